"""The per-pixel device functions (isce2_b200/csrc/*.cuh), compiled for the host by g++, against the oracle.

With the generic libm trigonometry (use_ref=False) the kernel arithmetic is the reference's own operation sequence
(no fused multiply-adds: the CUDA build uses -fmad=false; div_r / div_n / sqrt_n give IEEE results) except for the
re-associated cube-root term of XYZ->LLH, whose effect on the latitude is damped 300x: iteration counts must be
identical and >= 70 % of the outputs bit-identical, the rest within an ulp or two.  With the reference-angle
trigonometry (use_ref=True, the production path) the same bounds must hold.  This is a development aid for a
container without a GPU; the product never loads it."""
import numpy as np
import pytest

from isce2_b200 import synth
from oracle import oracle as orc
from tests import parity_util as pu
from tests.emu import emu


def _check(o, e):
    assert o["total_iters"] == e["iters"]
    for k, tol in (("lat", 1e-12), ("lon", 1e-12), ("hgt", 1e-7)):
        d = np.abs(o[k] - e[k])
        assert d.max() < tol, (k, d.max())
        assert (d == 0).mean() > (0.5 if k == "hgt" else 0.7), (k, (d == 0).mean())
    for k in ("los", "inc"):
        assert (o[k] != e[k]).mean() < 1e-4, k


@pytest.mark.parametrize("name,mid", [("BILINEAR", 1), ("BICUBIC", 2), ("NEAREST", 3), ("BIQUINTIC", 5), ("SINC", 0), ("AKIMA", 4)])
def test_pixel_functions_bit_exact(name, mid):
    sc = pu.rough_scene(12, 2048, dem_spacing_arcsec=1.0)
    o = orc.topo(**orc.scene_topo_kwargs(sc, dem_method=name, want_mask=False))
    for use_ref in (0, 1, 2):  # libm, reference-angle series, truncated series of narrow blocks
        e = emu.topo(sc, o, dem_method=mid, use_ref=use_ref)
        _check(o, e)


def test_native_doppler_and_left_looking_bit_exact():
    sc = synth.make_scene(8, 2048, sensor="nisar")
    o = orc.topo(**orc.scene_topo_kwargs(sc, dem_method="BIQUINTIC", orbit_method="LEGENDRE", want_mask=False))
    for use_ref in (0, 1, 2):
        _check(o, emu.topo(sc, o, dem_method=5, orbit_method=2, use_ref=use_ref))


@pytest.mark.parametrize("method,name", [(0, "HERMITE"), (2, "LEGENDRE")])
def test_orbit_window_polynomials_match_reference_interpolators(method, name, golden):
    """geo2rdr evaluates the orbit through per-window polynomials (isce2_b200/csrc/orbit_poly.h) expanded from the
    reference's own Hermite / Lagrange formulas: they must reproduce orbit.c to rounding (~1e-8 m, 1e-9 m/s), on the
    real state vectors of the reference's orbit test fixture and on a synthetic Keplerian orbit, at nodes, between
    nodes and in the clamped end windows."""
    import ctypes as C
    L = emu.build()
    dp = C.POINTER(C.c_double)
    rows = np.array(golden["orbit_rsc"])
    sc = synth.make_scene(1500, 64, dem=False)
    for t, pos, vel in ((rows[:, 0].copy(), rows[:, 1:4].copy(), rows[:, 4:7].copy()), (sc.orbit_t, sc.orbit_pos, sc.orbit_vel)):
        t, pos, vel = (np.ascontiguousarray(a, np.float64) for a in (t, pos, vel))
        rng = np.random.default_rng(3)
        tq = np.concatenate([rng.uniform(t[0], t[-1], 500), t, [t[0] + 1e-7, t[-1] - 1e-7]])
        out = np.zeros((len(tq), 9))
        assert L.emu_orbit_poly(method, len(t), t.ctypes.data_as(dp), pos.ctypes.data_as(dp), vel.ctypes.data_as(dp), len(tq),
                                tq.ctypes.data_as(dp), out.ctypes.data_as(dp)) == 0
        o = orc.Orbit(t, pos, vel)
        for q, tt in enumerate(tq):
            stat, p, v = o.interp(tt, name)
            assert stat == 0
            assert np.abs(out[q, :3] - p).max() < 2e-8, (tt, out[q, :3] - p)
            assert np.abs(out[q, 3:6] - v).max() < 1e-9, (tt, out[q, 3:6] - v)
        # acceleration = derivative of the velocity: compare with a central difference of the reference velocity
        tt = 0.5 * (t[0] + t[-1]) + 0.37
        out1 = np.zeros((1, 9))
        tq1 = np.array([tt])
        L.emu_orbit_poly(method, len(t), t.ctypes.data_as(dp), pos.ctypes.data_as(dp), vel.ctypes.data_as(dp), 1,
                         tq1.ctypes.data_as(dp), out1.ctypes.data_as(dp))
        fd = (o.interp(tt + 0.01, name)[2] - o.interp(tt - 0.01, name)[2]) / 0.02
        assert np.abs(out1[0, 6:9] - fd).max() < 1e-5
