"""The per-pixel device functions (isce2_b200/csrc/*.cuh), compiled for the host by g++, against the oracle.

With the generic libm trigonometry (use_ref=False) the kernel arithmetic is the reference's own operation sequence
(no fused multiply-adds: the CUDA build uses -fmad=false; div_r / div_n / sqrt_n give IEEE results) except for the
re-associated cube-root term of XYZ->LLH, whose effect on the latitude is damped 300x: iteration counts must be
identical and >= 80 % of the outputs bit-identical, the rest within an ulp or two.  With the reference-angle
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
        assert (d == 0).mean() > (0.5 if k == "hgt" else 0.8), (k, (d == 0).mean())
    for k in ("los", "inc"):
        assert (o[k] != e[k]).mean() < 1e-4, k


@pytest.mark.parametrize("name,mid", [("BILINEAR", 1), ("BICUBIC", 2), ("NEAREST", 3), ("BIQUINTIC", 5)])
def test_pixel_functions_bit_exact(name, mid):
    sc = pu.rough_scene(12, 2048, dem_spacing_arcsec=1.0)
    o = orc.topo(**orc.scene_topo_kwargs(sc, dem_method=name, want_mask=False))
    for use_ref in (False, True):
        e = emu.topo(sc, o, dem_method=mid, use_ref=use_ref)
        _check(o, e)


def test_native_doppler_and_left_looking_bit_exact():
    sc = synth.make_scene(8, 2048, sensor="nisar")
    o = orc.topo(**orc.scene_topo_kwargs(sc, dem_method="BIQUINTIC", orbit_method="LEGENDRE", want_mask=False))
    for use_ref in (False, True):
        _check(o, emu.topo(sc, o, dem_method=5, orbit_method=2, use_ref=use_ref))
