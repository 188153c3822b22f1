"""Size-independent properties at BASELINE.json's full sizes (the oracle would need hours there):

* configs[2] (13500 x 25000, BIQUINTIC + incidence + mask): topo followed by geo2rdr with the SAME orbit and timing must
  return every pixel to its own (line, sample): |azimuth offset|, |range offset| < 1e-3 pixel, all pixels valid;
* the swath computed as two azimuth line blocks (what N GPUs do) is bit-identical to the single run (lat + mask
  checksums), and the bounding box / convergence counters add up;
* a strip of the full-size run equals the oracle run on that strip alone (lines are independent, DEM crop is global).
"""
import zlib

import numpy as np
import pytest

from isce2_b200 import _capi, synth
from tests import parity_util as pu

pytestmark = pytest.mark.gpu


def _params(sc, line0=0, nlines=-1, dem_method="BIQUINTIC"):
    return _capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                             delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
                             side=sc.side, peg_heading=sc.peg_heading, dem_method=dem_method, line0=line0, nlines=nlines)


def _crc(a):
    return zlib.crc32(np.ascontiguousarray(a).view(np.uint8).reshape(-1))


def test_c2_full_swath_round_trip_and_line_block_identity():
    sc = synth.config_c2()
    assert (sc.length, sc.width) == (13500, 25000)
    slr = [[sc.r0, sc.dr]]
    tp = _capi.TopoPlan(_params(sc), sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, slr,
                        want_los=True, want_inc=True, want_mask=True)
    tp.execute()
    gp = _capi.geo_params(length=sc.length, width=sc.width, dem_shape=(sc.length, sc.width), r0=sc.r0, dr=sc.dr, prf=sc.prf,
                          t0=sc.t0, wvl=sc.wvl, side=sc.side, out_f32=True)
    g = _capi.GeoPlan(gp, topo_plan=tp)
    g.execute(gp, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, want=("azoff", "rgoff"))
    r = g.fetch()
    n = sc.length * sc.width
    # pixels ON the edge of the acquisition window (first / last line, first / last sample) solve to t = tstart or
    # rng = rngstart up to rounding, and the reference's strict bounds test (geo2rdr.f90:308-316) rejects the ones that
    # land an ulp outside; nothing else may be invalid
    bad = r["azoff"] == np.float32(-999999.0)
    assert np.array_equal(bad, r["rgoff"] == np.float32(-999999.0))
    assert int(bad.sum()) == n - r["num_valid"] and not bad[1:-1, 1:-1].any()
    assert float(np.abs(r["azoff"][~bad]).max()) < pu.TOL_OFFSET_PX
    # float32 offsets of a pixel index up to 25000 carry 2e-3 of rounding in (rng - r0)/dr - pixel; the double-precision
    # residual is checked on the range itself below
    assert float(np.abs(r["rgoff"][~bad]).max()) < 4e-3
    del r
    gp64 = _capi.geo_params(length=sc.length, width=sc.width, dem_shape=(sc.length, sc.width), r0=sc.r0, dr=sc.dr, prf=sc.prf,
                            t0=sc.t0, wvl=sc.wvl, side=sc.side, out_f32=False)
    g.execute(gp64, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, want=("rgm",))
    rg = g.fetch()["rgm"]
    want = sc.r0 + sc.dr * np.arange(sc.width)
    assert float(np.abs(rg - want[None, :])[~bad].max()) < pu.TOL_OFFSET_PX * sc.dr
    del rg
    g.close()
    full = tp.fetch()
    tp.close()
    assert full["converged"] > 0.999 * n
    hist = np.bincount(full["mask"].ravel().astype(np.uint8), minlength=4)
    assert hist.sum() == n and hist[0] > 0.5 * n
    assert np.isfinite(full["lat"]).all() and np.isfinite(full["hgt"]).all()
    # ---- two line blocks == one run ----
    half = sc.length // 2
    conv = iters = 0
    bbox = [1e300, -1e300, 1e300, -1e300]
    for a, b in ((0, half), (half, sc.length)):
        bp = _capi.TopoPlan(_params(sc, a, b - a), sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, slr,
                            want_los=True, want_inc=True, want_mask=True)
        bp.execute()
        blk = bp.fetch()
        bp.close()
        for k in ("lat", "lon", "hgt", "los", "inc", "mask"):
            assert _crc(blk[k]) == _crc(full[k][a:b]), (k, a, b)
        conv += blk["converged"]
        iters += blk["iterations"]
        bbox = [min(bbox[0], blk["min_lat"]), max(bbox[1], blk["max_lat"]), min(bbox[2], blk["min_lon"]), max(bbox[3], blk["max_lon"])]
        assert [blk[k] for k in ("dem_x0", "dem_y0", "dem_nx", "dem_ny")] == [full[k] for k in ("dem_x0", "dem_y0", "dem_nx", "dem_ny")]
    assert conv == full["converged"] and iters == full["iterations"]
    assert bbox == [full["min_lat"], full["max_lat"], full["min_lon"], full["max_lon"]]
    # ---- a strip of the full run against the oracle ----
    a, nl = 9000, 6
    c = pu.cpu_topo(sc, dem_method="BIQUINTIC", line0=a, nlines=nl)
    strip = {k: (full[k][a:a + nl] if isinstance(full[k], np.ndarray) else full[k]) for k in full}
    assert np.array_equal(strip["mask"], c["mask"])
    assert np.abs(strip["hgt"] - c["hgt"]).max() < pu.TOL_HGT_M
    assert (np.abs(strip["lat"] - c["lat"]) > pu.TOL_LATLON_DEG).sum() <= 2 and np.abs(strip["lat"] - c["lat"]).max() < 2e-7
    assert (np.abs(strip["lon"] - c["lon"]) > pu.TOL_LATLON_DEG).sum() <= 2 and np.abs(strip["lon"] - c["lon"]).max() < 2e-7


def test_c3_nisar_frame_round_trip_at_full_size():
    """BASELINE configs[3]: 60000 x 25000 (1.5 Gpixel), left-looking, native Doppler, Legendre orbit, BIQUINTIC +
    incidence + mask: ~90 GB resident on one B200.  Same closure property as above, with the Doppler polynomial of the
    scene handed to geo2rdr in its own convention (cycles / PRF versus range pixel, Geo2rdr.py:264-297)."""
    sc = synth.config_c3()
    assert (sc.length, sc.width) == (60000, 25000)
    p = _capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                          delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
                          side=sc.side, peg_heading=sc.peg_heading, dem_method="BIQUINTIC", orbit_method="LEGENDRE")
    tp = _capi.TopoPlan(p, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [[sc.r0, sc.dr]],
                        want_los=True, want_inc=True, want_mask=True)
    ms = tp.execute()
    gp = _capi.geo_params(length=sc.length, width=sc.width, dem_shape=(sc.length, sc.width), r0=sc.r0, dr=sc.dr, prf=sc.prf,
                          t0=sc.t0, wvl=sc.wvl, side=sc.side, orbit_method="LEGENDRE", out_f32=True)
    g = _capi.GeoPlan(gp, topo_plan=tp)
    dop = [d / sc.prf for d in sc.doppler_coeffs[0]]
    g.execute(gp, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, doppler_coeffs=dop, want=("azoff",))
    r = g.fetch()
    n = sc.length * sc.width
    bad = r["azoff"] == np.float32(-999999.0)
    assert int(bad.sum()) == n - r["num_valid"] and not bad[1:-1, 1:-1].any()
    assert float(np.abs(r["azoff"][~bad]).max()) < pu.TOL_OFFSET_PX
    g.close()
    res = _capi.TopoResult()
    import ctypes as C
    e = C.create_string_buffer(512)
    _capi._check(_capi.lib().b200_topo_plan_fetch(tp.handle, None, C.byref(res), e, 512), e)
    tp.close()
    assert res.converged > 0.999 * n and 3.0 < res.iterations / n < 12.0
    print("c3 topo device ms", ms)
