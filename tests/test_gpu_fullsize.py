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
    # ---- strips of the full run against the oracle: first burst, across a burst boundary, last burst (258 lines) ----
    tot = {}
    for a, nl in ((0, 86), (1457, 86), (sc.length - 86, 86)):
        c = pu.cpu_topo(sc, dem_method="BIQUINTIC", line0=a, nlines=nl)
        strip = {k: (full[k][a:a + nl] if isinstance(full[k], np.ndarray) else full[k]) for k in full}
        assert np.array_equal(strip["mask"], c["mask"])
        _accumulate(tot, strip, c)
    _report("c2_full_size_3_strips", tot)
    _assert_strip_totals(tot)


def _accumulate(tot, g, c):
    """Differences of one strip, added to the running totals (pixels, pixels over tolerance, largest difference)."""
    for k, tol in (("lat", pu.TOL_LATLON_DEG), ("lon", pu.TOL_LATLON_DEG), ("hgt", pu.TOL_HGT_M), ("los", pu.TOL_ANGLE_DEG),
                   ("inc", pu.TOL_ANGLE_DEG)):
        d = np.abs(g[k].astype(np.float64) - c[k].astype(np.float64))
        if k == "los":
            d = np.minimum(d, np.abs(d - 360.0))
        t = tot.setdefault(k, dict(n=0, n_over=0, max=0.0, n_exact=0, tol=tol))
        t["n"] += int(d.size)
        t["n_over"] += int((d > tol).sum())
        t["n_exact"] += int((d == 0).sum())
        t["max"] = max(t["max"], float(d.max()))


def _report(name, tot):
    """Outliers per 1e9 values, printed and left in gpurun_out/ for the profile summary."""
    import json
    import os
    rep = {k: dict(v, over_per_1e9=1e9 * v["n_over"] / v["n"], exact_fraction=v["n_exact"] / v["n"]) for k, v in tot.items()}
    print(name, json.dumps(rep))
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open(os.path.join("gpurun_out", f"parity_{name}.json"), "w") as f:
            json.dump(rep, f, indent=1)
    except OSError:
        pass


def _assert_strip_totals(tot):
    # hgt within 1 cm everywhere; lat / lon / angles within tolerance except for the float32 DEM-index flips described in
    # tests/test_gpu_parity.py, bounded by 10x the tolerance in lat / lon and 5e-3 deg in the angles and counted per 1e9 values in the report
    assert tot["hgt"]["n_over"] == 0, tot["hgt"]
    for k in ("lat", "lon"):
        assert tot[k]["n_over"] <= max(2, int(2e-5 * tot[k]["n"])) and tot[k]["max"] < 1e-7, (k, tot[k])
    for k in ("los", "inc"):
        assert tot[k]["n_over"] <= max(4, int(4e-5 * tot[k]["n"])) and tot[k]["max"] < 5e-3, (k, tot[k])


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
    # ---- a 64-line block from the middle of the full-size frame (global DEM crop, full-size geometry) against the oracle ----
    a, nl = 30000, 64
    p.line0, p.nlines = a, nl
    blk = _capi.topo_run(p, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [[sc.r0, sc.dr]], want_los=True,
                         want_inc=True, want_mask=True)
    c = pu.cpu_topo(sc, dem_method="BIQUINTIC", orbit_method="LEGENDRE", line0=a, nlines=nl)
    assert np.array_equal(blk["mask"], c["mask"]) and blk["iterations"] == c["total_iters"]
    assert [blk[k] for k in ("dem_x0", "dem_y0", "dem_nx", "dem_ny")] == [c[k] for k in ("ustartx", "ustarty", "udemwidth", "udemlength")]
    tot = {}
    _accumulate(tot, blk, c)
    _report("c3_full_size_strip", tot)
    _assert_strip_totals(tot)
    kw = dict(orbit_t=sc.orbit_t, orbit_pos=sc.orbit_pos, orbit_vel=sc.orbit_vel, length=sc.length, width=sc.width, r0=sc.r0,
              dr=sc.dr, prf=sc.prf, t0=sc.t0, wvl=sc.wvl, side=sc.side)
    from oracle import oracle as orc
    gblk = _capi.geo2rdr_run(_capi.geo_params(length=sc.length, width=sc.width, dem_shape=(sc.length, sc.width), r0=sc.r0, dr=sc.dr,
                                              prf=sc.prf, t0=sc.t0, wvl=sc.wvl, side=sc.side, orbit_method="LEGENDRE", line0=a,
                                              nlines=nl), c["lat"], c["lon"], c["hgt"], sc.orbit_t, sc.orbit_pos, sc.orbit_vel,
                             doppler_coeffs=dop, block_rows=True)
    # the oracle sees the block as an image whose first row is radar line a (offsets are relative to the row index)
    o = orc.geo2rdr(lat=c["lat"], lon=c["lon"], hgt=c["hgt"], orbit_method="LEGENDRE", doppler_coeffs=dop,
                    **dict(kw, t0=sc.t0 + a / sc.prf, length=sc.length - a))
    # same orbit, same window: the pixels of the first / last sample sit ON the edge of the range window and the strict
    # bounds test (geo2rdr.f90:308-316) decides them in the last bit, which the re-based description of the block moves
    v = o["azoff"] != -999999.0
    gv = gblk["azoff"] != -999999.0
    assert np.array_equal(gv[:, 1:-1], v[:, 1:-1]) and v[:, 1:-1].all()
    v = v & gv
    assert np.abs(gblk["rgoff"][v] - o["rgoff"][v]).max() < pu.TOL_OFFSET_PX
    assert np.abs(gblk["azt"][v] - o["azt"][v]).max() < pu.TOL_OFFSET_PX / sc.prf
