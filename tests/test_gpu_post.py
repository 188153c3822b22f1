"""GPU parity of the multilooking and mask-projection kernels (other consumers) against the oracle
restatements, which tests/test_post_oracle_cpu.py pins to the reference's own compiled takeLooks<T> templates and to
golden output of the reference's SWBDStitcher.toRadar.  Integer / byte work: bit-exact; so are the float means (same
additions in the same order, one IEEE division)."""
import os

import numpy as np
import pytest

import isce2_b200
from isce2_b200 import _capi, image as IF, looks as LK, watermask as WM
from isce2_b200.planet import Planet
from oracle import oracle as orc
from tests import parity_util as pu
from tests.test_gpu_components import _orbit, _write_dem
from tests.test_post_oracle_cpu import DTYPES, ROOT, _random

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt,code", DTYPES)
def test_looks_matches_oracle_bit_for_bit(dt, code):
    rng = np.random.default_rng(100 + code)
    cases = [((37, 61), "BIL", 3, 4), ((40, 2, 130), "BIL", 4, 7), ((33, 50, 3), "BIP", 5, 2), ((2, 29, 77), "BSQ", 2, 5),
             ((9, 9), "BIL", 1, 1), ((20, 4500), "BIL", 2, 9), ((6, 12), "BIL", 7, 2), ((6, 12), "BIL", 2, 13)]
    for shape, scheme, ld, la in cases:
        a = _random(rng, shape, dt)
        for method in ("AVERAGE", "NEAREST"):
            g, res = _capi.looks_run(a, ld, la, scheme=scheme, method=method)
            c = orc.looks(a, ld, la, scheme=scheme, method=method)
            assert g.shape == c.shape and g.dtype == c.dtype, (shape, scheme, ld, la, method)
            assert np.array_equal(g.view(np.uint8), c.view(np.uint8)), (dt, shape, scheme, ld, la, method)
            if g.size:
                assert res["gpu_launches"] >= 1


def test_looks_full_size_properties():
    """A burst-sized two-band float layer (los.rdr: 1500 x 2 x 25000 BIL) with stripmapStack-like looks: several pipeline
    blocks; the mean of a constant is the constant, the mean is linear, sampled blocks equal the oracle."""
    L, W, ld, la = 1500, 25000, 14, 4
    rng = np.random.default_rng(5)
    a = rng.normal(size=(L, 2, W)).astype(np.float32)
    g, res = _capi.looks_run(a, ld, la, scheme="BIL")
    assert g.shape == (L // ld, 2, W // la) and res["gpu_launches"] > 1
    for r0 in (0, 53, L // ld - 1):
        c = orc.looks(a[r0 * ld:(r0 + 1) * ld], ld, la, scheme="BIL")
        assert np.array_equal(g[r0:r0 + 1], c)
    k, _ = _capi.looks_run(np.full((L, W), 3.25, np.float64), ld, la)
    assert np.all(k == 3.25)
    m, _ = _capi.looks_run(np.full((L, W), -7, np.int8), ld, la)
    assert np.all(m == -7)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_mask_projection_matches_golden_and_oracle(dt):
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_toradar.npz"))
    d = float(g["delta"])
    lat, lon = g["lat"].astype(dt), g["lon"].astype(dt)
    out, res = _capi.mask_to_radar_run(g["mask"], float(g["start_lat"]), -d, float(g["start_lon"]), d, lat, lon)
    assert np.array_equal(out, g["out_f64" if dt == np.float64 else "out_f32"]) and res["gpu_launches"] == 1
    # a larger case with other mask types, NaN / infinite / far-away coordinates (clipped like numpy's astype(int) + clip)
    rng = np.random.default_rng(9)
    for mt in (np.int8, np.int16, np.int32, np.float32):
        mask = _random(rng, (301, 407), mt)
        la2 = (35.0 - 1e-3 * rng.uniform(-20, 320, (700, 900))).astype(dt)
        lo2 = (-118.0 + 1e-3 * rng.uniform(-20, 430, (700, 900))).astype(dt)
        la2[0, :4] = [np.nan, np.inf, -np.inf, 1e30]
        lo2[1, :4] = [np.nan, np.inf, -np.inf, -1e30]
        with np.errstate(invalid="ignore"):
            c = orc.mask_to_radar(mask, 35.0, -1e-3, -118.0, 1e-3, la2, lo2)
        o2, _ = _capi.mask_to_radar_run(mask, 35.0, -1e-3, -118.0, 1e-3, la2, lo2)
        assert np.array_equal(o2.view(np.uint8), c.view(np.uint8)), mt


def test_multilook_and_water_mask_through_the_host_mirrors(tmp_path):
    """topo() -> runMultilook of its layers (both methods) and toRadar of a geocoded mask through lat.rdr / lon.rdr: the
    call sequence of contrib/stack/stripmapStack/topo.py:365-441 and createWaterMask.py:66-71."""
    sc = pu.rough_scene(60, 1024)
    dem, _ = _write_dem(sc, str(tmp_path / "dem.dem"))
    geom = tmp_path / "geom_reference_full"
    topo = isce2_b200.createTopozero()
    topo.slantRangePixelSpacing, topo.prf, topo.radarWavelength = sc.dr, sc.prf, sc.wvl
    topo.orbit = _orbit(sc)
    topo.width, topo.length = sc.width, sc.length
    topo.wireInputPort(name="dem", object=dem)
    topo.wireInputPort(name="planet", object=Planet(pname="Earth"))
    topo.lookSide, topo.sensingStart, topo.rangeFirstSample = sc.side, sc.sensing_start, sc.r0
    topo.numberRangeLooks = topo.numberAzimuthLooks = 1
    topo.latFilename, topo.lonFilename, topo.heightFilename = (str(geom / f) for f in ("lat.rdr", "lon.rdr", "hgt.rdr"))
    topo.losFilename, topo.incFilename, topo.maskFilename = (str(geom / f) for f in ("los.rdr", "incLocal.rdr", "shadowMask.rdr"))
    topo.topo()

    # water mask on the DEM grid: "water" below the scene's median height (SWBD convention -1 water / 0 land)
    wb = np.where(sc.dem < np.median(np.fromfile(geom / "hgt.rdr")), -1, 0).astype(np.int8)
    wpath = str(tmp_path / "swbd.wbd")
    wb.tofile(wpath)
    wim = IF.createImage()
    wim.initImage(wpath, "read", wb.shape[1], "BYTE")
    wim.setLength(wb.shape[0])
    wim.coord1.coordStart, wim.coord1.coordDelta = sc.first_lon, sc.delta_lon
    wim.coord2.coordStart, wim.coord2.coordDelta = sc.first_lat, sc.delta_lat
    wim.coord1.coordSize, wim.coord2.coordSize = wb.shape[1], wb.shape[0]
    wim.renderHdr()
    WM.geo2radar(wpath, str(geom / "waterMask.rdr"), str(geom / "lat.rdr"), str(geom / "lon.rdr"))
    lat = np.fromfile(geom / "lat.rdr").reshape(sc.length, sc.width)
    lon = np.fromfile(geom / "lon.rdr").reshape(sc.length, sc.width)
    wm = np.fromfile(geom / "waterMask.rdr", np.int8).reshape(sc.length, sc.width)
    assert np.array_equal(wm, orc.mask_to_radar(wb, sc.first_lat, sc.delta_lat, sc.first_lon, sc.delta_lon, lat, lon))
    hdr = IF.createImage().load(str(geom / "waterMask.rdr.xml"))
    assert (hdr.dataType, hdr.width, hdr.length) == ("BYTE", sc.width, sc.length)
    assert 0 < (wm == 0).mean() < 1  # both water (0 after the +1) and land (1) occur

    for method in ("isce", "gdal"):
        out_dir = tmp_path / ("geom_reference_" + method)
        LK.runMultilook(str(geom), str(out_dir), 4, 3, method=method)
        for fbase, dt, bands in (("hgt", np.float64, 1), ("lat", np.float64, 1), ("lon", np.float64, 1), ("los", np.float32, 2),
                                 ("incLocal", np.float32, 2), ("shadowMask", np.int8, 1), ("waterMask", np.int8, 1)):
            full = np.fromfile(geom / (fbase + ".rdr"), dt).reshape((sc.length, sc.width) if bands == 1 else (sc.length, bands, sc.width))
            want = orc.looks(full, 4, 3, scheme="BIL", method="AVERAGE" if method == "isce" else "NEAREST")
            got = np.fromfile(out_dir / (fbase + ".rdr"), dt).reshape(want.shape)
            assert np.array_equal(got, want), (method, fbase)
            h = IF.createImage().load(str(out_dir / (fbase + ".rdr.xml")))
            assert (h.width, h.length, h.bands) == (sc.width // 3, sc.length // 4, bands)
            assert os.path.exists(out_dir / (fbase + ".rdr.full.xml")) and os.path.exists(out_dir / (fbase + ".rdr.full.vrt"))


def test_looks_of_resident_topo_layers():
    """b200_topo_plan_looks: multilooked layers straight from the plan's HBM copy == looks of the fetched layers."""
    sc = pu.rough_scene(50, 2048)
    p = _capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                          delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl, side=sc.side,
                          peg_heading=sc.peg_heading, dem_method="BILINEAR", line0=6, nlines=40)
    plan = _capi.TopoPlan(p, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [[sc.r0, sc.dr]], want_los=True,
                          want_inc=False, want_mask=True)
    with pytest.raises(_capi.B200Error):
        plan.looks("lat", 4, 3)  # not executed yet
    plan.execute()
    full = plan.fetch()
    for layer in ("lat", "lon", "hgt", "los", "mask"):
        for method in ("AVERAGE", "NEAREST"):
            g, res = plan.looks(layer, 7, 5, method=method)
            want = orc.looks(full[layer], 7, 5, scheme="BIL", method=method)
            assert g.shape == want.shape == ((5, 409) if layer != "los" else (5, 2, 409))
            assert np.array_equal(g, want), (layer, method)
            assert res["gpu_launches"] == 1 and res["ms_kernels"] > 0
    with pytest.raises(_capi.B200Error) as ei:
        plan.looks("inc", 2, 2)  # layer not requested
    assert ei.value.code == -1
    plan.close()
