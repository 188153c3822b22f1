"""The fused verb b200_topo_geo2rdr_run (topo, then geo2rdr jobs on the layers still resident in HBM) against the two
separate verbs it replaces -- the reference's sequence topo_Py -> .rdr files -> geo2rdr_Py
(components/isceobj/StripmapProc/runTopo.py:66-103, runGeo2rdr.py:57-110; contrib/stack/topsStack/geo2rdr.py:233-302
for one geometry x several secondary dates).  Outputs must be bit-identical to the unfused path, which the other
parity tests hold against the CPU oracle; one case is also compared with the oracle directly."""
import datetime
import os

import numpy as np
import pytest

import isce2_b200
from isce2_b200 import _capi, image as IF, synth
from isce2_b200.orbit import Orbit
from isce2_b200.planet import Planet
from oracle import oracle as orc
from tests import parity_util as pu
from tests.test_gpu_components import _orbit, _write_dem

pytestmark = pytest.mark.gpu

TOPO_KEYS = ("lat", "lon", "hgt", "los", "inc", "mask")
GEO_KEYS = ("azt", "rgm", "azoff", "rgoff")


def _topo_params(sc, dem_method, line0=0, nlines=-1):
    return _capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                             delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
                             side=sc.side, peg_heading=sc.peg_heading, a=sc.a, e2=sc.e2, dem_method=dem_method,
                             line0=line0, nlines=nlines)


def _geo_params(sc, kw, out_f32, line0=0, nlines=-1):
    return _capi.geo_params(length=kw["length"], width=kw["width"], dem_shape=(sc.length, sc.width), r0=kw["r0"], dr=kw["dr"],
                            prf=kw["prf"], t0=kw["t0"], wvl=kw["wvl"], side=kw["side"], line0=line0, nlines=nlines,
                            out_f32=out_f32)


def _secondaries(sc):
    out = []
    for seed, f32, want in ((7, True, ("azoff", "rgoff")), (11, False, GEO_KEYS)):
        sec = synth.config_c1_secondary(length=sc.length, width=sc.width, seed=seed)
        out.append((pu.secondary_kwargs(sc, sec, recenter=0.37), f32, want))
    return out


@pytest.mark.parametrize("dem_method,line0,nlines", [("BIQUINTIC", 0, -1), ("BILINEAR", 13, 30)])
def test_fused_equals_separate_verbs(dem_method, line0, nlines):
    sc = pu.rough_scene(64, 3072)
    slr = [[sc.r0, sc.dr * sc.nrnglooks]]
    secs = _secondaries(sc)
    tp = _topo_params(sc, dem_method, line0, nlines)
    jobs = [dict(params=_geo_params(sc, kw, f32), orbit=(kw["orbit_t"], kw["orbit_pos"], kw["orbit_vel"]), want=want)
            for kw, f32, want in secs]
    ft, fg = _capi.topo_geo2rdr_run(tp, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, jobs, slr,
                                    want_los=True, want_inc=True, want_mask=True)
    st = _capi.topo_run(tp, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, slr, want_los=True,
                        want_inc=True, want_mask=True)
    n = sc.length if nlines < 0 else nlines
    assert ft["lat"].shape == (n, sc.width)
    for k in TOPO_KEYS:
        assert np.array_equal(ft[k], st[k], equal_nan=True), k
    for k in ("min_lat", "max_lat", "min_lon", "max_lon", "converged", "iterations"):
        assert ft[k] == st[k], k
    # the separate verb reads whole images and selects the block itself
    full = {k: np.zeros((sc.length, sc.width)) for k in ("lat", "lon", "hgt")}
    for k in full:
        full[k][line0:line0 + n] = st[k]
    for (kw, f32, want), g in zip(secs, fg):
        sg = _capi.geo2rdr_run(_geo_params(sc, kw, f32, line0, n), full["lat"], full["lon"], full["hgt"], kw["orbit_t"],
                               kw["orbit_pos"], kw["orbit_vel"], want=want)
        for k in GEO_KEYS:
            if k in want:
                assert g[k].dtype == (np.float32 if f32 else np.float64) and g[k].shape == (n, sc.width)
                assert np.array_equal(g[k], sg[k]), k
            else:
                assert g[k] is None
        for k in ("num_outside", "num_valid", "num_converged", "iterations"):
            assert g[k] == sg[k], k
        assert g["num_valid"] > 0.5 * n * sc.width
    # and against the oracle on the oracle's own layers (the unfused verbs are pinned to it elsewhere)
    if line0 == 0 and nlines < 0:
        kw = secs[1][0]
        o = orc.geo2rdr(lat=st["lat"], lon=st["lon"], hgt=st["hgt"], **kw)
        s2 = pu.compare_geo(fg[1], o)
        assert s2["valid"]["gpu"] == s2["valid"]["cpu"]
        assert s2["azoff"]["max"] < pu.TOL_OFFSET_PX and s2["rgoff"]["max"] < pu.TOL_OFFSET_PX
        assert s2["azoff"]["n_valid_mismatch"] == 0


def test_fused_argument_errors():
    sc = pu.rough_scene(16, 1024)
    slr = [[sc.r0, sc.dr * sc.nrnglooks]]
    kw, _, _ = _secondaries(sc)[0]
    tp = _topo_params(sc, "BILINEAR")
    orbit = (kw["orbit_t"], kw["orbit_pos"], kw["orbit_vel"])
    bad = _capi.geo_params(length=kw["length"], width=kw["width"], dem_shape=(sc.length, sc.width - 1), r0=kw["r0"], dr=kw["dr"],
                           prf=kw["prf"], t0=kw["t0"], wvl=kw["wvl"], side=kw["side"])
    with pytest.raises(_capi.B200Error) as ei:
        _capi.topo_geo2rdr_run(tp, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs,
                               [dict(params=bad, orbit=orbit)], slr)
    assert ei.value.code == -1 and "topo grid" in str(ei.value)
    good = _geo_params(sc, kw, True)
    with pytest.raises(_capi.B200Error) as ei:  # Geo2rdr.py:271-274
        _capi.topo_geo2rdr_run(tp, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs,
                               [dict(params=good, orbit=orbit, want=())], slr)
    assert "No outputs requested" in str(ei.value)
    with pytest.raises(_capi.B200Error) as ei:  # too few state vectors for Hermite (geo2rdr.f90:137-143)
        _capi.topo_geo2rdr_run(tp, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs,
                               [dict(params=good, orbit=tuple(a[:3] for a in orbit))], slr)
    assert ei.value.code == -4
    # no jobs: plain topo
    ft, fg = _capi.topo_geo2rdr_run(tp, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [], slr)
    st = _capi.topo_run(tp, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, slr)
    assert fg == [] and np.array_equal(ft["hgt"], st["hgt"])


def _configure_topo(sc, dem, d):
    topo = isce2_b200.createTopozero()
    topo.slantRangePixelSpacing, topo.prf, topo.radarWavelength = sc.dr, sc.prf, sc.wvl
    topo.orbit = _orbit(sc)
    topo.width, topo.length = sc.width, sc.length
    topo.wireInputPort(name="dem", object=dem)
    topo.wireInputPort(name="planet", object=Planet(pname="Earth"))
    topo.lookSide = sc.side
    topo.sensingStart = sc.sensing_start
    topo.rangeFirstSample = sc.r0
    topo.numberRangeLooks = topo.numberAzimuthLooks = 1
    topo.demInterpolationMethod = "BIQUINTIC"
    topo.latFilename, topo.lonFilename, topo.heightFilename, topo.losFilename = (str(d / f) for f in
                                                                                  ("lat.rdr", "lon.rdr", "z.rdr", "los.rdr"))
    topo.maskFilename = str(d / "shadowMask.rdr")
    return topo


def _configure_geo(sc, kw, sec, d, precision):
    grdr = isce2_b200.createGeo2rdr()
    grdr.configure()
    grdr.slantRangePixelSpacing, grdr.prf, grdr.radarWavelength = sc.dr, sc.prf, sc.wvl
    day = sc.sensing_start.replace(hour=0, minute=0, second=0, microsecond=0)
    grdr.orbit = Orbit.from_arrays(day, sec.orbit_t, sec.orbit_pos, sec.orbit_vel)
    grdr.width, grdr.length = sc.width, sc.length
    grdr.wireInputPort(name="planet", object=Planet(pname="Earth"))
    grdr.lookSide = sc.side
    grdr.setSensingStart(day + datetime.timedelta(seconds=kw["t0"]))
    grdr.rangeFirstSample = kw["r0"]
    grdr.numberRangeLooks = grdr.numberAzimuthLooks = 1
    grdr.dopplerCentroidCoeffs = [0.]
    grdr.rangeOffsetImageName = str(d / "range.off")
    grdr.azimuthOffsetImageName = str(d / "azimuth.off")
    grdr.outputPrecision = precision
    return grdr


@pytest.mark.parametrize("devices", [[0], [0, 0, 0]])
def test_chained_components_write_the_same_files(tmp_path, devices):
    """Topo.chainGeo2rdr(grdr) + topo.topo() == topo.topo() then grdr.geo2rdr() on its files, byte for byte, XML included."""
    sc = pu.rough_scene(45, 2048)
    dem, _ = _write_dem(sc, str(tmp_path / "dem.dem"))
    sec = synth.config_c1_secondary(length=sc.length, width=sc.width)
    kw = pu.secondary_kwargs(sc, sec, recenter=0.37)

    a = tmp_path / "separate"
    topo = _configure_topo(sc, dem, a)
    topo.topo()
    grdr = _configure_geo(sc, kw, sec, a, "single")
    for attr, f in (("demImage", "z.rdr"), ("latImage", "lat.rdr"), ("lonImage", "lon.rdr")):
        img = IF.createImage()
        img.load(str(a / (f + ".xml")))
        img.setAccessMode("READ")
        setattr(grdr, attr, img)
    grdr.geo2rdr()

    b = tmp_path / "chained"
    topo2 = _configure_topo(sc, dem, b)
    topo2.gpuDevices = devices
    grdr2 = topo2.chainGeo2rdr(_configure_geo(sc, kw, sec, b, "single"))
    topo2.topo()
    for f in ("lat.rdr", "lon.rdr", "z.rdr", "los.rdr", "shadowMask.rdr", "range.off", "azimuth.off"):
        assert np.array_equal(np.fromfile(a / f, np.uint8), np.fromfile(b / f, np.uint8)), f
        assert os.path.exists(b / (f + ".xml")) and os.path.exists(b / (f + ".vrt"))
    ha, hb = (IF.createImage().load(str(d / "range.off.xml")) for d in (a, b))
    assert (ha.dataType, ha.width, ha.length, ha.bands) == (hb.dataType, hb.width, hb.length, hb.bands) == ("FLOAT", sc.width, sc.length, 1)
    assert (grdr2.numValid, grdr2.numOutsideImage, grdr2.numConverged) == (grdr.numValid, grdr.numOutsideImage, grdr.numConverged)
    assert topo2.snwe == topo.snwe and topo2.totalConverged == topo.totalConverged
    assert grdr2.numValid > 0.5 * sc.length * sc.width


def test_frozen_stack_geometry_is_bit_identical():
    """b200_geo_plan_freeze_geometry: the ECEF copy of a fixed reference geometry serves every secondary date; outputs
    equal the per-pixel LLH -> XYZ path bit for bit; another ellipsoid falls back to that path."""
    sc = pu.rough_scene(40, 2048)
    c = pu.cpu_topo(sc, dem_method="BILINEAR", want_inc=False, want_mask=False)
    outs = {}
    for frozen in (False, True):
        kw0 = _secondaries(sc)[0][0]
        plan = _capi.GeoPlan(_geo_params(sc, kw0, False), lat=c["lat"], lon=c["lon"], hgt=c["hgt"])
        if frozen:
            plan.freeze_geometry()
        for i, (kw, f32, want) in enumerate(_secondaries(sc)):
            p = _geo_params(sc, kw, f32)
            plan.execute(p, kw["orbit_t"], kw["orbit_pos"], kw["orbit_vel"], want=GEO_KEYS)
            outs[(frozen, i)] = plan.fetch()
        # a sphere of the same radius: not the frozen ellipsoid
        p2 = _capi.geo_params(length=kw0["length"], width=kw0["width"], dem_shape=(sc.length, sc.width), r0=kw0["r0"], dr=kw0["dr"],
                              prf=kw0["prf"], t0=kw0["t0"], wvl=kw0["wvl"], side=kw0["side"], e2=0.0)
        plan.execute(p2, kw0["orbit_t"], kw0["orbit_pos"], kw0["orbit_vel"], want=("rgm",))
        outs[(frozen, "sphere")] = plan.fetch()
        plan.close()
    for i in (0, 1):
        for k in GEO_KEYS:
            assert np.array_equal(outs[(False, i)][k], outs[(True, i)][k]), (i, k)
        assert outs[(False, i)]["num_valid"] == outs[(True, i)]["num_valid"] > 0
    assert np.array_equal(outs[(False, "sphere")]["rgm"], outs[(True, "sphere")]["rgm"])
    assert not np.array_equal(outs[(True, "sphere")]["rgm"], outs[(True, 0)]["rgm"])
