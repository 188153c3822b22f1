"""Host logic of the chained (fused) topo -> geo2rdr path on CPU: Topo.chainGeo2rdr(), the per-device line blocks, the
output rasters and their XML -- with the one library call (b200_topo_geo2rdr_run) replaced, for this test only, by the
oracle evaluating the very arguments the component hands to the C ABI.  The hardware counterpart is tests/test_gpu_fused.py."""
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

DEM_NAMES = {v: k for k, v in _capi.DEM_METHODS.items()}
ORB_NAMES = {v: k for k, v in _capi.ORBIT_METHODS.items()}


def _poly2d_coeffs(p):
    n = (p.azimuth_order + 1) * (p.range_order + 1)
    return np.array([p.coeffs[i] for i in range(n)]).reshape(p.azimuth_order + 1, p.range_order + 1)


def _oracle_as_library(calls):
    def topo_geo2rdr_run(params, dem, t, pos, vel, dop, geo_jobs, slr=None, rho_image=None, want_los=True, want_inc=False,
                         want_mask=False, out=None, doppler_poly=None, slrng_poly=None):
        p = params
        sl = _poly2d_coeffs(slrng_poly)
        o = orc.topo(dem=np.asarray(dem, np.float32), first_lat=p.first_lat, first_lon=p.first_lon, delta_lat=p.delta_lat,
                     delta_lon=p.delta_lon, orbit_t=t, orbit_pos=pos, orbit_vel=vel, length=p.length, width=p.width,
                     r0=float(sl[0, 0]), dr=float(sl[0, 1]), prf=p.prf, t0=p.t0, wvl=p.wvl, side=p.look_side,
                     peg_heading=p.peg_heading, doppler_coeffs=_poly2d_coeffs(doppler_poly), a=p.major, e2=p.e2,
                     dem_method=DEM_NAMES[p.dem_method], orbit_method=ORB_NAMES[p.orbit_method], numiter=p.numiter,
                     extraiter=p.extraiter, thresh=p.thresh, want_inc=want_inc, want_mask=want_mask, line0=p.line0, nlines=p.nlines)
        for k in ("lat", "lon", "hgt", "los", "inc", "mask"):
            if out.get(k) is not None:
                out[k][...] = o[k]
        calls.append((p.line0, p.nlines, p.device))
        tres = dict(min_lat=o["min_lat"], max_lat=o["max_lat"], min_lon=o["min_lon"], max_lon=o["max_lon"], converged=o["totalconv"],
                    iterations=o["total_iters"], ms_setup=0.0, ms_kernels=0.0, ms_pixels=0.0, ms_mask=0.0, ms_total=0.0, gpu_launches=4)
        geos = []
        for jb in geo_jobs:
            q = jb["params"]
            assert (q.dem_length, q.dem_width, q.line0, q.nlines) == (p.length, p.width, p.line0, p.nlines)
            coeffs, mean, norm = jb["doppler"]
            g = orc.geo2rdr(lat=o["lat"], lon=o["lon"], hgt=o["hgt"], orbit_t=jb["orbit"][0], orbit_pos=jb["orbit"][1],
                            orbit_vel=jb["orbit"][2], length=q.length, width=q.width, r0=q.rho0, dr=q.drho, prf=q.prf, t0=q.t0,
                            wvl=q.wvl, side=q.look_side, doppler_coeffs=coeffs, doppler_mean=mean, doppler_norm=norm, a=q.major,
                            e2=q.e2, orbit_method=ORB_NAMES[q.orbit_method], bistatic=bool(q.bistatic))
            # the block's rows sit at lines line0.. of the full grid: azimuth offsets are relative to the absolute line
            g["azoff"] = np.where(g["azoff"] == -999999.0, -999999.0, g["azoff"] - p.line0)
            for k in ("azt", "rgm", "azoff", "rgoff"):
                if jb["out"].get(k) is not None:
                    jb["out"][k][...] = g[k].astype(jb["out"][k].dtype)
            geos.append(dict(num_outside=g["num_outside"], num_valid=g["num_valid"], num_converged=g["num_conv"],
                             iterations=g["total_iters"], ms_setup=0.0, ms_kernels=0.0, ms_total=0.0, gpu_launches=2))
        return tres, geos
    return topo_geo2rdr_run


@pytest.mark.parametrize("devices", [[0], [0, 1, 2]])
def test_chained_components_host_logic(tmp_path, monkeypatch, devices):
    calls = []
    monkeypatch.setattr(_capi, "topo_geo2rdr_run", _oracle_as_library(calls))
    sc = pu.rough_scene(24, 512)
    dem = IF.createDemImage()
    sc.dem.tofile(tmp_path / "dem.dem")
    dem.initImage(str(tmp_path / "dem.dem"), "read", sc.dem.shape[1], "FLOAT")
    dem.setLength(sc.dem.shape[0])
    dem.firstLatitude, dem.firstLongitude, dem.deltaLatitude, dem.deltaLongitude = sc.first_lat, sc.first_lon, sc.delta_lat, sc.delta_lon
    day = sc.sensing_start.replace(hour=0, minute=0, second=0, microsecond=0)
    topo = isce2_b200.createTopozero()
    topo.slantRangePixelSpacing, topo.prf, topo.radarWavelength = sc.dr, sc.prf, sc.wvl
    topo.orbit = Orbit.from_arrays(day, sc.orbit_t, sc.orbit_pos, sc.orbit_vel)
    topo.width, topo.length = sc.width, sc.length
    topo.wireInputPort(name="dem", object=dem)
    topo.wireInputPort(name="planet", object=Planet(pname="Earth"))
    topo.lookSide, topo.sensingStart, topo.rangeFirstSample = sc.side, sc.sensing_start, sc.r0
    topo.numberRangeLooks = topo.numberAzimuthLooks = 1
    topo.pegHeading = sc.peg_heading  # the component's own default (from the orbit) differs from the scene's in the last digits
    g = tmp_path / "geom"
    topo.latFilename, topo.lonFilename, topo.heightFilename, topo.losFilename = (str(g / f) for f in ("lat.rdr", "lon.rdr", "z.rdr", "los.rdr"))
    topo.gpuDevices = devices
    sec = synth.config_c1_secondary(length=sc.length, width=sc.width)
    kw = pu.secondary_kwargs(sc, sec, recenter=0.37)
    grdr = isce2_b200.createGeo2rdr()
    grdr.configure()
    grdr.slantRangePixelSpacing, grdr.prf, grdr.radarWavelength = sc.dr, sc.prf, sc.wvl
    grdr.orbit = Orbit.from_arrays(day, sec.orbit_t, sec.orbit_pos, sec.orbit_vel)
    grdr.width, grdr.length = sc.width, sc.length
    grdr.wireInputPort(name="planet", object=Planet(pname="Earth"))
    grdr.lookSide = sc.side
    grdr.setSensingStart(day + datetime.timedelta(seconds=kw["t0"]))
    grdr.rangeFirstSample = kw["r0"]
    grdr.numberRangeLooks = grdr.numberAzimuthLooks = 1
    grdr.dopplerCentroidCoeffs = [0.]
    grdr.rangeOffsetImageName, grdr.azimuthOffsetImageName = str(g / "range.off"), str(g / "azimuth.off")
    assert topo.chainGeo2rdr(grdr) is grdr
    topo.topo()

    # one library call per device, contiguous blocks that tile the lines
    assert [c[2] for c in calls] == devices or sorted(c[2] for c in calls) == sorted(devices)
    blocks = sorted((c[0], c[1]) for c in calls)
    assert blocks[0][0] == 0 and sum(n for _, n in blocks) == sc.length
    assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(len(blocks) - 1))
    # files == the oracle on the whole grid
    o = orc.topo(**orc.scene_topo_kwargs(sc, dem_method="BILINEAR", want_inc=False, want_mask=False))
    lat = np.fromfile(g / "lat.rdr").reshape(sc.length, sc.width)
    assert np.array_equal(lat, o["lat"]) and np.array_equal(np.fromfile(g / "z.rdr").reshape(sc.length, sc.width), o["hgt"])
    gg = orc.geo2rdr(lat=o["lat"], lon=o["lon"], hgt=o["hgt"], **kw)
    az = np.fromfile(g / "azimuth.off", np.float32).reshape(sc.length, sc.width)
    rg = np.fromfile(g / "range.off", np.float32).reshape(sc.length, sc.width)
    assert np.array_equal(rg, gg["rgoff"].astype(np.float32))
    assert np.array_equal(az == -999999.0, gg["azoff"] == -999999.0)
    v = az != -999999.0
    assert v.any() and np.abs(az[v] - gg["azoff"][v]).max() < 1e-3
    for f in ("lat.rdr", "lon.rdr", "z.rdr", "los.rdr", "range.off", "azimuth.off"):
        assert os.path.exists(g / (f + ".xml")) and os.path.exists(g / (f + ".vrt")), f
    hdr = IF.createImage().load(str(g / "azimuth.off.xml"))
    assert (hdr.dataType, hdr.width, hdr.length) == ("FLOAT", sc.width, sc.length)
    assert grdr.numValid == gg["num_valid"] and grdr.numOutsideImage == gg["num_outside"]
    assert topo.totalConverged == o["totalconv"] and abs(topo.minimumLatitude - o["min_lat"]) < 1e-12


def test_bench_component_runners_host_logic(tmp_path, monkeypatch):
    """The two runners bench.py times for `e2e_component` -- topo() with geo2rdr chained, and topo() followed by a separate
    geo2rdr() on the rasters it wrote -- produce the same offset rasters (library calls replaced by the oracle), and the
    separate geo2rdr() really reads lat / lon / hgt back from the files, declared as the file mappings they are."""
    from isce2_b200 import synth_components as comp
    calls, declared = [], []
    fused = _oracle_as_library(calls)
    monkeypatch.setattr(_capi, "topo_geo2rdr_run", fused)

    def topo_run(params, dem, t, pos, vel, dop, slr=None, rho_image=None, want_los=True, want_inc=False, want_mask=False,
                 out=None, doppler_poly=None, slrng_poly=None):
        return fused(params, dem, t, pos, vel, dop, [], slr, rho_image, want_los, want_inc, want_mask, out, doppler_poly,
                     slrng_poly)[0]

    def geo2rdr_run(params, lat, lon, hgt, t, pos, vel, coeffs=(0.0,), mean=0.0, norm=1.0, want=None, out=None, block_rows=False):
        q = params
        for a in (lat, lon, hgt):  # what the component hands over are views of the rasters' mappings
            assert IF._file_range(a, False) is not None
        g = orc.geo2rdr(lat=np.asarray(lat), lon=np.asarray(lon), hgt=np.asarray(hgt), orbit_t=t, orbit_pos=pos, orbit_vel=vel,
                        length=q.length, width=q.width, r0=q.rho0, dr=q.drho, prf=q.prf, t0=q.t0, wvl=q.wvl, side=q.look_side,
                        doppler_coeffs=coeffs, doppler_mean=mean, doppler_norm=norm, a=q.major, e2=q.e2,
                        orbit_method=ORB_NAMES[q.orbit_method], bistatic=bool(q.bistatic))
        for k in ("azt", "rgm", "azoff", "rgoff"):
            if out.get(k) is not None:
                out[k][...] = g[k][q.line0:q.line0 + q.nlines].astype(out[k].dtype)
        return dict(num_outside=g["num_outside"], num_valid=g["num_valid"], num_converged=g["num_conv"], iterations=g["total_iters"],
                    ms_setup=0.0, ms_kernels=0.0, ms_total=0.0, gpu_launches=2)

    monkeypatch.setattr(_capi, "topo_run", topo_run)
    monkeypatch.setattr(_capi, "geo2rdr_run", geo2rdr_run)
    real_register = _capi.host_file_register
    monkeypatch.setattr(_capi, "host_file_register", lambda addr, n, fd, off: (declared.append(n), real_register(addr, n, fd, off))[1])
    sc = synth.make_scene(20, 384)
    sec = synth.make_scene(20, 384, dem=False, perturb=dict(da=120.0, d_cross=80.0, d_along_s=0.0))
    dem_img = comp.prepare_dem(sc, str(tmp_path / "dem.dem"))
    a = comp.run_components(sc, sec, dem_img, str(tmp_path / "chained"), dem_method="BILINEAR", inc=True, mask=True, devices=[0])
    n_chained = len(declared)
    b = comp.run_components_separately(sc, sec, dem_img, str(tmp_path / "separate"), dem_method="BILINEAR", inc=True, mask=True,
                                       devices=[0])
    assert a["files"] == b["files"] and a["bytes_written"] == b["bytes_written"] == sc.pixels * 49
    assert b["seconds_topo"] > 0 and b["seconds_geo2rdr"] > 0 and a["num_valid"] == b["num_valid"] > 0.5 * sc.pixels
    for f in a["files"]:
        x, y = np.fromfile(tmp_path / "chained" / f, np.uint8), np.fromfile(tmp_path / "separate" / f, np.uint8)
        assert np.array_equal(x, y), f
    # chained: 8 rasters declared for writing; separate: 6 by topo(), then 2 + the 3 it reads by geo2rdr()
    assert n_chained == 8 and len(declared) - n_chained == 11
