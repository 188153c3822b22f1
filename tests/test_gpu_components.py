"""End-to-end through the drop-in Component classes on the GPU: createTopozero().topo() writes the .rdr layers with
their XML/VRT, createGeo2rdr().geo2rdr() reads them back and writes the .off rasters -- the call sequence of
components/isceobj/StripmapProc/runTopo.py:66-103 and runGeo2rdr.py:57-110 -- checked against the CPU oracle."""
import datetime
import os

import numpy as np
import pytest

import isce2_b200
from isce2_b200 import image as IF, synth
from isce2_b200.orbit import Orbit
from isce2_b200.planet import Planet
from isce2_b200.poly import Poly2D
from oracle import oracle as orc
from tests import parity_util as pu

pytestmark = pytest.mark.gpu


def _write_dem(sc, path, as_int16=False):
    dem = IF.createDemImage()
    arr = np.round(sc.dem).astype(np.int16) if as_int16 else sc.dem
    arr.tofile(path)
    dem.initImage(path, "read", sc.dem.shape[1], "SHORT" if as_int16 else "FLOAT")
    dem.setLength(sc.dem.shape[0])
    dem.firstLatitude, dem.firstLongitude = sc.first_lat, sc.first_lon
    dem.deltaLatitude, dem.deltaLongitude = sc.delta_lat, sc.delta_lon
    dem.renderHdr()
    return IF.createDemImage().load(path + ".xml"), arr.astype(np.float32)


def _orbit(sc):
    day = sc.sensing_start.replace(hour=0, minute=0, second=0, microsecond=0)
    return Orbit.from_arrays(day, sc.orbit_t, sc.orbit_pos, sc.orbit_vel)


def test_topo_then_geo2rdr_through_components(tmp_path):
    sc = pu.rough_scene(40, 3000)
    dem, demf = _write_dem(sc, str(tmp_path / "dem.dem"), as_int16=True)
    geom = tmp_path / "geometry"
    topo = isce2_b200.createTopozero()
    topo.slantRangePixelSpacing = sc.dr
    topo.prf = sc.prf
    topo.radarWavelength = sc.wvl
    topo.orbit = _orbit(sc)
    topo.width, topo.length = sc.width, sc.length
    topo.wireInputPort(name="dem", object=dem)
    topo.wireInputPort(name="planet", object=Planet(pname="Earth"))
    topo.numberRangeLooks = 1
    topo.numberAzimuthLooks = 1
    topo.lookSide = sc.side
    topo.sensingStart = sc.sensing_start
    topo.rangeFirstSample = sc.r0
    topo.demInterpolationMethod = "BIQUINTIC"
    topo.latFilename = str(geom / "lat.rdr")
    topo.lonFilename = str(geom / "lon.rdr")
    topo.heightFilename = str(geom / "z.rdr")
    topo.losFilename = str(geom / "los.rdr")
    topo.incFilename = str(geom / "incLocal.rdr")
    topo.maskFilename = str(geom / "shadowMask.rdr")
    topo.topo()

    sc.dem = demf
    c = pu.cpu_topo(sc, dem_method="BIQUINTIC")
    for f in ("lat.rdr", "lon.rdr", "z.rdr", "los.rdr", "incLocal.rdr", "shadowMask.rdr"):
        assert os.path.exists(geom / (f + ".xml")) and os.path.exists(geom / (f + ".vrt"))
    lat = IF.createImage().load(str(geom / "lat.rdr.xml"))
    assert (lat.width, lat.length, lat.dataType, lat.bands) == (sc.width, sc.length, "DOUBLE", 1)
    los = IF.createImage().load(str(geom / "los.rdr.xml"))
    assert (los.bands, los.scheme, los.dataType, los.imageType) == (2, "BIL", "FLOAT", "bil")
    g = dict(lat=np.asarray(lat.memMap()), lon=np.fromfile(geom / "lon.rdr").reshape(sc.length, sc.width),
             hgt=np.fromfile(geom / "z.rdr").reshape(sc.length, sc.width), los=np.asarray(los.memMap()),
             inc=np.fromfile(geom / "incLocal.rdr", np.float32).reshape(sc.length, 2, sc.width),
             mask=np.fromfile(geom / "shadowMask.rdr", np.int8).reshape(sc.length, sc.width))
    assert np.abs(g["lat"] - c["lat"]).max() < 1e-7 and (np.abs(g["lat"] - c["lat"]) > pu.TOL_LATLON_DEG).sum() <= 2
    assert np.abs(g["lon"] - c["lon"]).max() < 1e-7 and (np.abs(g["lon"] - c["lon"]) > pu.TOL_LATLON_DEG).sum() <= 2
    assert np.abs(g["hgt"] - c["hgt"]).max() < pu.TOL_HGT_M
    assert np.array_equal(g["mask"], c["mask"])
    assert (np.abs(g["los"].astype(np.float64) - c["los"]) > pu.TOL_ANGLE_DEG).sum() <= 4
    assert abs(topo.minimumLatitude - c["min_lat"]) < 1e-9 and abs(topo.maximumLongitude - c["max_lon"]) < 1e-9
    assert topo.snwe[0] < topo.snwe[1]

    # ---- geo2rdr from the files topo wrote, secondary orbit, DOUBLE outputs as StripmapProc/runGeo2rdr.py:108 ----
    sec = synth.config_c1_secondary(length=sc.length, width=sc.width)
    kw = pu.secondary_kwargs(sc, sec, recenter=0.37)
    grdr = isce2_b200.createGeo2rdr()
    grdr.configure()
    grdr.slantRangePixelSpacing = sc.dr
    grdr.prf = sc.prf
    grdr.radarWavelength = sc.wvl
    day = sc.sensing_start.replace(hour=0, minute=0, second=0, microsecond=0)
    grdr.orbit = Orbit.from_arrays(day, sec.orbit_t, sec.orbit_pos, sec.orbit_vel)
    grdr.width, grdr.length = sc.width, sc.length
    grdr.wireInputPort(name="planet", object=Planet(pname="Earth"))
    grdr.lookSide = sc.side
    grdr.setSensingStart(day + datetime.timedelta(seconds=kw["t0"]))
    grdr.rangeFirstSample = kw["r0"]
    grdr.numberRangeLooks = 1
    grdr.numberAzimuthLooks = 1
    grdr.dopplerCentroidCoeffs = [0.]
    grdr.fmrateCoeffs = [0.]
    off = tmp_path / "offsets"
    grdr.rangeOffsetImageName = str(off / "range.off")
    grdr.azimuthOffsetImageName = str(off / "azimuth.off")
    for attr, f in (("demImage", "z.rdr"), ("latImage", "lat.rdr"), ("lonImage", "lon.rdr")):
        img = IF.createImage()
        img.load(str(geom / (f + ".xml")))
        img.setAccessMode("READ")
        setattr(grdr, attr, img)
    grdr.outputPrecision = "DOUBLE"
    grdr.geo2rdr()
    o = orc.geo2rdr(lat=g["lat"], lon=g["lon"], hgt=g["hgt"], **kw)
    rg = np.fromfile(off / "range.off").reshape(sc.length, sc.width)
    az = np.fromfile(off / "azimuth.off").reshape(sc.length, sc.width)
    assert np.array_equal(rg == -999999.0, o["rgoff"] == -999999.0)
    v = rg != -999999.0
    assert v.mean() > 0.5
    assert np.abs(rg[v] - o["rgoff"][v]).max() < pu.TOL_OFFSET_PX and np.abs(az[v] - o["azoff"][v]).max() < pu.TOL_OFFSET_PX
    hdr = IF.createImage().load(str(off / "range.off.xml"))
    assert (hdr.dataType, hdr.width, hdr.length) == ("DOUBLE", sc.width, sc.length)

    # 'single' precision: FLOAT rasters (the reference narrows through a DoubleToFloat caster)
    grdr2 = isce2_b200.createGeo2rdr()
    grdr2.configure()
    for a in ("slantRangePixelSpacing prf radarWavelength orbit width length lookSide sensingStart rangeFirstSample "
              "numberRangeLooks numberAzimuthLooks dopplerCentroidCoeffs demImage latImage lonImage").split():
        setattr(grdr2, a, getattr(grdr, a))
    grdr2.rangeOffsetImageName = str(off / "range_f.off")
    grdr2.geo2rdr()
    rgf = np.fromfile(off / "range_f.off", np.float32).reshape(sc.length, sc.width)
    assert np.array_equal(rgf, rg.astype(np.float32))


def test_topo_component_native_doppler_and_line_sharding_attribute(tmp_path):
    """polyDoppler given by the caller (StripmapProc/runTopo.py:91-101) and gpuDevices=[0, 0]: two line blocks on the
    same device must tile the output files exactly like a single block."""
    sc = synth.make_scene(24, 2048, sensor="nisar")
    dem, _ = _write_dem(sc, str(tmp_path / "dem.dem"))
    outs = {}
    for tag, devs in (("one", [0]), ("two", [0, 0])):
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
        dop = Poly2D()
        dop.setWidth(sc.width); dop.setLength(sc.length)
        dop.initPoly(rangeOrder=len(sc.doppler_coeffs[0]) - 1, azimuthOrder=0, coeffs=sc.doppler_coeffs)
        topo.polyDoppler = dop
        topo.orbitInterpolationMethod = "LEGENDRE"
        topo.demInterpolationMethod = "BIQUINTIC"
        d = tmp_path / tag
        topo.latFilename, topo.lonFilename, topo.heightFilename, topo.losFilename = (str(d / f) for f in
                                                                                      ("lat.rdr", "lon.rdr", "z.rdr", "los.rdr"))
        topo.maskFilename = str(d / "mask.rdr")
        topo.gpuDevices = devs
        topo.topo()
        outs[tag] = {f: np.fromfile(d / f, np.uint8) for f in ("lat.rdr", "lon.rdr", "z.rdr", "los.rdr", "mask.rdr")}
        outs[tag]["snwe"] = topo.snwe
    for f in ("lat.rdr", "lon.rdr", "z.rdr", "los.rdr", "mask.rdr"):
        assert np.array_equal(outs["one"][f], outs["two"][f]), f
    assert outs["one"]["snwe"] == outs["two"]["snwe"]
    c = pu.cpu_topo(sc, dem_method="BIQUINTIC", orbit_method="LEGENDRE", want_inc=False)
    lat = outs["one"]["lat.rdr"].view(np.float64).reshape(sc.length, sc.width)
    assert np.abs(lat - c["lat"]).max() < 1e-7 and (np.abs(lat - c["lat"]) > pu.TOL_LATLON_DEG).sum() <= 2
    assert np.array_equal(outs["one"]["mask.rdr"].view(np.int8).reshape(sc.length, sc.width), c["mask"])
