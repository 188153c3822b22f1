"""Host-side logic of the drop-in Component classes (no GPU needed): attribute surface, defaults and validation
messages of Topozero.py / Geo2rdr.py, ports, ISCE XML + VRT metadata round trip, zerodop aliasing."""
import datetime
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

import isce2_b200
from isce2_b200 import image as IF, synth
from isce2_b200.geo2rdr import Geo2rdr
from isce2_b200.orbit import Orbit, enu_heading_deg, export_rows
from isce2_b200.planet import Planet
from isce2_b200.poly import Poly1D, Poly2D
from isce2_b200.topozero import Topo


def test_factories_and_alias():
    t = isce2_b200.createTopozero()
    g = isce2_b200.createGeo2rdr()
    assert isinstance(t, Topo) and isinstance(g, Geo2rdr)
    isce2_b200.install_as_zerodop()
    from zerodop.geo2rdr import createGeo2rdr
    from zerodop.topozero import createTopozero
    assert isinstance(createTopozero(), Topo) and isinstance(createGeo2rdr(), Geo2rdr)


def test_topo_attribute_surface_matches_reference():
    # components/zerodop/topozero/Topozero.py:613-670
    t = Topo()
    for a in ("numberIterations secondaryIterations threshold demWidth demLength orbit sensingStart firstLatitude "
              "firstLongitude deltaLatitude deltaLongitude ellipsoidMajorSemiAxis ellipsoidEccentricitySquared length width "
              "slantRangePixelSpacing rangeFirstSample numberRangeLooks numberAzimuthLooks pegHeading prf radarWavelength "
              "demFilename latFilename lonFilename heightFilename losFilename incFilename maskFilename slantRangeFilename "
              "demImage latImage lonImage heightImage losImage incImage maskImage slantRangeImage minimumLatitude "
              "minimumLongitude maximumLatitude maximumLongitude lookSide polyDoppler demInterpolationMethod "
              "orbitInterpolationMethod").split():
        assert hasattr(t, a), a
    assert set(t.listInputPorts()) == {"frame", "planet", "dem", "interferogram"} if hasattr(t, "listInputPorts") else True
    assert t.interpolationMethods["BIQUINTIC"] == 5 and t.orbitInterpolationMethods["LEGENDRE"] == 2
    assert "MINIMUM_LATITUDE" in t.dictionaryOfOutputVariables and "PEG_HEADING" in t.dictionaryOfVariables
    t.snwe = (1.0, 2.0, 3.0, 4.0)
    assert t.getMinimumLatitude() == 1.0 and t.getMaximumLongitude() == 4.0
    for s in ("setNumberIterations setDemWidth setOrbit setFirstLatitude setPegHeading setPRF setRadarWavelength "
              "setLosFilename setIncidenceFilename setMaskFilename setLookSide setPolyDoppler").split():
        assert callable(getattr(t, s)), s


def test_geo2rdr_parameter_surface_matches_reference():
    # components/zerodop/geo2rdr/Geo2rdr.py:45-185,506-544
    g = Geo2rdr()
    g.configure()
    assert g.outputPrecision == "single" and g.ellipsoidMajorSemiAxis == 6378137.0
    for a in ("slantRangePixelSpacing rangeFirstSample prf radarWavelength sensingStart numberRangeLooks numberAzimuthLooks "
              "lookSide bistaticDelayCorrectionFlag orbitInterpolationMethod rangeImageName azimuthImageName "
              "rangeOffsetImageName azimuthOffsetImageName latImage lonImage demImage dopplerCentroidCoeffs orbit width "
              "length demWidth demLength polyDoppler").split():
        assert hasattr(g, a), a
    assert [p.public_name for p in g.parameter_list][:4] == ["RANGE_FILENAME", "AZIMUTH_FILENAME", "RANGE_OFFSET_FILENAME",
                                                              "AZIMUTH_OFFSET_FILENAME"]
    with pytest.raises(Exception, match="No orbit provided for geocoding"):
        g.geo2rdr()
    with pytest.raises(KeyError):
        g.wireInputPort(name="nonexistent", object=None)


def test_ports_fill_attributes():
    class Instr:
        def getRangePixelSize(self): return 2.33
        def getPulseRepetitionFrequency(self): return 486.0
        def getRadarWavelength(self): return 0.0555

    class Frame:
        def getInstrument(self): return Instr()
        def getOrbit(self): return "ORB"

    dem = IF.createDemImage()
    dem.initImage("x.dem", "read", 100)
    dem.setLength(50)
    dem.firstLatitude, dem.firstLongitude, dem.deltaLatitude, dem.deltaLongitude = 35.0, -118.0, -1 / 3600, 1 / 3600
    t = Topo()
    t.wireInputPort(name="frame", object=Frame())
    t.wireInputPort(name="planet", object=Planet(pname="Earth"))
    t.wireInputPort(name="dem", object=dem)
    for port in t._inputPorts:
        port()
    assert (t.slantRangePixelSpacing, t.prf, t.radarWavelength, t.orbit) == (2.33, 486.0, 0.0555, "ORB")
    assert t.ellipsoidMajorSemiAxis == 6378137.0 and t.demWidth == 100 and t.demLength == 50 and t.firstLatitude == 35.0


def test_topo_defaults_and_validation_messages():
    sc = synth.make_scene(8, 64, dem_spacing_arcsec=3.0)
    t = Topo()
    t.width, t.length, t.prf = sc.width, sc.length, sc.prf
    t.sensingStart = sc.sensing_start
    t.orbit = Orbit.from_arrays(sc.sensing_start.replace(hour=0, minute=0, second=0, microsecond=0), sc.orbit_t, sc.orbit_pos,
                                sc.orbit_vel)
    with pytest.raises(Exception, match="slantRangePixelSpacing cannot be None"):
        t.setDefaults()
    t.slantRangePixelSpacing, t.rangeFirstSample = sc.dr, sc.r0
    t.demInterpolationMethod = "WRONG"
    with pytest.raises(Exception, match="Interpolation method must be one of"):
        t.setDefaults()
    t.demInterpolationMethod = None
    t.setDefaults()
    assert (t.numberIterations, t.secondaryIterations, t.threshold) == (25, 10, 0.05)
    assert t.demInterpolationMethod == "BILINEAR" and t.orbitInterpolationMethod == "HERMITE"
    assert (t.latFilename, t.lonFilename, t.heightFilename, t.losFilename) == ("lat.rdr", "lon.rdr", "z.rdr", "los.rdr")
    assert t.polyDoppler.getCoeffs() == [[0.0]]
    # default peg heading = ENU heading at mid-scene (Topozero.py:162-165); synth computes the same quantity
    assert abs(t.pegHeading - sc.peg_heading) < 1e-9
    bad = Poly2D()
    bad.initPoly(rangeOrder=0, azimuthOrder=0, coeffs=[[0.0]])
    bad.setWidth(sc.width + 1)
    bad.setLength(sc.length)
    t.polyDoppler = bad
    with pytest.raises(Exception, match="same width"):
        t.setDefaults()


def test_orbit_export_matches_reference_convention():
    day = datetime.datetime(2026, 10, 17)
    t = np.array([21590.0, 21600.0, 21610.0, 21620.0, 21630.0])
    sc = synth.make_scene(4, 16, dem=False)
    o = Orbit.from_arrays(day, sc.orbit_t, sc.orbit_pos, sc.orbit_vel)
    tt, pos, vel = export_rows(o, day + datetime.timedelta(seconds=21600.5))  # reference = sensingStart -> its midnight
    assert np.array_equal(tt, sc.orbit_t) and np.array_equal(pos, sc.orbit_pos) and np.array_equal(vel, sc.orbit_vel)
    assert abs(np.radians(enu_heading_deg(o, day + datetime.timedelta(seconds=sc.t0 + 0.5 * 4 / sc.prf))) - sc.peg_heading) < 1e-9


def test_image_xml_vrt_round_trip(tmp_path):
    fn = str(tmp_path / "los.rdr")
    img = IF.createImage()
    img.initImage(fn, "write", 7, "FLOAT", bands=2, scheme="BIL")
    img.setLength(3)
    mm = img.createImage()
    assert mm.shape == (3, 2, 7)
    mm[:] = np.arange(42, dtype=np.float32).reshape(3, 2, 7)
    img.setImageType("bil")
    img.addDescription("test layer")
    img.finalizeImage()
    img.renderHdr()
    # raw layout: per line [band0 x width][band1 x width] (BILAccessor.cpp:11-37)
    raw = np.fromfile(fn, np.float32)
    assert np.array_equal(raw, np.arange(42, dtype=np.float32))
    root = ET.parse(fn + ".xml").getroot()
    assert root.tag == "imageFile"
    props = {e.get("name"): e.find("value").text for e in root.findall("property")}
    assert props["width"] == "7" and props["length"] == "3" and props["number_bands"] == "2" and props["scheme"] == "BIL"
    assert props["data_type"] == "FLOAT" and props["byte_order"] == "l" and props["image_type"] == "bil"
    comps = {e.get("name"): e for e in root.findall("component")}
    assert set(comps) == {"coordinate1", "coordinate2"}
    assert comps["coordinate1"].find("factoryname").text == "createCoordinate"
    vrt = ET.parse(fn + ".vrt").getroot()
    assert vrt.get("rasterXSize") == "7" and vrt.get("rasterYSize") == "3"
    bands = vrt.findall("VRTRasterBand")
    assert [b.get("dataType") for b in bands] == ["Float32", "Float32"] and bands[0].get("subClass") == "VRTRawRasterBand"
    # Image.py:563-566
    assert [b.find("ImageOffset").text for b in bands] == ["0", str(7 * 4)]
    assert bands[1].find("PixelOffset").text == "4" and bands[1].find("LineOffset").text == str(2 * 7 * 4)
    back = IF.createImage().load(fn + ".xml")
    assert (back.width, back.length, back.bands, back.dataType, back.scheme) == (7, 3, 2, "FLOAT", "BIL")
    assert np.array_equal(np.asarray(back.memMap()), np.arange(42, dtype=np.float32).reshape(3, 2, 7))


def test_dem_xml_old_style_uppercase(tmp_path):
    """Old upper-case ISCE headers (e.g. components/isceobj/Util/test/resampImage.int.xml) load too."""
    fn = str(tmp_path / "dem.dem")
    np.arange(12, dtype=np.int16).tofile(fn)
    xml = """<imageFile><property name="WIDTH"><value>4</value></property><property name="LENGTH"><value>3</value></property>
    <property name="DATA_TYPE"><value>SHORT</value></property><property name="NUMBER_BANDS"><value>1</value></property>
    <property name="SCHEME"><value>BIP</value></property><property name="FILE_NAME"><value>dem.dem</value></property>
    <component name="Coordinate1"><property name="startingValue"><value>-118.0</value></property>
    <property name="delta"><value>0.000277777777778</value></property><property name="size"><value>4</value></property></component>
    <component name="Coordinate2"><property name="startingValue"><value>36.0</value></property>
    <property name="delta"><value>-0.000277777777778</value></property><property name="size"><value>3</value></property></component>
    </imageFile>"""
    open(fn + ".xml", "w").write(xml)
    dem = IF.createDemImage().load(fn + ".xml")
    assert (dem.getWidth(), dem.getLength(), dem.dataType) == (4, 3, "SHORT")
    assert dem.getFirstLatitude() == 36.0 and dem.getFirstLongitude() == -118.0 and dem.getDeltaLatitude() < 0
    assert IF.read_raster(dem).shape == (3, 4)


def test_poly_evaluation_order_matches_reference_c():
    from oracle import oracle as orc
    p = Poly2D()
    p.initPoly(rangeOrder=2, azimuthOrder=1, coeffs=[[1.5, -2.0e-3, 3e-8], [0.25, 1e-5, -2e-9]])
    p.setMeanRange(100.0); p.setNormRange(50.0); p.setMeanAzimuth(10.0); p.setNormAzimuth(4.0)
    ref = orc.Poly2D([[1.5, -2.0e-3, 3e-8], [0.25, 1e-5, -2e-9]], 100.0, 10.0, 50.0, 4.0)
    for az, rg in ((0, 0), (3, 777), (1499, 24999)):
        assert p(az, rg) == ref(az, rg)
    q = Poly1D()
    q.initPoly(order=2, coeffs=[0.1, 0.02, -3e-4]); q.setMean(5.0); q.setNorm(2.0)
    r1 = orc.Poly1D([0.1, 0.02, -3e-4], 5.0, 2.0)
    assert q(123.0) == r1(123.0)


def test_install_as_zerodop_resolves_both_import_forms():
    """zerodop.<name> is a package in the reference; the class modules must be importable too."""
    import importlib
    import sys
    import isce2_b200
    saved = {k: v for k, v in sys.modules.items() if k == "zerodop" or k.startswith("zerodop.")}
    for k in saved:
        del sys.modules[k]
    try:
        isce2_b200.install_as_zerodop()
        from zerodop.topozero import createTopozero
        from zerodop.geo2rdr import createGeo2rdr
        from zerodop.topozero.Topozero import Topo
        from zerodop.geo2rdr.Geo2rdr import Geo2rdr
        from zerodop.geozero.Geozero import Geocode
        assert isinstance(createTopozero(), Topo) and isinstance(createGeo2rdr(), Geo2rdr) and Geocode is not None
        assert importlib.import_module("zerodop.geo2rdr.Geo2rdr").Geo2rdr is Geo2rdr
    finally:
        for k in [k for k in sys.modules if k == "zerodop" or k.startswith("zerodop.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_image_byte_order_long_vrt_and_foreign_images(tmp_path):
    """big-endian rasters are read as such, LONG renders a VRT, foreign (isceobj-like) images are honoured
    in their own interleaving on input and mapped writable on output."""
    import numpy as np
    from isce2_b200 import image as IF
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    p = str(tmp_path / "be.bin")
    a.astype(">f4").tofile(p)
    img = IF.createImage()
    img.initImage(p, "read", 4, "FLOAT")
    img.setLength(3)
    img.byteOrder = "b"
    got = IF.read_raster(img)
    assert got.dtype == np.float32 and got.dtype.isnative and np.array_equal(got, a)
    lg = IF.createImage()
    q = str(tmp_path / "long.bin")
    np.arange(6, dtype=np.int64).tofile(q)
    lg.initImage(q, "read", 3, "LONG")
    lg.setLength(2)
    lg.renderHdr()
    assert "Int64" in open(q + ".vrt").read()

    class Foreign:  # what an isceobj Image exposes
        def __init__(self, fn, width, length, bands, scheme, dataType):
            self.filename, self.width, self.length, self.bands, self.scheme, self.dataType = fn, width, length, bands, scheme, dataType
            self.byteOrder = "l"

        def getFilename(self):
            return self.filename

        def memMap(self, mode="r", band=None):  # read-only, (length, 1, width) for one band: unusable as an output
            raise AssertionError("the foreign image's own memMap must not be used")

    bip = np.arange(2 * 3 * 2, dtype=np.float32).reshape(2, 3, 2)  # [line][sample][band]
    r = str(tmp_path / "bip.bin")
    bip.tofile(r)
    assert np.array_equal(IF.read_raster(Foreign(r, 3, 2, 2, "BIP", "FLOAT")), bip)
    o = str(tmp_path / "out" / "lat.rdr")
    mm = IF.output_memmap(Foreign(o, 5, None, 1, "BIL", "DOUBLE"), 4, 5)
    assert mm.shape == (4, 5) and mm.dtype == np.float64 and mm.flags.writeable
    mm[:] = 7.0
    mm.flush()
    assert np.fromfile(o).sum() == 7.0 * 20
    import pytest
    with pytest.raises(ValueError):
        IF.output_memmap(Foreign(o, 6, None, 1, "BIL", "DOUBLE"), 4, 5)
