"""geozero on the GPU against the CPU oracle: the solved image coordinates, the cropped DEM, the
counters and the geocoded samples for the four interpolators, real and complex, single- and multi-band; then the
drop-in ``createGeozero()`` Component end to end with files.

Tolerances (written here because BASELINE's north_star names none for geozero): image coordinates within 1e-6 pixel
(the GPU evaluates the Hermite orbit through per-window polynomials; measured ~1e-9), cropped DEM and validity
bit-exact; geocoded samples: nearest-neighbour bit-exact, interpolating methods within 1e-6 of the image's dynamic
range (a float32 ulp or two, >90 % of the samples bit-identical), except where a coordinate difference of that size
flips an integer index / sinc phase bin / the reference's |dt| < 5e-7 s stopping test: at most 1e-4 of the pixels,
and then by less than 2 % of the dynamic range.  For scale: the reference's own stopping tolerance leaves
2e-4 pixel of azimuth error in every coordinate."""
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
from tests.test_geozero_oracle_cpu import _grid_kw, _inner_box
from tests.test_gpu_components import _write_dem

pytestmark = pytest.mark.gpu


def _textured(sc, seed, complex_=False):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:sc.length, 0:sc.width]
    a = (np.sin(xx / 37.0) * np.cos(yy / 11.0) * 50.0 + rng.normal(scale=3.0, size=xx.shape)).astype(np.float32)
    if not complex_:
        return a
    b = (np.cos(xx / 23.0 + yy / 17.0) * 40.0 + rng.normal(scale=3.0, size=xx.shape)).astype(np.float32)
    return (a + 1j * b).astype(np.complex64)


def _gparams(sc, snwe, dem):
    return _capi.geozero_params(dem_shape=dem.shape, side=sc.side, length=sc.length, width=sc.width, **_grid_kw(sc, snwe))


def _compare(gpu, cpu, image, method):
    assert gpu.shape == cpu.shape and gpu.dtype == cpu.dtype
    d = np.abs(gpu.astype(np.complex128) - cpu.astype(np.complex128))
    span = float(np.abs(image).max()) * 2.0
    # validity (zero fill) must agree except where a coordinate sits within 1e-6 px of the image-edge tests
    assert int(((gpu == 0) != (cpu == 0)).sum()) <= 2
    if method == "NEAREST":  # pure index work: bit-exact, up to a rounding flip of nint() on a handful of pixels
        assert int((d != 0).sum()) <= max(2, int(1e-4 * d.size)), (method, int((d != 0).sum()), d.size)
        return
    # interpolating methods: coordinates agree to ~1e-9 px, which moves a float32 sample by an ulp here and there
    nloose = int((d > 1e-6 * span).sum())
    assert nloose <= max(2, int(1e-4 * d.size)), (method, nloose, d.size)
    assert float(d.max()) <= 0.02 * span, (method, float(d.max()))
    assert float((d == 0).mean()) > 0.9, (method, float((d == 0).mean()))


@pytest.mark.parametrize("complex_", [False, True])
def test_geozero_parity_all_methods(complex_):
    sc = pu.rough_scene(400, 3000)
    c = pu.cpu_topo(sc, dem_method="BILINEAR", want_inc=False, want_mask=False)
    snwe = _inner_box(c["lat"], c["lon"], -0.1, -0.1)  # larger than the footprint: image edges and the outside are in
    img = _textured(sc, 11, complex_)
    kw = dict(dem=sc.dem, orbit_t=sc.orbit_t, orbit_pos=sc.orbit_pos, orbit_vel=sc.orbit_vel, side=sc.side, **_grid_kw(sc, snwe))
    plan = _capi.GeozeroPlan(_gparams(sc, snwe, sc.dem), sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel)
    first = True
    for method in ("BILINEAR", "BICUBIC", "SINC", "NEAREST"):
        o = orc.geozero(image=img, method=method, **kw)
        g = plan.geocode(img, method=method)
        r = plan.fetch(want_indices=True)
        assert (r["geo_length"], r["geo_width"]) == (o["geolength"], o["geowidth"]) == g.shape
        if first:
            first = False
            assert np.array_equal(r["dem_crop"], o["dem_crop"])
            # the oracle reports coordinates only for pixels that reached the interpolator; the plan keeps them for every
            # pixel that went through the solve
            v = np.isfinite(o["az_idx"])
            assert v.sum() > 0.2 * v.size and np.isfinite(r["az_idx"][v]).all()
            assert np.abs(r["az_idx"][v] - o["az_idx"][v]).max() < 1e-6
            assert np.abs(r["rng_idx"][v] - o["rng_idx"][v]).max() < 1e-6
            assert abs(r["iterations"] - o["total_iters"]) <= max(4, 1e-4 * o["total_iters"])
            for k_g, k_o in (("geo_min_lat", "geomin_lat"), ("geo_max_lat", "geomax_lat"), ("geo_min_lon", "geomin_lon"),
                             ("geo_max_lon", "geomax_lon")):
                assert r[k_g] == o[k_o]
        assert abs(r["num_valid"] - o["num_valid"]) <= 2 and abs(r["num_outside_image"] - o["num_outside_image"]) <= 2
        assert r["num_outside_dem"] == o["num_outside_dem"] == 0
        assert o["num_valid"] > 0.2 * g.size and o["num_outside_image"] > 0
        _compare(g, o["geo"], img, method)
    plan.close()


def test_geozero_multiband_schemes_native_doppler_and_left_looking():
    sc = synth.make_scene(300, 2000, sensor="nisar")
    c = pu.cpu_topo(sc, dem_method="BILINEAR", orbit_method="HERMITE", want_inc=False, want_mask=False)
    snwe = _inner_box(c["lat"], c["lon"], 0.1, 0.1)
    # geozero's Doppler polynomial is in cycles / PRF versus range pixel (Geozero.py setDefaults, geozero.f90:196-207)
    dop = [d / sc.prf for d in sc.doppler_coeffs[0]]
    b0, b1 = _textured(sc, 1), _textured(sc, 2)
    kw = dict(dem=sc.dem, orbit_t=sc.orbit_t, orbit_pos=sc.orbit_pos, orbit_vel=sc.orbit_vel, side=sc.side,
              doppler_coeffs=dop, **_grid_kw(sc, snwe))
    o0 = orc.geozero(image=b0, method="BILINEAR", **kw)
    o1 = orc.geozero(image=b1, method="BILINEAR", **kw)
    assert o0["num_valid"] > 0.15 * o0["geo"].size  # the footprint is a tilted strip inside its lat/lon box
    p = _gparams(sc, snwe, sc.dem)
    for scheme, stack, axis in (("BIL", np.stack([b0, b1], axis=1), 1), ("BIP", np.stack([b0, b1], axis=2), 2),
                                ("BSQ", np.stack([b0, b1], axis=0), 0)):
        r = _capi.geozero_run(p, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, np.ascontiguousarray(stack), method="BILINEAR",
                              nbands=2, scheme=scheme, doppler_coeffs=dop)
        g0, g1 = np.take(r["geo"], 0, axis=axis), np.take(r["geo"], 1, axis=axis)
        _compare(np.ascontiguousarray(g0), o0["geo"], b0, "BILINEAR")
        _compare(np.ascontiguousarray(g1), o1["geo"], b1, "BILINEAR")
        assert np.array_equal(r["dem_crop"], o0["dem_crop"])
        assert abs(r["num_valid"] - o1["num_valid"]) <= 2
    # the native-Doppler terms are live: the zero-Doppler solution of this squinted pass lies a hundred lines away
    # (fd * wvl * R / (2 v^2) in seconds), which moves the footprint inside the box
    z = orc.geozero(image=b0, method="BILINEAR", **{**kw, "doppler_coeffs": (0.0,)})
    assert abs(z["num_valid"] - o0["num_valid"]) > 0.05 * o0["num_valid"]


def test_geozero_edges_voids_int16_dem_and_errors():
    sc = pu.rough_scene(160, 1200)
    c = pu.cpu_topo(sc, dem_method="BILINEAR", want_inc=False, want_mask=False)
    dem16 = np.round(sc.dem).astype(np.int16)
    box = _inner_box(c["lat"], c["lon"], 0.1, 0.1)
    # reach north and west of the DEM: rows / columns outside it
    snwe = (box[0], sc.first_lat + 7.5 * abs(sc.delta_lat), sc.first_lon - 3.5 * sc.delta_lon, box[3])
    g0 = orc.geozero_grid(orc.geozero_params(dem_shape=dem16.shape, length=sc.length, width=sc.width, **_grid_kw(sc, snwe)))
    assert g0["max_lat_idx"] == -7 and g0["min_lon_idx"] == -3
    i, j = g0["geo_len"] - 40, g0["geo_wid"] - 60
    dem16[g0["max_lat_idx"] + i, g0["min_lon_idx"] + j] = -32768
    img = _textured(sc, 3)
    o = orc.geozero(dem=dem16.astype(np.float32), image=img, orbit_t=sc.orbit_t, orbit_pos=sc.orbit_pos, orbit_vel=sc.orbit_vel,
                    method="NEAREST", side=sc.side, **_grid_kw(sc, snwe))
    p = _gparams(sc, snwe, dem16)
    r = _capi.geozero_run(p, dem16, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, img, method="NEAREST")
    assert r["num_outside_dem"] == o["num_outside_dem"] == 7 * dem16.shape[1]
    assert np.array_equal(r["dem_crop"], o["dem_crop"]) and r["dem_crop"][i, j] == -32768 and not r["dem_crop"][:7].any()
    _compare(r["geo"], o["geo"], img, "NEAREST")
    assert o["num_valid"] > 1000 and abs(r["num_valid"] - o["num_valid"]) <= 2
    # wrong look side: all zeros
    pw = _capi.geozero_params(dem_shape=dem16.shape, side=-sc.side, length=sc.length, width=sc.width, **_grid_kw(sc, snwe))
    w = _capi.geozero_run(pw, dem16, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, img, method="BILINEAR")
    assert not w["geo"].any() and w["num_valid"] == 0 and w["num_outside_image"] == 0
    # errors come back as exceptions, not as a dead process
    with pytest.raises(_capi.B200Error, match="empty output grid"):
        _capi.geozero_run(_gparams(sc, (box[1], box[0], box[2], box[3]), dem16), dem16, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, img)
    with pytest.raises(_capi.B200Error, match="4 state vectors"):
        _capi.geozero_run(p, dem16, sc.orbit_t[:3], sc.orbit_pos[:3], sc.orbit_vel[:3], img)
    with pytest.raises(ValueError):
        _capi.geozero_run(p, dem16, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, img[:-1])


def test_geocode_component_end_to_end(tmp_path):
    sc = pu.rough_scene(200, 1600)
    dem, demf = _write_dem(sc, str(tmp_path / "dem.dem"), as_int16=True)
    c = pu.cpu_topo(sc, dem_method="BILINEAR", want_inc=False, want_mask=False)
    snwe = _inner_box(c["lat"], c["lon"], 0.05, 0.05)
    day = sc.sensing_start.replace(hour=0, minute=0, second=0, microsecond=0)
    orbit = Orbit.from_arrays(day, sc.orbit_t, sc.orbit_pos, sc.orbit_vel)
    # a 2-band BIL float product (like los.rdr) and a complex one (like an interferogram)
    los = np.stack([_textured(sc, 1), _textured(sc, 2)], axis=1)
    ifg = _textured(sc, 4, complex_=True)
    prods = {}
    for name, arr, dtype, bands, itype in (("los.rdr", los, "FLOAT", 2, "bil"), ("topophase.flat", ifg, "CFLOAT", 1, "cpx")):
        path = str(tmp_path / name)
        arr.tofile(path)
        img = IF.createImage()
        img.initImage(path, "read", sc.width, dtype, bands=bands, scheme="BIL")
        img.setLength(sc.length)
        img.imageType = itype
        img.addDescription("test product " + name)
        img.renderHdr()
        prods[name] = (IF.createImage().load(path + ".xml"), arr)
    for name, method in (("los.rdr", "nearest"), ("topophase.flat", "sinc")):
        inimg, arr = prods[name]
        ge = isce2_b200.createGeozero()
        ge.wireInputPort(name="planet", object=Planet(pname="Earth"))
        ge.wireInputPort(name="dem", object=dem)
        ge.wireInputPort(name="tobegeocoded", object=inimg)
        ge.snwe = snwe
        ge.demCropFilename = str(tmp_path / ("dem.crop." + name))
        ge.dopplerCentroidCoeffs = [0.0]
        ge.setSensingStart(sc.sensing_start)
        ge.rangeFirstSample = sc.r0
        ge.slantRangePixelSpacing = sc.dr
        ge.prf = sc.prf
        ge.radarWavelength = sc.wvl
        ge.lookSide = sc.side
        ge.orbit = orbit
        ge.numberRangeLooks = 1
        ge.numberAzimuthLooks = 1
        ge.geocode(method=method)
        assert ge.geoFilename == inimg.getFilename() + ".geo"
        hdr = IF.createImage().load(ge.geoFilename + ".xml")
        assert (hdr.width, hdr.length, hdr.bands, hdr.dataType, hdr.scheme) == (ge.geoWidth, ge.geoLength, inimg.bands,
                                                                                 inimg.dataType, "BIL")
        assert hdr.coord2.coordStart == ge.maximumGeoLatitude and hdr.coord2.coordDelta == sc.delta_lat
        assert hdr.coord1.coordStart == ge.minimumGeoLongitude and hdr.coord1.coordDelta == sc.delta_lon
        assert hdr.description == "test product " + name and os.path.exists(ge.geoFilename + ".vrt")
        assert ge.geoWidth == ge.computeGeoImageWidth()
        geo = np.fromfile(ge.geoFilename, arr.dtype).reshape((ge.geoLength,) + ((inimg.bands,) if inimg.bands > 1 else ()) + (ge.geoWidth,))
        kw = dict(dem=demf, orbit_t=sc.orbit_t, orbit_pos=sc.orbit_pos, orbit_vel=sc.orbit_vel, side=sc.side, **_grid_kw(sc, snwe))
        if inimg.bands > 1:
            for b in range(inimg.bands):
                o = orc.geozero(image=np.ascontiguousarray(arr[:, b, :]), method=method.upper(), **kw)
                _compare(np.ascontiguousarray(geo[:, b, :]), o["geo"], arr, method.upper())
        else:
            o = orc.geozero(image=arr, method=method.upper(), **kw)
            _compare(geo, o["geo"], arr, method.upper())
        crop = np.fromfile(ge.demCropFilename, np.int16).reshape(ge.geoLength, ge.geoWidth)
        assert np.array_equal(crop, o["dem_crop"])
        ch = IF.createDemImage().load(ge.demCropFilename + ".xml")
        assert (ch.width, ch.length, ch.dataType) == (ge.geoWidth, ge.geoLength, "SHORT")
        assert abs(ge.numValid - o["num_valid"]) <= 2 and ge.numValid > 0.5 * crop.size
