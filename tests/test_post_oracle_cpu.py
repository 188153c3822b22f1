"""CPU pins of the multilooking / mask-projection restatements (other consumers):

* oracle.looks() against the reference's own takeLooks<T> / takeLookscpx<T> templates, compiled unchanged from
  components/mroipac/looks/bindings/looksmodule.cpp into oracle/_ref/libisce2_looks_ref.so (oracle/Makefile) and driven
  through an in-memory DataAccessor (oracle/ref_looks_shim.cpp): bit-identical for every element type;
* oracle.mask_to_radar() against golden output of the reference's own SWBDStitcher.toRadar
  (tests/golden/make_golden_post.py -> ref_toradar.npz);
* the host mirrors' file handling that needs no device, and the no-CPU-fallback rule for the new entry points.
"""
import ctypes as C
import os

import numpy as np
import pytest

from isce2_b200 import _capi
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LOOKS = os.path.join(ROOT, "oracle", "_ref", "libisce2_looks_ref.so")
DTYPES = [(np.int8, 0), (np.int16, 1), (np.int32, 2), (np.int64, 3), (np.float32, 4), (np.float64, 5), (np.complex64, 6)]


def _random(rng, shape, dt):
    if np.issubdtype(dt, np.integer):
        info = np.iinfo(dt)
        lo, hi = max(info.min, -2**40), min(info.max, 2**40)
        return rng.integers(lo, hi, shape, dtype=np.int64, endpoint=True).astype(dt)
    if dt == np.complex64:
        return (rng.normal(size=shape) * 1e3 + 1j * rng.normal(size=shape)).astype(dt)
    return (rng.normal(size=shape) * 10.0 ** rng.integers(-3, 6, shape)).astype(dt)


@pytest.mark.skipif(not os.path.exists(REF_LOOKS), reason="oracle/_ref not built (reference tree absent)")
@pytest.mark.parametrize("dt,code", DTYPES)
def test_looks_restatement_is_bit_identical_to_reference_templates(dt, code):
    ref = C.CDLL(REF_LOOKS)
    rng = np.random.default_rng(code)
    for (length, width, bands, ld, la) in ((23, 41, 1, 3, 4), (16, 30, 2, 4, 3), (9, 9, 3, 1, 1), (7, 50, 1, 7, 13), (5, 6, 2, 6, 2)):
        a = _random(rng, (length, width, bands), dt)  # pixel-interleaved lines, as the accessors deliver them
        out = np.zeros((length // ld, width // la, bands), dt)
        rc = ref.ref_take_looks(code, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), length, width, bands, ld, la)
        assert rc == 0
        mine = orc.looks(a, ld, la, scheme="BIP")
        assert mine.shape == out.shape and mine.dtype == out.dtype
        assert np.array_equal(mine.view(np.uint8), out.view(np.uint8)), (dt, length, width, bands, ld, la)


@pytest.mark.skipif(not os.path.exists(REF_LOOKS), reason="oracle/_ref not built (reference tree absent)")
def test_looks_restatement_random_shapes_against_reference_templates():
    """Forty random (shape, bands, looks, type) draws, including looks that do not divide the image and looks larger
    than the image (empty output)."""
    ref = C.CDLL(REF_LOOKS)
    rng = np.random.default_rng(2026)
    for _ in range(40):
        dt, code = DTYPES[int(rng.integers(0, len(DTYPES)))]
        length, width, bands = int(rng.integers(1, 40)), int(rng.integers(1, 60)), int(rng.integers(1, 4))
        ld, la = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        a = _random(rng, (length, width, bands), dt)
        out = np.zeros((length // ld, width // la, bands), dt)
        if out.size:
            assert ref.ref_take_looks(code, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), length, width, bands, ld, la) == 0
        mine = orc.looks(a, ld, la, scheme="BIP")
        assert mine.shape == out.shape
        assert np.array_equal(mine.view(np.uint8), out.view(np.uint8)), (dt, length, width, bands, ld, la)


def test_looks_restatement_layouts_and_nearest():
    rng = np.random.default_rng(3)
    a = rng.normal(size=(12, 2, 20)).astype(np.float32)  # BIL
    bil = orc.looks(a, 3, 4, scheme="BIL")
    bip = orc.looks(np.ascontiguousarray(np.moveaxis(a, 1, 2)), 3, 4, scheme="BIP")
    bsq = orc.looks(np.ascontiguousarray(np.moveaxis(a, 1, 0)), 3, 4, scheme="BSQ")
    assert np.array_equal(bil, np.moveaxis(bip, 2, 1)) and np.array_equal(bil, np.moveaxis(bsq, 0, 1))
    assert np.array_equal(bil[:, 0], orc.looks(a[:, 0], 3, 4))
    # exact means of integers, truncation toward zero of static_cast<T>
    b = np.array([[-3, -4, 5, 6], [-3, -3, 5, 5]], np.int16)
    assert orc.looks(b, 2, 2).tolist() == [[-3, 5]]  # -13/4 = -3.25 -> -3 ; 21/4 = 5.25 -> 5
    # nearest: source index floor((i + 0.5) * looks)
    c = np.arange(7 * 11).reshape(7, 11).astype(np.int32)
    n = orc.looks(c, 3, 4, method="NEAREST")
    assert n.tolist() == [[c[1, 2], c[1, 6]], [c[4, 2], c[4, 6]]]


def test_mask_projection_restatement_matches_reference_toRadar():
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_toradar.npz"))
    d = float(g["delta"])
    for tag, dt in (("f64", np.float64), ("f32", np.float32)):
        out = orc.mask_to_radar(g["mask"], float(g["start_lat"]), -d, float(g["start_lon"]), d, g["lat"].astype(dt), g["lon"].astype(dt))
        assert out.dtype == np.int8 and np.array_equal(out, g["out_" + tag]), tag
    assert set(np.unique(g["out_f64"]).tolist()) <= {0, 1, 2}


@pytest.mark.skipif(_capi.device_count() > 0, reason="a CUDA device is present")
def test_post_products_have_no_cpu_fallback():
    a = np.zeros((8, 8), np.float32)
    with pytest.raises(_capi.B200Error) as ei:
        _capi.looks_run(a, 2, 2)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)
    with pytest.raises(_capi.B200Error) as ei:
        _capi.mask_to_radar_run(np.zeros((4, 4), np.int8), 1.0, -0.1, 0.0, 0.1, np.zeros(5), np.zeros(5))
    assert ei.value.code == -2


def test_post_products_argument_validation_precedes_device_use():
    a = np.zeros((8, 8), np.float32)
    with pytest.raises(_capi.B200Error) as ei:
        _capi.looks_run(a, 0, 2)
    assert ei.value.code == -1 and "looks must be >= 1" in str(ei.value)
    with pytest.raises(TypeError):  # "Error. Unrecognized data type" (looksmodule.cpp:124-127)
        _capi.looks_run(a.astype(np.float16), 2, 2)
    with pytest.raises(_capi.B200Error) as ei:  # more column sums than a tile holds
        _capi.looks_run(np.zeros((2, 5000, 2), np.float32), 1, 2500, scheme="BIP")
    assert ei.value.code == -1
    with pytest.raises(TypeError):
        _capi.mask_to_radar_run(np.zeros((4, 4), np.float64), 1.0, -0.1, 0.0, 0.1, np.zeros(5), np.zeros(5))


def test_looks_component_with_fewer_lines_than_looks_writes_an_empty_raster(tmp_path):
    """Looks.py:44-45: outLength = length // downLooks may be 0; the reference's loops then do not execute.  Needs no
    device (nothing is computed)."""
    from isce2_b200 import image as IF, looks as LK
    np.zeros((3, 20), np.float32).tofile(tmp_path / "hgt.rdr")
    im = IF.createImage()
    im.initImage(str(tmp_path / "hgt.rdr"), "read", 20, "FLOAT")
    im.setLength(3)
    im.coord1.coordStart, im.coord1.coordDelta = 100.0, 2.0
    im.renderHdr()
    lk = LK.Looks()
    lk.setDownLooks(4)
    lk.setAcrossLooks(2)
    lk.setInputImage(im)
    lk.setOutputFilename(str(tmp_path / "ml" / "hgt.rdr"))
    o = lk.looks()
    assert (o.width, o.length) == (10, 0) and os.path.getsize(tmp_path / "ml" / "hgt.rdr") == 0
    assert os.path.exists(tmp_path / "ml" / "hgt.rdr.xml")
    # geo-referenced images: delta scales with the looks, start moves to the centre of the first look (Looks.py:49-55)
    assert o.coord1.coordDelta == 4.0 and o.coord1.coordStart == 101.0


def test_multilook_and_water_mask_host_logic_with_the_oracle_standing_in(tmp_path, monkeypatch):
    """File handling of the host mirrors (names, XML / VRT, .full copies, layouts, data types) without a device: for this
    test only, the two library calls are replaced by the oracle restatements."""
    from isce2_b200 import image as IF, looks as LK, watermask as WM

    def fake_looks(image, ld, la, *, scheme="BIL", method="AVERAGE", out=None, device=0):
        return orc.looks(np.asarray(image), ld, la, scheme=scheme, method=method), dict(ms_kernels=0.0, ms_total=0.0, gpu_launches=1)

    def fake_mask(mask, lat0, dlat, lon0, dlon, lat, lon, *, out=None, device=0):
        return orc.mask_to_radar(np.asarray(mask), lat0, dlat, lon0, dlon, lat, lon), dict(ms_kernels=0.0, ms_total=0.0, gpu_launches=1)

    monkeypatch.setattr(_capi, "looks_run", fake_looks)
    monkeypatch.setattr(_capi, "mask_to_radar_run", fake_mask)
    rng = np.random.default_rng(4)
    L, W = 23, 31
    geom = tmp_path / "geom_full"
    geom.mkdir()
    layers = {"hgt": (rng.normal(size=(L, W)) * 100, "DOUBLE", 1), "lat": (35.0 - 1e-3 * rng.random((L, W)), "DOUBLE", 1),
              "lon": (-118.0 + 1e-3 * rng.random((L, W)), "DOUBLE", 1), "los": (rng.normal(size=(L, 2, W)).astype(np.float32), "FLOAT", 2),
              "shadowMask": (rng.integers(0, 4, (L, W)).astype(np.int8), "BYTE", 1)}
    for name, (arr, dt, bands) in layers.items():
        arr.tofile(geom / (name + ".rdr"))
        im = IF.createImage()
        im.initImage(str(geom / (name + ".rdr")), "read", W, dt, bands=bands, scheme="BIL")
        im.setLength(L)
        im.renderHdr()
        im.renderVRT()
    # water mask: geocoded BYTE raster + header with its geographic coordinates
    wb = rng.integers(-1, 1, (40, 50)).astype(np.int8)
    wb.tofile(tmp_path / "swbd.wbd")
    wim = IF.createImage()
    wim.initImage(str(tmp_path / "swbd.wbd"), "read", 50, "BYTE")
    wim.setLength(40)
    wim.coord1.coordStart, wim.coord1.coordDelta, wim.coord1.coordSize = -118.0, 2.5e-5, 50
    wim.coord2.coordStart, wim.coord2.coordDelta, wim.coord2.coordSize = 35.0, -2.5e-5, 40
    wim.renderHdr()
    rdr = WM.geo2radar(str(tmp_path / "swbd.wbd"), str(geom / "waterMask.rdr"), str(geom / "lat.rdr"), str(geom / "lon.rdr"))
    assert rdr == str(geom / "waterMask.rdr")
    wm = np.fromfile(geom / "waterMask.rdr", np.int8).reshape(L, W)
    assert np.array_equal(wm, orc.mask_to_radar(wb, 35.0, -2.5e-5, -118.0, 2.5e-5, layers["lat"][0], layers["lon"][0]))
    h = IF.createImage().load(str(geom / "waterMask.rdr.xml"))
    assert (h.dataType, h.width, h.length) == ("BYTE", W, L)
    IF.createImage().load(str(geom / "waterMask.rdr.xml")).renderVRT()  # runMultilook only takes layers that have all three files

    for method, omethod in (("isce", "AVERAGE"), ("gdal", "NEAREST")):
        out_dir = tmp_path / ("geom_" + method)
        assert LK.runMultilook(str(geom), str(out_dir), 4, 3, method=method) == str(out_dir)
        for name, (arr, dt, bands) in list(layers.items()) + [("waterMask", (wm, "BYTE", 1))]:
            want = orc.looks(arr, 4, 3, scheme="BIL", method=omethod)
            got = np.fromfile(out_dir / (name + ".rdr"), arr.dtype).reshape(want.shape)
            assert np.array_equal(got, want), (method, name)
            hh = IF.createImage().load(str(out_dir / (name + ".rdr.xml")))
            src_scheme = IF.createImage().load(str(geom / (name + ".rdr.xml"))).scheme  # the output keeps the input's interleaving
            assert (hh.width, hh.length, hh.bands, hh.dataType, hh.scheme) == (W // 3, L // 4, bands, dt, src_scheme)
            for ext in (".rdr.vrt", ".rdr.full.xml", ".rdr.full.vrt"):
                assert (out_dir / (name + ext)).exists(), (name, ext)
        assert not (out_dir / "incLocal.rdr").exists()  # absent inputs are skipped (topo.py:381)
    with pytest.raises(ValueError):
        LK.runMultilook(str(geom), str(tmp_path / "x"), 2, 2, method="bogus")
