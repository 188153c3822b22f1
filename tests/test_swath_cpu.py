"""Host logic of the merged-swath driver: union grid, burst windows, burst-view VRTs and the reader that
resolves them -- no GPU needed.  Reference behaviour: components/isceobj/TopsProc/runTopo.py:159-172, :316-319, :362-423."""
import datetime
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from isce2_b200 import image as IF, swath, synth


def test_union_grid_equals_the_scene_grid_and_windows_tile_it():
    sc, frames = synth.make_tops_acquisition(n_swaths=3, n_bursts=4, dem=False)
    g = swath.union_grid(frames)
    assert (g.length, g.width) == (sc.length, sc.width)
    assert g.t0 == sc.sensing_start and abs(g.r0 - sc.r0) < 1e-9 and abs(g.dt - 1.0 / sc.prf) < 1e-15
    cover = np.zeros((g.length, g.width), np.int32)
    for f in frames:
        for b in f.bursts:
            top, bottom, left, right = g.window(b)
            assert (top, bottom, left, right) == b.window
            cover[top:bottom, left:right] += 1
    assert cover.max() >= 2            # bursts overlap
    assert cover[0, 0] == 1 and cover[-1, -1] == 1


def test_union_grid_without_frame_level_properties():
    from types import SimpleNamespace
    sc, frames = synth.make_tops_acquisition(dem=False)
    bare = [SimpleNamespace(bursts=f.bursts) for f in frames]
    assert swath.union_grid(bare) == swath.union_grid(frames)


def test_merged_orbit_collects_vectors_outside_the_running_span():
    sc, frames = synth.make_tops_acquisition(dem=False, sv_per_burst=6)
    orb = swath.merged_orbit(frames)
    times = [sv.getTime() for sv in orb]
    assert times == sorted(times) and len(set(times)) == len(times)
    day = sc.sensing_start.replace(hour=0, minute=0, second=0, microsecond=0)
    all_t = {day + datetime.timedelta(seconds=float(t)) for t in sc.orbit_t}
    assert set(times) <= all_t and len(times) > 6


@pytest.mark.parametrize("bands,dtype,npdt", [(1, "DOUBLE", np.float64), (2, "FLOAT", np.float32), (1, "BYTE", np.int8)])
def test_burst_view_vrt_round_trip(tmp_path, bands, dtype, npdt):
    W, L = 50, 40
    parent = str(tmp_path / "geom" / "layer.rdr")
    img = IF.createImage()
    img.initImage(parent, "write", W, dtype, bands=bands, scheme="BIL")
    img.setLength(L)
    m = img.createImage()
    rng = np.random.default_rng(3)
    m[...] = (rng.uniform(-100, 100, m.shape)).astype(npdt)
    ref = np.array(m)
    img.finalizeImage()
    img.renderHdr()
    box = [5, 29, 7, 43]  # top, bottom, left, right
    dst = str(tmp_path / "geom" / "IW1" / "layer_03.rdr")
    os.makedirs(os.path.dirname(dst))
    swath.build_vrt(parent, dst, [W, L], box, bands=bands, dtype=dtype)
    # header of the view
    hdr = IF.createImage().load(dst + ".xml")
    assert (hdr.width, hdr.length, hdr.bands, hdr.dataType) == (36, 24, bands, dtype)
    # the VRT text is what buildVRT writes
    root = ET.parse(dst + ".vrt").getroot()
    assert root.get("rasterXSize") == "36" and root.get("rasterYSize") == "24"
    vb = root.findall("VRTRasterBand")
    assert len(vb) == bands
    gd = {"DOUBLE": "Float64", "FLOAT": "Float32", "BYTE": "UInt8"}[dtype]
    for i, b in enumerate(vb):
        assert b.get("dataType") == gd and b.get("band") == str(i + 1)
        s = b.find("SimpleSource")
        assert s.find("SourceFilename").text == os.path.join("..", "layer.rdr.vrt") and s.find("SourceFilename").get("relativeToVRT") == "1"
        assert s.find("SourceBand").text == str(i + 1)
        sp = s.find("SourceProperties")
        assert (sp.get("RasterXSize"), sp.get("RasterYSize"), sp.get("DataType")) == (str(W), str(L), gd)
        r = s.find("SrcRect")
        assert [r.get(k) for k in ("xOff", "yOff", "xSize", "ySize")] == ["7", "5", "36", "24"]
        d = s.find("DstRect")
        assert [d.get(k) for k in ("xOff", "yOff", "xSize", "ySize")] == ["0", "0", "36", "24"]
        assert b.find("NoDataValue").text == "0.0"
    # and it resolves to the window of the parent
    v = IF.read_view(dst)
    want = ref[5:29, 7:43] if bands == 1 else ref[5:29, :, 7:43]
    assert v.shape == want.shape and np.array_equal(np.asarray(v), want)
    # a plain raster resolves to itself
    assert np.array_equal(np.asarray(IF.read_view(parent + ".xml")), ref)


def test_build_vrt_rejects_unknown_types(tmp_path):
    with pytest.raises(Exception, match="Unsupported type"):
        swath.build_vrt(str(tmp_path / "a"), str(tmp_path / "b"), [4, 4], [0, 2, 0, 2], dtype="CFLOAT")


def test_stack_bookkeeping_without_a_device():
    st = swath.Geo2rdrStack(devices=[0, 1])
    z = np.zeros((4, 5))
    st.add_geometry("a", z, z, z)
    st.add_geometry("b", z, z, z)
    st.add_geometry("c", z, z, z)
    assert [st._geoms[k]["slot"] for k in "abc"] == [0, 1, 0]
    with pytest.raises(KeyError):
        st.add_job("nope", None, "r", "a")
    with pytest.raises(Exception, match="one shape"):
        st.add_geometry("d", z, z, np.zeros((4, 6)))
