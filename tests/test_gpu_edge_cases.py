"""Edge cases of the hot path on the GPU against the oracle: ragged and tiny image shapes (every kernel's tail handling:
the solve kernel's line segments and lane refill, the one-thread-per-pixel kernels' last CTA, the layover pass on lines
shorter than a warp), a DEM that is smaller than the scene (every DEM index clamps, topozero.f90:525-536 / 666-677), a DEM
with voids and spikes (the -500 m floor of :541), and geo2rdr grids of one line / a few samples."""
import dataclasses

import numpy as np
import pytest

from isce2_b200 import _capi
from oracle import oracle as orc
from tests import parity_util as pu
from tests.test_gpu_parity import _assert_topo

pytestmark = pytest.mark.gpu

RAGGED = [(1, 2), (1, 33), (3, 31), (2, 257), (5, 1025), (2, 1500), (1, 4099)]


@pytest.mark.parametrize("method", ["BIQUINTIC", "BILINEAR"])  # split solve / final kernels and the fused kernel
@pytest.mark.parametrize("shape", RAGGED, ids=[f"{a}x{b}" for a, b in RAGGED])
def test_ragged_and_tiny_shapes(method, shape):
    sc = pu.rough_scene(*shape)
    g = pu.gpu_topo(sc, dem_method=method)
    c = pu.cpu_topo(sc, dem_method=method)
    st = pu.compare_topo(g, c)
    _assert_topo(st)
    assert st["mask"]["hist_gpu"] == st["mask"]["hist_cpu"]


def test_single_sample_lines_are_an_argument_error():
    # width 1: the reference's layover pass indexes zsch(2) of a one-element line (topozero.f90:748-755) and its bounding box
    # corners coincide; the library refuses such a grid instead of guessing (B200_EINVAL), with or without the mask
    sc = pu.rough_scene(1, 1)
    for want_mask in (False, True):
        with pytest.raises(_capi.B200Error) as ei:
            pu.gpu_topo(sc, dem_method="BIQUINTIC", want_mask=want_mask)
        assert ei.value.code == -1 and "bad radar grid" in str(ei.value)


def _cut_dem(sc, frac_lat=(0.35, 0.65), frac_lon=(0.3, 0.7)):
    """The scene with only the middle of its DEM: pixels whose iterates fall outside clamp to the edge posts."""
    ny, nx = sc.dem.shape
    y0, y1 = int(frac_lat[0] * ny), int(frac_lat[1] * ny)
    x0, x1 = int(frac_lon[0] * nx), int(frac_lon[1] * nx)
    return dataclasses.replace(sc, dem=np.ascontiguousarray(sc.dem[y0:y1, x0:x1]), first_lat=sc.first_lat + y0 * sc.delta_lat,
                               first_lon=sc.first_lon + x0 * sc.delta_lon)


@pytest.mark.parametrize("method", ["BIQUINTIC", "BILINEAR", "SINC"])
def test_dem_smaller_than_the_scene(method):
    sc = _cut_dem(pu.rough_scene(24, 3000))
    g = pu.gpu_topo(sc, dem_method=method)
    c = pu.cpu_topo(sc, dem_method=method)
    st = pu.compare_topo(g, c)
    _assert_topo(st)
    assert st["crop"]["gpu"] == [1, 1, sc.dem.shape[1], sc.dem.shape[0]]  # the whole (small) DEM is the crop


@pytest.mark.parametrize("method", ["BIQUINTIC", "BILINEAR", "NEAREST", "AKIMA"])
def test_dem_with_voids_and_spikes(method):
    sc = pu.rough_scene(16, 2500)
    clean = pu.cpu_topo(sc, dem_method=method, want_inc=False, want_mask=False)
    dem = sc.dem.copy()
    rng = np.random.default_rng(11)
    ny, nx = dem.shape
    for _ in range(400):  # SRTM-style void blocks and single-post spikes spread over the footprint
        y, x = int(rng.integers(0, ny - 8)), int(rng.integers(0, nx - 8))
        dem[y:y + int(rng.integers(1, 8)), x:x + int(rng.integers(1, 8))] = -32768.0
    ys, xs = rng.integers(0, ny, 3000), rng.integers(0, nx, 3000)
    dem[ys, xs] = 8800.0
    sc = dataclasses.replace(sc, dem=dem)
    g = pu.gpu_topo(sc, dem_method=method)
    c = pu.cpu_topo(sc, dem_method=method)
    st = pu.compare_topo(g, c)
    _assert_topo(st)
    assert (np.abs(c["hgt"] - clean["hgt"]) > 50.0).sum() >= 50  # the voids / spikes are actually hit
    assert c["totalconv"] < 0.97 * c["lat"].size  # and send pixels through the secondary iterations


@pytest.mark.parametrize("shape", [(1, 1), (1, 5), (3, 31), (2, 1025)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_geo2rdr_tiny_grids(shape):
    sc = pu.rough_scene(*shape)
    c = pu.cpu_topo(sc, want_inc=False, want_mask=False)
    kw = dict(orbit_t=sc.orbit_t, orbit_pos=sc.orbit_pos, orbit_vel=sc.orbit_vel, length=sc.length, width=sc.width, r0=sc.r0 - 1.7,
              dr=sc.dr, prf=sc.prf, t0=sc.t0 - 0.013, wvl=sc.wvl, side=sc.side)
    for out_f32 in (False, True):
        g = pu.gpu_geo2rdr(c["lat"], c["lon"], c["hgt"], kw, out_f32=out_f32)
        o = orc.geo2rdr(lat=c["lat"], lon=c["lon"], hgt=c["hgt"], **kw)
        if out_f32:
            for k in ("azt", "rgm", "azoff", "rgoff"):
                o[k] = o[k].astype(np.float32)
        st = pu.compare_geo(g, o)
        assert st["valid"]["gpu"] == st["valid"]["cpu"]
        for k in ("azoff", "rgoff"):
            assert st[k]["n_valid_mismatch"] == 0 and st[k]["max"] < pu.TOL_OFFSET_PX, st[k]


def test_empty_line_block_is_an_argument_error():
    sc = pu.rough_scene(4, 64)
    with pytest.raises(_capi.B200Error):
        pu.gpu_topo(sc, line0=4, nlines=0)
    with pytest.raises(_capi.B200Error):
        pu.gpu_topo(sc, line0=4, nlines=-1)  # "to the end" from the end


def _ridge_scene(length=24, width=8000):
    """Flat ground with two north-south ridges of 59 and 68 degrees of slope (3000 m over 60 posts, 1500 m over 20): the slant
    range of the layover pass's cross-track grid runs backwards over ~950 and ~300 consecutive samples of every line,
    i.e. disorder windows far beyond the length at which the co-sort's rank counting switches to its tiled group walk
    (B2_RANK_TILE_MIN, topo_kernels.cu)."""
    from isce2_b200 import synth
    sc = synth.make_scene(length, width, hmin=0.0, hmax=1.0)
    ny, nx = sc.dem.shape
    x = np.arange(nx)
    dem = np.zeros((ny, nx), np.float32)
    for x0, h, half in ((nx * 0.45, 3000.0, 60), (nx * 0.6, 1500.0, 20)):
        dem = np.maximum(dem, np.maximum(0.0, h * (1 - np.abs(x - x0) / half)).astype(np.float32)[None, :])
    return dataclasses.replace(sc, dem=dem)


@pytest.mark.parametrize("method", ["BIQUINTIC", "BILINEAR"])
def test_long_range_fold_over(method):
    sc = _ridge_scene()
    g = pu.gpu_topo(sc, dem_method=method)
    c = pu.cpu_topo(sc, dem_method=method)
    st = pu.compare_topo(g, c)
    _assert_topo(st)
    hist = st["mask"]["hist_cpu"]
    assert hist[2] + hist[3] > 0.1 * c["mask"].size and hist[1] > 0.1 * c["mask"].size  # layover and shadow on every line
    assert st["mask"]["hist_gpu"] == hist
