"""Pins of the geozero oracle (oracle/zerodop_oracle.c: orc_geozero) that need no GPU.

The reference holds no test or fixture for geozero, and its Fortran cannot be built here, so the
restatement is pinned by (a) the primitives it shares with geo2rdr (orbit / polynomial evaluators: bit-identical to the
reference's C sources, tests/test_oracle_pins.py), (b) the inverse relation with topozero: geozero's (azimuth, range)
image coordinates of a DEM node, pushed through the topo oracle's lat/lon layers, must give back that node, and
(c) exactness properties of the four interpolators as the reference calls them (index conventions, f_delay offsets,
band / complex handling)."""
import numpy as np
import pytest

from isce2_b200 import synth
from oracle import oracle as orc
from tests import parity_util as pu


def _scene(flat=None, length=240, width=1800, **kw):
    sc = synth.make_scene(length, width, **kw)
    if flat is not None:
        sc.dem = np.full_like(sc.dem, flat)
    return sc


def _grid_kw(sc, snwe):
    return dict(first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat, delta_lon=sc.delta_lon, snwe=snwe,
                r0=sc.r0, dr=sc.dr, prf=sc.prf, t0=sc.t0, wvl=sc.wvl)


def _inner_box(lat, lon, frac_lat=0.15, frac_lon=0.15):
    """Box shrunk (or, with negative fractions, grown) by a fraction of the footprint's extent on every side."""
    pad_lat = frac_lat * float(lat.max() - lat.min())
    pad_lon = frac_lon * float(lon.max() - lon.min())
    return (float(lat.min()) + pad_lat, float(lat.max()) - pad_lat, float(lon.min()) + pad_lon, float(lon.max()) - pad_lon)


def _bilin(a, y, x):
    y0, x0 = np.floor(y).astype(int), np.floor(x).astype(int)
    fy, fx = y - y0, x - x0
    return (a[y0, x0] * (1 - fy) * (1 - fx) + a[y0, x0 + 1] * (1 - fy) * fx + a[y0 + 1, x0] * fy * (1 - fx)
            + a[y0 + 1, x0 + 1] * fy * fx)


def test_geozero_inverts_topo_on_a_flat_dem():
    sc = _scene(flat=250.0)
    c = pu.cpu_topo(sc, dem_method="BILINEAR")
    snwe = _inner_box(c["lat"], c["lon"])
    img = np.zeros((sc.length, sc.width), np.float32)
    r = orc.geozero(dem=sc.dem, image=img, orbit_t=sc.orbit_t, orbit_pos=sc.orbit_pos, orbit_vel=sc.orbit_vel,
                    method="BILINEAR", side=sc.side, **_grid_kw(sc, snwe))
    g = r["grid"]
    assert r["geo"].shape == (g["geo_len"], g["geo_wid"]) and g["geo_len"] > 20 and g["geo_wid"] > 50
    v = np.isfinite(r["az_idx"]) & (r["az_idx"] > 2) & (r["az_idx"] < sc.length - 2) & (r["rng_idx"] > 2) & (r["rng_idx"] < sc.width - 2)
    assert v.mean() > 0.3
    glat = (sc.first_lat + (g["max_lat_idx"] + np.arange(g["geo_len"])) * sc.delta_lat)[:, None] * np.ones((1, g["geo_wid"]))
    glon = (sc.first_lon + (g["min_lon_idx"] + np.arange(g["geo_wid"])) * sc.delta_lon)[None, :] * np.ones((g["geo_len"], 1))
    # 1-based fractional image coordinates -> topo layers (double)
    la = _bilin(c["lat"], r["az_idx"][v] - 1.0, r["rng_idx"][v] - 1.0)
    lo = _bilin(c["lon"], r["az_idx"][v] - 1.0, r["rng_idx"][v] - 1.0)
    # topo stops at |range residual| <= 5 cm; a pixel is 2.3 m x 14 m: 1e-6 deg ~ 0.1 m pins the coordinates to ~0.03 px
    assert np.abs(la - glat[v]).max() < 1e-6 and np.abs(lo - glon[v]).max() < 1e-6
    assert np.array_equal(r["dem_crop"], np.full_like(r["dem_crop"], 250))
    assert r["num_valid"] + r["num_outside_image"] == g["geo_len"] * g["geo_wid"] and r["num_outside_dem"] == 0
    # geographic limits reported back (geozero.f90:419-422)
    assert abs(r["geomax_lat"] - glat[0, 0]) < 1e-12 and abs(r["geomin_lon"] - glon[0, 0]) < 1e-12


def test_interpolators_on_images_with_known_answers():
    sc = _scene(flat=0.0, length=200, width=1500)
    c = pu.cpu_topo(sc, dem_method="BILINEAR")
    snwe = _inner_box(c["lat"], c["lon"], -0.2, -0.2)  # a box larger than the footprint: the image edges are inside it
    kw = dict(dem=sc.dem, orbit_t=sc.orbit_t, orbit_pos=sc.orbit_pos, orbit_vel=sc.orbit_vel, side=sc.side, **_grid_kw(sc, snwe))
    rr, aa = np.meshgrid(np.arange(1, sc.width + 1, dtype=np.float64), np.arange(1, sc.length + 1, dtype=np.float64))
    ramp = (0.25 * rr - 1.5 * aa + 3.0)  # exactly representable in float32 at these sizes
    assert np.array_equal(ramp.astype(np.float32).astype(np.float64), ramp)
    out = {m: orc.geozero(image=ramp.astype(np.float32), method=m, **kw) for m in ("BILINEAR", "NEAREST", "SINC", "BICUBIC")}
    b = out["BILINEAR"]
    v = np.isfinite(b["az_idx"]) & (b["geo"] != 0)
    want = 0.25 * b["rng_idx"] - 1.5 * b["az_idx"] + 3.0
    # bilinear reproduces a plane (float32 output rounding only)
    assert np.abs(b["geo"][v] - want[v]).max() < 2e-4
    n = out["NEAREST"]
    vn = np.isfinite(n["az_idx"]) & (n["geo"] != 0)
    wn = 0.25 * np.round(n["rng_idx"]) - 1.5 * np.round(n["az_idx"]) + 3.0
    assert np.array_equal(n["geo"][vn].astype(np.float64), wn[vn])
    # a constant image stays constant under the normalised sinc and the bicubic
    const = np.full((sc.length, sc.width), 7.5, np.float32)
    for m in ("SINC", "BICUBIC"):
        r = orc.geozero(image=const, method=m, **kw)
        inside = np.isfinite(r["az_idx"]) & (r["az_idx"] > 9) & (r["az_idx"] < sc.length - 9) & (r["rng_idx"] > 9) & (r["rng_idx"] < sc.width - 9)
        assert inside.sum() > 1000 and np.abs(r["geo"][inside] - 7.5).max() < 1e-5
    # sinc of a plane: symmetric kernel quantised to 1/8192 px
    s = out["SINC"]
    vs = np.isfinite(s["az_idx"]) & (s["geo"] != 0)
    ws = 0.25 * s["rng_idx"] - 1.5 * s["az_idx"] + 3.0
    assert np.abs(s["geo"][vs] - ws[vs]).max() < 0.05
    # f_delay margins: sinc needs 4 samples of margin, bicubic 3, bilinear / nearest 2 (geozeroMethods.F:66-79)
    assert out["SINC"]["num_valid"] < out["BICUBIC"]["num_valid"] < out["BILINEAR"]["num_valid"] == out["NEAREST"]["num_valid"]


def test_complex_image_is_interpolated_componentwise_and_real_images_ignore_the_imaginary_part():
    sc = _scene(length=160, width=1200)
    c = pu.cpu_topo(sc, dem_method="BILINEAR")
    snwe = _inner_box(c["lat"], c["lon"])
    kw = dict(dem=sc.dem, orbit_t=sc.orbit_t, orbit_pos=sc.orbit_pos, orbit_vel=sc.orbit_vel, side=sc.side, **_grid_kw(sc, snwe))
    rng = np.random.default_rng(5)
    re = rng.normal(size=(sc.length, sc.width)).astype(np.float32)
    im = rng.normal(size=(sc.length, sc.width)).astype(np.float32)
    for m in ("BILINEAR", "BICUBIC", "SINC", "NEAREST"):
        z = orc.geozero(image=(re + 1j * im).astype(np.complex64), method=m, **kw)["geo"]
        a = orc.geozero(image=re, method=m, **kw)["geo"]
        b = orc.geozero(image=im, method=m, **kw)["geo"]
        assert np.array_equal(z.real, a) and np.array_equal(z.imag, b), m


def test_bad_dem_samples_wrong_look_side_and_rows_outside_the_dem():
    sc = _scene(length=120, width=900)
    c = pu.cpu_topo(sc, dem_method="BILINEAR")
    snwe = _inner_box(c["lat"], c["lon"])
    dem = sc.dem.copy()
    g0 = orc.geozero_grid(orc.geozero_params(dem_shape=dem.shape, length=sc.length, width=sc.width, **_grid_kw(sc, snwe)))
    i, j = g0["geo_len"] // 2, g0["geo_wid"] // 2  # the footprint is a tilted parallelogram: its centre is surely imaged
    dem[g0["max_lat_idx"] + i, g0["min_lon_idx"] + j] = -32768.0  # SRTM void
    img = np.ones((sc.length, sc.width), np.float32)
    kw = dict(orbit_t=sc.orbit_t, orbit_pos=sc.orbit_pos, orbit_vel=sc.orbit_vel, **_grid_kw(sc, snwe))
    r = orc.geozero(dem=dem, image=img, method="NEAREST", side=sc.side, **kw)
    assert r["geo"][i, j] == 0 and r["dem_crop"][i, j] == -32768 and np.isnan(r["az_idx"][i, j])
    assert r["geo"][i, j + 1] == 1 and r["geo"][i + 1, j] == 1
    # the radar looks the other way: nothing is geocoded, nothing is counted
    w = orc.geozero(dem=dem, image=img, method="NEAREST", side=-sc.side, **kw)
    assert not w["geo"].any() and w["num_valid"] == 0 and w["num_outside_image"] == 0
    # a box reaching north of the DEM: those lines are skipped and counted demwidth each (geozero.f90:250-253)
    north = (snwe[0], sc.first_lat + 5.5 * abs(sc.delta_lat), snwe[2], snwe[3])
    kw2 = dict(kw)
    kw2["snwe"] = north
    n = orc.geozero(dem=dem, image=img, method="NEAREST", side=sc.side, **kw2)
    assert n["grid"]["max_lat_idx"] == -5
    assert n["num_outside_dem"] == 5 * dem.shape[1] and not n["geo"][:5].any() and not n["dem_crop"][:5].any()


def test_geozero_solve_against_the_reference_pythons_own_geo2rdr(golden):
    """tests/golden/ref_python_vectors.json holds single-point solutions computed by importing the reference's Python
    (isceobj.Orbit.Orbit.rdr2geo / geo2rdr on its 15-vector orbit fixture): geozero's fixed-point iteration solves the
    same zero-Doppler equation, so the image coordinates it assigns to that ground point must be the golden azimuth
    time and range (to geozero's own stopping tolerance of 5e-7 s and the 1 us resolution of Python datetimes)."""
    rows = np.array(golden["orbit_rsc"])
    n_checked = 0
    for e in golden["rdr2geo"]:
        lat, lon, h = e["llh"][0], e["llh"][1], e["height"]
        spacing, half = 1.0 / 1200, 40
        dem = np.full((2 * half + 1, 2 * half + 1), h, np.float32)
        first_lat, first_lon = lat + half * spacing, lon - half * spacing
        prf, dr = 1000.0, 1.0
        t0, r0 = e["t"] - 0.5, e["rng"] - 2000.0
        L, W = 1001, 4001
        snwe = (lat - 2.2 * spacing, lat + 2.2 * spacing, lon - 2.2 * spacing, lon + 2.2 * spacing)
        p = orc.geozero_params(dem_shape=dem.shape, first_lat=first_lat, first_lon=first_lon, delta_lat=-spacing,
                               delta_lon=spacing, snwe=snwe, length=L, width=W, r0=r0, dr=dr, prf=prf, t0=t0, wvl=0.056)
        g = orc.geozero_grid(p)
        r = orc.geozero(dem=dem, image=np.zeros((L, W), np.float32), orbit_t=rows[:, 0], orbit_pos=rows[:, 1:4],
                        orbit_vel=rows[:, 4:7], method="NEAREST", side=e["side"], first_lat=first_lat, first_lon=first_lon,
                        delta_lat=-spacing, delta_lon=spacing, snwe=snwe, r0=r0, dr=dr, prf=prf, t0=t0, wvl=0.056)
        i, j = half - g["max_lat_idx"], half - g["min_lon_idx"]  # the grid node that is the golden point
        assert 0 <= i < g["geo_len"] and 0 <= j < g["geo_wid"]
        t_geo = t0 + (r["az_idx"][i, j] - 1.0) / prf
        rng_geo = r0 + (r["rng_idx"][i, j] - 1.0) * dr
        assert abs(t_geo - e["t"]) < 6e-7 and abs(t_geo - e["geo2rdr_t"]) < 2e-6, (t_geo, e["t"], e["geo2rdr_t"])
        assert abs(rng_geo - e["rng"]) < 2e-4 and abs(rng_geo - e["geo2rdr_rng"]) < 2e-4, (rng_geo, e["rng"])
        n_checked += 1
    assert n_checked >= 4
