"""GPU parity of the remaining method paths of the two Components and of the host-memory kinds the
C ABI accepts:

  * geo2rdr with the SCH orbit interpolator (the non-polynomial kernel k_geo2rdr; orbit.c:119-172);
  * azimuth-varying 2-D Doppler polynomial in topo, alone and in the fused verb (Topozero.py:305-334);
  * lat / lon handed to the Geo2rdr component as Poly2D objects (Geo2rdr.py:216-226);
  * slant-range image input in the fused verb (slantRangeFilename, Topozero.py:337-347);
  * line blocks on two different devices (Topo.gpuDevices = [0, 1]) next to the same-device case;
  * an orbit that barely covers the scene: the pixels whose iterates leave the state-vector span are invalid exactly
    where the reference's 51-step loop says so (geo2rdr.f90:287-291);
  * results delivered into pageable host memory (numpy arrays, numpy.memmap over a file) equal those delivered into
    page-locked buffers.
"""
import datetime
import os

import numpy as np
import pytest

import isce2_b200
from isce2_b200 import _capi, image as IF, synth, synth_components as comp
from isce2_b200.orbit import Orbit
from isce2_b200.planet import Planet
from isce2_b200.poly import Poly2D
from oracle import oracle as orc
from tests import parity_util as pu
from tests.test_gpu_parity import _assert_topo

pytestmark = pytest.mark.gpu


def _same_orbit_kwargs(sc, dt0=0.0, dr0=0.0, pad=0):
    """geo2rdr window of the scene's own acquisition; pad > 0 widens it by that many lines / samples on every side, so
    that no pixel of the scene sits ON an edge of the window (where the strict bounds tests of geo2rdr.f90:308-316 are
    decided in the last bit)."""
    return dict(orbit_t=sc.orbit_t, orbit_pos=sc.orbit_pos, orbit_vel=sc.orbit_vel, length=sc.length + 2 * pad,
                width=sc.width + 2 * pad, r0=sc.r0 + dr0 - pad * sc.dr, dr=sc.dr, prf=sc.prf, t0=sc.t0 + dt0 - pad / sc.prf,
                wvl=sc.wvl, side=sc.side)


def test_geo2rdr_sch_orbit():
    sc = synth.config_c0(length=48, width=4096)
    c = pu.cpu_topo(sc, want_inc=False, want_mask=False)
    sec = synth.config_c1_secondary(length=sc.length, width=sc.width)
    kw = pu.secondary_kwargs(sc, sec, recenter=0.37)
    g = pu.gpu_geo2rdr(c["lat"], c["lon"], c["hgt"], kw, orbit_method="SCH")
    o = orc.geo2rdr(lat=c["lat"], lon=c["lon"], hgt=c["hgt"], orbit_method="SCH", **kw)
    st = pu.compare_geo(g, o)
    assert st["valid"]["gpu"] == st["valid"]["cpu"] and st["valid"]["cpu"] > 0.5 * c["lat"].size
    for k in ("azoff", "rgoff"):
        assert st[k]["n_valid_mismatch"] == 0 and st[k]["max"] < pu.TOL_OFFSET_PX, st[k]
    assert st["azt"]["max"] < 1e-8 and st["rgm"]["max"] < 1e-5
    # the same iteration as the reference's (no polynomial shortcut for this interpolator): same step count, up to the
    # pixels whose last step lands within rounding of the 5e-9 s stopping threshold
    # (the Lagrange sum over all state vectors is ill-conditioned: which side of the threshold a step lands on depends on
    # the last bits of the host's libm variant as well)
    assert abs(st["iters"]["gpu"] - st["iters"]["cpu"]) <= 5e-3 * st["iters"]["cpu"]


def test_topo_azimuth_varying_doppler_alone_and_fused():
    sc = synth.make_scene(40, 3072, sensor="nisar", beta=1.6, hmax=2500.0)
    # Hz vs (azimuth line, range pixel): second row = the azimuth derivative (Topozero.py:305-334 evaluates row by row)
    sc.doppler_coeffs = [[-120.0, 2.0e-2, -1.5e-6, 3e-11], [0.35, -4.0e-5, 0.0, 0.0], [-2e-3, 0.0, 0.0, 0.0]]
    c = pu.cpu_topo(sc, dem_method="BIQUINTIC", orbit_method="LEGENDRE")
    g = pu.gpu_topo(sc, dem_method="BIQUINTIC", orbit_method="LEGENDRE")
    _assert_topo(pu.compare_topo(g, c))
    flat = dict(sc.__dict__)
    sc0 = type(sc)(**{**flat, "doppler_coeffs": [sc.doppler_coeffs[0]]})
    g0 = pu.gpu_topo(sc0, dem_method="BIQUINTIC", orbit_method="LEGENDRE")
    assert np.abs(g0["lat"] - g["lat"]).max() > 1e-7  # the azimuth terms really move the solution
    # fused verb with the same polynomial
    p = _capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                          delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
                          side=sc.side, peg_heading=sc.peg_heading, a=sc.a, e2=sc.e2, dem_method="BIQUINTIC",
                          orbit_method="LEGENDRE")
    kw = _same_orbit_kwargs(sc, pad=3)
    dop1d = tuple(x / sc.prf for x in sc.doppler_coeffs[0])
    job = dict(params=_capi.geo_params(length=kw["length"], width=kw["width"], dem_shape=(sc.length, sc.width), r0=kw["r0"],
                                       dr=sc.dr, prf=sc.prf, t0=kw["t0"], wvl=sc.wvl, side=sc.side, orbit_method="LEGENDRE"),
               orbit=(sc.orbit_t, sc.orbit_pos, sc.orbit_vel), doppler=(dop1d, 0.0, 1.0), want=("azoff", "rgoff"))
    ft, fg = _capi.topo_geo2rdr_run(p, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [job],
                                    [[sc.r0, sc.dr]], want_los=True, want_inc=True, want_mask=True)
    for k in ("lat", "lon", "hgt", "los", "inc", "mask"):
        assert np.array_equal(ft[k], g[k], equal_nan=True), k
    o = orc.geo2rdr(lat=c["lat"], lon=c["lon"], hgt=c["hgt"], orbit_method="LEGENDRE", doppler_coeffs=dop1d, **kw)
    st = pu.compare_geo(fg[0], o)
    # (geo2rdr is given the range part of the Doppler only, so part of the grid solves to lines outside the short window)
    assert st["valid"]["gpu"] == st["valid"]["cpu"] and st["valid"]["cpu"] > 0.3 * sc.length * sc.width
    assert st["azoff"]["n_valid_mismatch"] == 0
    assert st["azoff"]["max"] < pu.TOL_OFFSET_PX and st["rgoff"]["max"] < pu.TOL_OFFSET_PX


def test_fused_verb_with_slant_range_image():
    sc = pu.rough_scene(24, 2048)
    rho = np.empty((sc.length, sc.width))
    slr = orc.Poly2D([[sc.r0, sc.dr]])
    for j in range(sc.width):
        rho[:, j] = slr(0.0, j)
    rho += 0.25 * sc.dr * np.sin(np.arange(sc.length))[:, None]  # line-dependent: only an image can carry this
    p = _capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                          delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
                          side=sc.side, peg_heading=sc.peg_heading, dem_method="BILINEAR")
    sec = synth.config_c1_secondary(length=sc.length, width=sc.width)
    kw = pu.secondary_kwargs(sc, sec, recenter=0.37)
    gp = _capi.geo_params(length=kw["length"], width=kw["width"], dem_shape=(sc.length, sc.width), r0=kw["r0"], dr=kw["dr"],
                          prf=kw["prf"], t0=kw["t0"], wvl=kw["wvl"], side=kw["side"])
    job = dict(params=gp, orbit=(kw["orbit_t"], kw["orbit_pos"], kw["orbit_vel"]), want=("azoff", "rgoff"))
    ft, fg = _capi.topo_geo2rdr_run(p, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [job], None,
                                    rho_image=rho, want_los=True, want_inc=True, want_mask=True)
    c = orc.topo(**orc.scene_topo_kwargs(sc, dem_method="BILINEAR", rho_image=rho))
    _assert_topo(pu.compare_topo(ft, c))
    o = orc.geo2rdr(lat=c["lat"], lon=c["lon"], hgt=c["hgt"], **kw)
    st = pu.compare_geo(fg[0], o)
    assert st["valid"]["gpu"] == st["valid"]["cpu"] and st["azoff"]["max"] < pu.TOL_OFFSET_PX and st["rgoff"]["max"] < pu.TOL_OFFSET_PX


def test_geo2rdr_component_with_poly2d_lat_lon(tmp_path):
    """Geo2rdr.py:216-226: latImage / lonImage may be Poly2D objects evaluated at the 0-based (line, sample) of the
    height image; here planes fitted to a topo run, so that most of the grid is inside the radar image."""
    sc = synth.config_c0(length=32, width=2048)
    c = pu.cpu_topo(sc, want_inc=False, want_mask=False)
    rows, cols = c["lat"].shape
    az, rg = np.arange(rows, dtype=np.float64), np.arange(cols, dtype=np.float64)
    A = np.stack([np.ones(rows * cols), np.tile((rg - cols / 2) / cols, rows), np.repeat((az - rows / 2) / rows, cols)], 1)
    polys = {}
    for k in ("lat", "lon"):
        co, *_ = np.linalg.lstsq(A, c[k].ravel(), rcond=None)
        p = Poly2D()
        p.initPoly(rangeOrder=1, azimuthOrder=1, coeffs=[[co[0], co[1]], [co[2], 0.0]])
        p.setMeanRange(cols / 2); p.setNormRange(float(cols)); p.setMeanAzimuth(rows / 2); p.setNormAzimuth(float(rows))
        p.setWidth(cols); p.setLength(rows)  # Geo2rdr.setDefaults compares them with the height image (Geo2rdr.py:283-289)
        polys[k] = p
    hgt_path = str(tmp_path / "hgt.rdr")
    c["hgt"].tofile(hgt_path)
    himg = IF.createImage()
    himg.initImage(hgt_path, "read", cols, "DOUBLE")
    himg.setLength(rows)
    himg.renderHdr()
    g = comp.make_geo2rdr(sc, sc, str(tmp_path), t0=sc.t0, r0=sc.r0, double=True)
    g.demImage, g.latImage, g.lonImage = himg, polys["lat"], polys["lon"]
    g.geo2rdr()
    # the oracle on the evaluated planes (evalPoly2d accumulation order, poly2d.c:92-111)
    ev = {k: np.array([[orc.Poly2D(polys[k].getCoeffs(), cols / 2, rows / 2, float(cols), float(rows))(i, j) for j in range(cols)]
                       for i in range(rows)]) for k in ("lat", "lon")}
    o = orc.geo2rdr(lat=ev["lat"], lon=ev["lon"], hgt=c["hgt"], **_same_orbit_kwargs(sc))  # (same window as the component's)
    az_off = np.fromfile(tmp_path / "azimuth.off").reshape(rows, cols)
    rg_off = np.fromfile(tmp_path / "range.off").reshape(rows, cols)
    assert np.array_equal(az_off == -999999.0, o["azoff"] == -999999.0)
    v = az_off != -999999.0
    assert v.mean() > 0.3
    assert np.abs(az_off[v] - o["azoff"][v]).max() < pu.TOL_OFFSET_PX and np.abs(rg_off[v] - o["rgoff"][v]).max() < pu.TOL_OFFSET_PX


@pytest.mark.parametrize("devices", [[0, 0], [0, 1]])
def test_topo_component_line_blocks_on_devices(tmp_path, devices):
    if max(devices) >= _capi.device_count():
        pytest.skip(f"needs {max(devices) + 1} CUDA devices")
    sc = pu.rough_scene(30, 2048)
    sec = synth.config_c1_secondary(length=sc.length, width=sc.width)
    dem_img = comp.prepare_dem(sc, str(tmp_path / "dem.dem"))
    outs = {}
    for tag, devs in (("one", [0]), ("blocks", devices)):
        d = tmp_path / tag
        comp.run_components(sc, sec, dem_img, str(d), devices=devs, misreg_az=0.013 + 0.37)
        outs[tag] = {f: np.fromfile(d / f, np.uint8) for f in sorted(os.listdir(d)) if f.endswith((".rdr", ".off"))}
    assert set(outs["one"]) == {"lat.rdr", "lon.rdr", "hgt.rdr", "los.rdr", "incLocal.rdr", "shadowMask.rdr", "range.off", "azimuth.off"}
    for f in outs["one"]:
        assert np.array_equal(outs["one"][f], outs["blocks"][f]), f


def test_geo2rdr_orbit_barely_covering_the_scene():
    """the polynomial kernel tests the span on the reference's first iterate and on its own Newton
    iterates, the reference on every one of its 9-11 iterates; with state vectors ending within a second of the scene
    the two must still agree on which pixels are invalid."""
    sc = synth.config_c0(length=1024, width=512)
    c = pu.cpu_topo(sc, want_inc=False, want_mask=False)
    dur = (sc.length - 1) / sc.prf
    n_mismatch = 0
    diag = []
    for lo, hi in ((0.05, 0.05), (0.6, 0.3), (-0.2, 0.4), (0.3, -0.25)):
        # resample the orbit so that its first / last state vector sit lo / hi seconds outside the scene's time span
        t = np.linspace(sc.t0 - lo, sc.t0 + dur + hi, 12)
        pos = np.array([synth.hermite_point(sc.orbit_t, sc.orbit_pos, sc.orbit_vel, x)[0] for x in t])
        vel = np.array([synth.hermite_point(sc.orbit_t, sc.orbit_pos, sc.orbit_vel, x)[1] for x in t])
        kw = dict(_same_orbit_kwargs(sc, pad=3), orbit_t=t, orbit_pos=pos, orbit_vel=vel)
        for method in ("HERMITE", "LEGENDRE"):
            g = pu.gpu_geo2rdr(c["lat"], c["lon"], c["hgt"], kw, orbit_method=method)
            o = orc.geo2rdr(lat=c["lat"], lon=c["lon"], hgt=c["hgt"], orbit_method=method, **kw)
            bad_g, bad_o = g["azoff"] == -999999.0, o["azoff"] == -999999.0
            n_mismatch += int((bad_g != bad_o).sum())
            if (bad_g != bad_o).any():
                ii = np.argwhere(bad_g != bad_o)
                diag.append(dict(case=(lo, hi, method), n=len(ii), first=ii[:6].tolist(), gpu_bad=int(bad_g[bad_g != bad_o].sum()),
                                 azt_gpu=[float(g["azt"][a, b]) for a, b in ii[:4]], azt_cpu=[float(o["azt"][a, b]) for a, b in ii[:4]],
                                 span=(float(t[0]), float(t[-1])), window=(kw["t0"], kw["t0"] + (kw["length"] - 1) / kw["prf"])))
            both = ~bad_g & ~bad_o
            assert both.sum() > 0.5 * both.size
            assert np.abs(g["azoff"][both] - o["azoff"][both]).max() < pu.TOL_OFFSET_PX, (lo, hi, method)
            if lo < 0 or hi < 0:
                assert bad_o.sum() > 0  # part of the scene really lies beyond the state vectors
    if diag:
        import json
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(diag, open(os.path.join("gpurun_out", "diag_orbit_span.json"), "w"), indent=1)
    assert n_mismatch == 0, diag


def _grow_and_map(path, dtype, shape, offset):
    """A writable mapping that starts `offset` bytes into an existing file (grown to hold the array)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    with open(path, "r+b") as f:
        f.truncate(offset + n)
    return np.memmap(str(path), dtype=dtype, mode="r+", shape=shape, offset=offset)


def test_pageable_and_page_locked_destinations_agree(tmp_path):
    """Results bound for pageable memory (plain numpy arrays, a numpy.memmap over a new file) go through the bounce ring
    and the copier threads, page-locked ones are written by DMA: same bytes.  Sized so that every layer spans several
    32 MB slots and a ragged tail."""
    sc = synth.config_c0(length=700, width=6000)
    p = _capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                          delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
                          side=sc.side, peg_heading=sc.peg_heading, dem_method="BILINEAR")
    sec = synth.config_c1_secondary(length=sc.length, width=sc.width)
    kw = pu.secondary_kwargs(sc, sec)
    gp = _capi.geo_params(length=kw["length"], width=kw["width"], dem_shape=(sc.length, sc.width), r0=kw["r0"], dr=kw["dr"],
                          prf=kw["prf"], t0=kw["t0"], wvl=kw["wvl"], side=kw["side"], out_f32=True)
    shapes = dict(lat=((sc.length, sc.width), np.float64), lon=((sc.length, sc.width), np.float64),
                  hgt=((sc.length, sc.width), np.float64), los=((sc.length, 2, sc.width), np.float32),
                  inc=((sc.length, 2, sc.width), np.float32), mask=((sc.length, sc.width), np.int8))

    def run(alloc):
        out = {k: alloc(k, s, d) for k, (s, d) in shapes.items()}
        gout = dict(azt=None, rgm=None, azoff=alloc("azoff", (sc.length, sc.width), np.float32),
                    rgoff=alloc("rgoff", (sc.length, sc.width), np.float32))
        job = dict(params=gp, orbit=(kw["orbit_t"], kw["orbit_pos"], kw["orbit_vel"]), want=("azoff", "rgoff"), out=gout)
        _capi.topo_geo2rdr_run(p, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [job], [[sc.r0, sc.dr]],
                               want_los=True, want_inc=True, want_mask=True, out=out)
        return {**out, "azoff": gout["azoff"], "rgoff": gout["rgoff"]}

    pinned = run(lambda k, s, d: _capi.pinned_empty(s, d))
    pageable = run(lambda k, s, d: np.full(s, 77, d))
    mm = run(lambda k, s, d: np.memmap(str(tmp_path / (k + ".bin")), dtype=d, mode="w+", shape=s))
    # ... and memmaps declared to the library as the file mappings they are are written with pwrite (image.file_backed):
    # the mapping sees the same pages; a mapping that starts inside its file (offset) lands where it should
    from isce2_b200 import image as IF

    def map_inside_a_file(k, s, d):
        with open(tmp_path / (k + ".fb"), "wb") as f:
            f.write(b"\xab" * 8192)  # a header the library must not touch
        return _grow_and_map(tmp_path / (k + ".fb"), d, s, 8192)

    pre = {k: map_inside_a_file(k, s, d) for k, (s, d) in shapes.items()}
    pre.update({k: map_inside_a_file(k, (sc.length, sc.width), np.float32) for k in ("azoff", "rgoff")})
    before = _capi.host_file_bytes()
    with IF.file_backed(list(pre.values())):
        fb = run(lambda k, s, d: pre[k])
    assert _capi.host_file_bytes() - before == sum(a.nbytes for a in pre.values())  # every byte went through pwrite
    assert not _capi.host_file_unregister(pre["lat"].ctypes.data)  # the context manager has taken the registrations back
    for k in pinned:
        assert np.array_equal(pinned[k], pageable[k], equal_nan=True), k
        assert np.array_equal(pinned[k], np.asarray(mm[k]), equal_nan=True), k
        mm[k].flush()
        assert np.array_equal(np.fromfile(tmp_path / (k + ".bin"), pinned[k].dtype).reshape(pinned[k].shape), pinned[k], equal_nan=True)
        assert np.array_equal(pinned[k], np.asarray(fb[k]), equal_nan=True), k
        raw = np.fromfile(tmp_path / (k + ".fb"), np.uint8)
        assert (raw[:8192] == 0xab).all(), k
        assert np.array_equal(raw[8192:].view(pinned[k].dtype).reshape(pinned[k].shape), pinned[k], equal_nan=True), k
    # inputs in pageable memory (memmaps of the rasters just written) go up through the mirror-image bounce path
    gp64 = _capi.geo_params(length=kw["length"], width=kw["width"], dem_shape=(sc.length, sc.width), r0=kw["r0"], dr=kw["dr"],
                            prf=kw["prf"], t0=kw["t0"], wvl=kw["wvl"], side=kw["side"])
    ins_mm = [np.memmap(str(tmp_path / (k + ".bin")), dtype=np.float64, mode="r", shape=(sc.length, sc.width)) for k in ("lat", "lon", "hgt")]
    ins_pin = []
    for k in ("lat", "lon", "hgt"):
        a = _capi.pinned_empty((sc.length, sc.width), np.float64)
        a[...] = pinned[k]
        ins_pin.append(a)
    ga = _capi.geo2rdr_run(gp64, *ins_mm, kw["orbit_t"], kw["orbit_pos"], kw["orbit_vel"])
    gb = _capi.geo2rdr_run(gp64, *ins_pin, kw["orbit_t"], kw["orbit_pos"], kw["orbit_vel"])
    read0 = _capi.host_file_bytes_read()
    with IF.file_backed([], inputs=[np.ascontiguousarray(m) for m in ins_mm]):  # ... or with pread, declared as the files they are
        gc = _capi.geo2rdr_run(gp64, *ins_mm, kw["orbit_t"], kw["orbit_pos"], kw["orbit_vel"])
    assert _capi.host_file_bytes_read() - read0 == sum(m.nbytes for m in ins_mm)
    for k in ("azt", "rgm", "azoff", "rgoff"):
        assert np.array_equal(ga[k], gb[k]), k
        assert np.array_equal(gc[k], gb[k]), k
    assert np.array_equal(ga["azoff"].astype(np.float32), pinned["azoff"])  # and they are the fused call's offsets
    plan = _capi.GeoPlan(gp64, *ins_mm)
    plan.execute(gp64, kw["orbit_t"], kw["orbit_pos"], kw["orbit_vel"], want=("azoff",))
    assert np.array_equal(plan.fetch()["azoff"], gb["azoff"])
    plan.close()
    # the plan form fetches through the same sink
    tp = _capi.TopoPlan(p, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [[sc.r0, sc.dr]], want_los=True,
                        want_inc=True, want_mask=True)
    tp.execute()
    f = tp.fetch()
    tp.close()
    for k in shapes:
        assert np.array_equal(f[k], pinned[k], equal_nan=True), k
