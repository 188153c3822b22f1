#!/usr/bin/env python
"""Kernel timings + parity spot check of the current build (run on the GPU box).  Prints one JSON line per case."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isce2_b200 import _capi, synth  # noqa: E402
from tests import parity_util as pu  # noqa: E402


def perf(lines=1500, width=25000, reps=3, rough=False):
    sc = pu.rough_scene(lines, width) if rough else synth.make_scene(lines, width)
    sec = synth.make_scene(lines, width, dem=False, perturb=dict(da=120.0, d_cross=80.0, d_along_s=0.37))
    out = {}
    for method, inc, mask in ((("BIQUINTIC", True, True),) if rough else (("BILINEAR", False, False), ("BIQUINTIC", True, True))):
        p = _capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                              delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
                              side=sc.side, peg_heading=sc.peg_heading, dem_method=method)
        tp = _capi.TopoPlan(p, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [[sc.r0, sc.dr]],
                            want_los=True, want_inc=inc, want_mask=mask)
        best = None
        for _ in range(reps):
            tp.execute()
            import ctypes as C
            res = _capi.TopoResult()
            e = C.create_string_buffer(512)
            _capi._check(_capi.lib().b200_topo_plan_fetch(tp.handle, None, C.byref(res), e, 512), e)
            if best is None:
                best = (res.ms_pixels, res.ms_mask, res.iterations / float(lines * width), res.ms_solve)
            else:  # every time is the best of the repetitions on its own (a first call can carry one-off costs)
                best = (min(best[0], res.ms_pixels), min(best[1], res.ms_mask), best[2], min(best[3], res.ms_solve))
        out[method] = dict(ms_pixels=round(best[0], 3), ms_solve=round(best[3], 3), ms_mask=round(best[1], 3), K=round(best[2], 3),
                           gpix_s=round(lines * width / best[0] / 1e6, 3))
        if method == "BILINEAR":
            gp = _capi.geo_params(length=sc.length, width=sc.width, dem_shape=(sc.length, sc.width), r0=sc.r0 - 1.7, dr=sc.dr,
                                  prf=sc.prf, t0=sc.t0 - 0.013, wvl=sc.wvl, side=sc.side, out_f32=True)
            g = _capi.GeoPlan(gp, topo_plan=tp)
            ms = min(g.execute(gp, sec.orbit_t, sec.orbit_pos, sec.orbit_vel, want=("azoff", "rgoff")) for _ in range(reps))
            out["GEO2RDR_HERMITE"] = dict(ms=round(ms, 3), gpix_s=round(lines * width / ms / 1e6, 3))
            g.close()
        tp.close()
    return out


def parity(L=96, W=8192):
    sc = pu.rough_scene(L, W)
    rep = {}
    for method in ("BILINEAR", "BIQUINTIC"):
        g = pu.gpu_topo(sc, dem_method=method)
        c = pu.cpu_topo(sc, dem_method=method)
        st = pu.compare_topo(g, c)
        rep[method] = {k: (st[k]["n_over"], float("%.3g" % st[k]["max"]), round(st[k]["n_exact"] / st[k]["n"], 4))
                       for k in ("lat", "lon", "hgt", "los", "inc")}
        rep[method]["mask_diff"] = st["mask"]["n_diff"]
        rep[method]["iters_equal"] = st["iters"]["gpu"] == st["iters"]["cpu"]
    return rep


def geozero_perf(lines=13500, width=25000, reps=3):
    """Geocode a full-swath product (float32 and complex64) onto the 1-arcsec DEM grid under the C2 footprint."""
    sc = synth.make_scene(lines, width)
    # the scene builder pads the DEM by 0.2 deg around the footprint: geocode the DEM's inner part
    nlat, nlon = sc.dem.shape
    snwe = (sc.first_lat + (nlat - 1 - 600) * sc.delta_lat, sc.first_lat + 600 * sc.delta_lat,
            sc.first_lon + 600 * sc.delta_lon, sc.first_lon + (nlon - 1 - 600) * sc.delta_lon)
    p = _capi.geozero_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                             delta_lon=sc.delta_lon, snwe=snwe, length=lines, width=width, r0=sc.r0, dr=sc.dr, prf=sc.prf,
                             t0=sc.t0, wvl=sc.wvl, side=sc.side)
    _capi.GeozeroPlan(p, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel).close()  # module load, workspace cache
    t0 = time.perf_counter()
    plan = _capi.GeozeroPlan(p, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel)
    t_plan = time.perf_counter() - t0
    r0 = plan.fetch()
    npx = r0["geo_length"] * r0["geo_width"]
    out = {"grid": [r0["geo_length"], r0["geo_width"]], "Mpx": npx / 1e6, "ms_setup": round(r0["ms_setup"], 3), "ms_solve": round(r0["ms_solve"], 3),
           "plan_wall_ms": round(1e3 * t_plan, 1), "iters_per_px": round(r0["iterations"] / npx, 3)}
    rng = np.random.default_rng(1)
    real = rng.normal(size=(lines, width)).astype(np.float32)
    for name, img in (("f32", real), ("c64", (real + 1j * real[::-1]).astype(np.complex64))):
        for m in ("NEAREST", "BILINEAR", "BICUBIC", "SINC"):
            best, wall = 1e30, 1e30
            for _ in range(reps):
                t0 = time.perf_counter()
                plan.geocode(img, method=m)
                wall = min(wall, time.perf_counter() - t0)
                best = min(best, plan.ms_kernels)
            r = plan.fetch()
            out[f"{name}_{m}"] = {"ms_gather": round(best, 3), "wall_ms": round(1e3 * wall, 1), "valid_frac": round(r["num_valid"] / npx, 3),
                                  "gpix_s": round(npx / best / 1e6, 2)}
    plan.close()
    return out


def resamp_perf(lines=1500, width=25000, reps=3):
    """Resample one S1 burst (complex64) with float32 .off residuals, a TOPS-like azimuth carrier and flattening."""
    rng = np.random.default_rng(2)
    z = (rng.normal(size=(lines, width)) + 1j * rng.normal(size=(lines, width))).astype(np.complex64)
    y, x = np.mgrid[0:lines, 0:width]
    ra = (0.4 + 1e-4 * y).astype(np.float32)
    rr = (-0.7 + 1e-5 * x).astype(np.float32)
    out = {}
    for name, kw in (("zero_carrier_zero_doppler", {}),
                     ("az_carrier_doppler_flatten", dict(az_carrier=[[0.0, 1e-4], [0.3, 0.0], [2e-4, 0.0]], doppler=[[0.02, 1e-6]],
                                                         flatten=True, ref_r0=100.0))):
        best, wall = 1e30, 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            r = _capi.resamp_slc_run(z, z.shape, resid_az=ra, resid_rg=rr, **kw)
            wall = min(wall, time.perf_counter() - t0)
            best = min(best, r["ms_kernels"])
        npx = lines * width
        out[name] = {"ms_kernels": round(best, 3), "wall_ms": round(1e3 * wall, 1), "gpix_s": round(npx / best / 1e6, 2),
                     "valid_frac": round(r["num_valid"] / npx, 3), "launches": r["gpu_launches"],
                     "algorithmic_GBs": round(npx * (8 + 8 + 8) / best / 1e6, 1)}
    return out


def post_perf(lines=1500, width=25000, ld=14, la=4):
    """Multilooking from the resident topo layers (device time of the kernel alone) and through the host-buffer verbs."""
    sc = synth.make_scene(lines, width)
    p = _capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                          delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
                          side=sc.side, peg_heading=sc.peg_heading, dem_method="BILINEAR")
    tp = _capi.TopoPlan(p, sc.dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [[sc.r0, sc.dr]],
                        want_los=True, want_inc=False, want_mask=True)
    tp.execute()
    out = {"looks": [ld, la]}
    npx = lines * width
    for layer, bpp in (("lat", 8), ("los", 8), ("mask", 1)):
        for method in ("AVERAGE", "NEAREST"):
            ms = min(tp.looks(layer, ld, la, method=method)[1]["ms_kernels"] for _ in range(5))
            algo = npx * bpp * (1.0 + 1.0 / (ld * la)) if method == "AVERAGE" else 2.0 * npx * bpp / (ld * la)
            out[f"{layer}_{method}"] = {"ms_kernel": round(ms, 4), "algorithmic_GBs": round(algo / ms / 1e6, 1)}
    full = tp.fetch()
    tp.close()
    t0 = time.perf_counter()
    _, r = _capi.looks_run(full["lat"], ld, la)
    out["host_verb_lat"] = {"wall_ms": round(1e3 * (time.perf_counter() - t0), 1), "ms_total": round(r["ms_total"], 1),
                            "launches": r["gpu_launches"]}
    mask = (sc.dem < np.median(sc.dem)).astype(np.int8)
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        _, r = _capi.mask_to_radar_run(mask, sc.first_lat, sc.delta_lat, sc.first_lon, sc.delta_lon, full["lat"], full["lon"])
        best = min(best, time.perf_counter() - t0)
    out["mask_to_radar"] = {"wall_ms": round(1e3 * best, 1), "ms_kernels_incl_h2d_wait": round(r["ms_kernels"], 2),
                            "launches": r["gpu_launches"]}
    return out


if __name__ == "__main__":
    if "--post" in sys.argv:
        print(json.dumps({"post": post_perf()}), flush=True)
        sys.exit(0)
    if "--resamp" in sys.argv:
        print(json.dumps({"resamp_slc": resamp_perf()}), flush=True)
        sys.exit(0)
    if "--geozero" in sys.argv:
        print(json.dumps({"geozero": geozero_perf()}), flush=True)
        sys.exit(0)
    tag = os.environ.get("B200GEOM_LIB", "default")
    print(json.dumps({"lib": tag, "perf": perf()}), flush=True)
    if "--rough" in sys.argv:
        print(json.dumps({"lib": tag, "perf_rough_terrain": perf(rough=True)}), flush=True)
    if "--no-parity" not in sys.argv:
        print(json.dumps({"lib": tag, "parity(n_over,max,exact_frac)": parity()}), flush=True)
