"""Shared helpers of the parity tests: run the CUDA path (through the C ABI) and the CPU oracle on one
synthetic scene and compare at the tolerances BASELINE.json's north_star states."""
import numpy as np

from isce2_b200 import _capi, synth
from oracle import oracle as orc

# north_star tolerances
TOL_LATLON_DEG = 1e-8
TOL_HGT_M = 1e-2
TOL_ANGLE_DEG = 1e-6
TOL_OFFSET_PX = 1e-3


def rough_scene(length, width, **kw):
    """Steeper terrain (amplitude ~ f^-1.6, 2.6 km relief) so that layover and shadow actually occur."""
    kw.setdefault("beta", 1.6)
    kw.setdefault("hmin", -100.0)
    kw.setdefault("hmax", 2500.0)
    return synth.make_scene(length, width, **kw)


def gpu_topo(sc, *, dem_method="BILINEAR", orbit_method="HERMITE", want_inc=True, want_mask=True, line0=0, nlines=-1,
             device=0, dem=None, **kw):
    p = _capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                          delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
                          side=sc.side, peg_heading=sc.peg_heading, a=sc.a, e2=sc.e2, dem_method=dem_method,
                          orbit_method=orbit_method, line0=line0, nlines=nlines, device=device, **kw)
    return _capi.topo_run(p, sc.dem if dem is None else dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs,
                          [[sc.r0, sc.dr * sc.nrnglooks]], want_los=True, want_inc=want_inc, want_mask=want_mask)


def cpu_topo(sc, **kw):
    return orc.topo(**orc.scene_topo_kwargs(sc, **kw))


def compare_topo(g, c):
    """Returns a dict of error statistics between GPU (g) and oracle (c) outputs."""
    st = {}
    for k, tol in (("lat", TOL_LATLON_DEG), ("lon", TOL_LATLON_DEG), ("hgt", TOL_HGT_M)):
        d = np.abs(g[k] - c[k])
        st[k] = dict(max=float(d.max()), n_over=int((d > tol).sum()), n_exact=int((d == 0).sum()), n=int(d.size), tol=tol)
    for k in ("los", "inc"):
        if g.get(k) is not None and c.get(k) is not None:
            d = np.abs(g[k].astype(np.float64) - c[k].astype(np.float64))
            # los channel 2 is an azimuth angle: compare modulo 360
            d = np.minimum(d, np.abs(d - 360.0))
            st[k] = dict(max=float(np.nanmax(d)), n_over=int((d > TOL_ANGLE_DEG).sum()), n_exact=int((d == 0).sum()),
                         n=int(d.size), tol=TOL_ANGLE_DEG, n_nan_mismatch=int((np.isnan(g[k]) != np.isnan(c[k])).sum()))
    if g.get("mask") is not None and c.get("mask") is not None:
        st["mask"] = dict(n_diff=int((g["mask"] != c["mask"]).sum()), n=int(g["mask"].size),
                          hist_gpu=np.bincount(g["mask"].ravel().astype(np.uint8), minlength=4).tolist(),
                          hist_cpu=np.bincount(c["mask"].ravel().astype(np.uint8), minlength=4).tolist())
    st["iters"] = dict(gpu=int(g["iterations"]), cpu=int(c["total_iters"]))
    st["converged"] = dict(gpu=int(g["converged"]), cpu=int(c["totalconv"]))
    st["bbox"] = dict(gpu=[g["min_lat"], g["max_lat"], g["min_lon"], g["max_lon"]],
                      cpu=[c["min_lat"], c["max_lat"], c["min_lon"], c["max_lon"]])
    st["crop"] = dict(gpu=[g["dem_x0"], g["dem_y0"], g["dem_nx"], g["dem_ny"]],
                      cpu=[c["ustartx"], c["ustarty"], c["udemwidth"], c["udemlength"]])
    return st


def secondary_kwargs(sc, sec, misreg_az=0.013, misreg_rg=1.7, recenter=0.0):
    """geo2rdr inputs of a secondary acquisition as topsStack builds them (contrib/stack/topsStack/geo2rdr.py:90-91):
    sensingStart - misreg_az, startingRange - misreg_rg.  The synthetic secondary flies 0.37 s ahead of the
    reference; short test scenes pass recenter=0.37 so that its acquisition window still covers the scene."""
    return dict(orbit_t=sec.orbit_t, orbit_pos=sec.orbit_pos, orbit_vel=sec.orbit_vel, length=sc.length, width=sc.width,
                r0=sc.r0 - misreg_rg, dr=sc.dr, prf=sc.prf, t0=sc.t0 - misreg_az - recenter, wvl=sc.wvl, side=sc.side)


def gpu_geo2rdr(lat, lon, hgt, kw, *, out_f32=False, orbit_method="HERMITE", bistatic=False, doppler_coeffs=(0.0,),
                line0=0, nlines=-1, device=0):
    p = _capi.geo_params(length=kw["length"], width=kw["width"], dem_shape=lat.shape, r0=kw["r0"], dr=kw["dr"],
                         prf=kw["prf"], t0=kw["t0"], wvl=kw["wvl"], side=kw["side"], orbit_method=orbit_method,
                         bistatic=bistatic, line0=line0, nlines=nlines, device=device, out_f32=out_f32)
    return _capi.geo2rdr_run(p, lat, lon, hgt, kw["orbit_t"], kw["orbit_pos"], kw["orbit_vel"], doppler_coeffs)


def compare_geo(g, c, dtaz=None):
    st = {}
    bad = -999999.0
    for k in ("azt", "rgm", "azoff", "rgoff"):
        if g.get(k) is None or c.get(k) is None:
            continue
        gv, cv = np.asarray(g[k], np.float64), np.asarray(c[k], np.float64)
        inval_g, inval_c = gv == bad, cv == bad
        both = ~inval_g & ~inval_c
        d = np.abs(gv[both] - cv[both])
        st[k] = dict(max=float(d.max()) if d.size else 0.0, n_valid_mismatch=int((inval_g != inval_c).sum()), n=int(gv.size),
                     n_exact=int((d == 0).sum()))
    st["iters"] = dict(gpu=int(g["iterations"]), cpu=int(c["total_iters"]))
    st["valid"] = dict(gpu=int(g["num_valid"]), cpu=int(c["num_valid"]))
    return st
