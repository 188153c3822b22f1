#!/usr/bin/env python
"""Where does the end-to-end time of one reference-facing call go? (run on the GPU box)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isce2_b200 import _capi, synth

def main(lines=1500, width=21000):
    sc = synth.make_scene(lines, width)
    sec = synth.make_scene(lines, width, dem=False, perturb=dict(da=120.0, d_cross=80.0, d_along_s=0.37))
    p = _capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                          delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
                          side=sc.side, peg_heading=sc.peg_heading)
    dem = _capi.pinned_empty(sc.dem.shape, np.float32); dem[...] = sc.dem
    outs = dict(lat=_capi.pinned_empty((lines, width), np.float64), lon=_capi.pinned_empty((lines, width), np.float64),
                hgt=_capi.pinned_empty((lines, width), np.float64), los=_capi.pinned_empty((lines, 2, width), np.float32),
                inc=None, mask=None)
    gp = _capi.geo_params(length=lines, width=width, dem_shape=(lines, width), r0=sc.r0 - 1.7, dr=sc.dr, prf=sc.prf,
                          t0=sc.t0 - 0.013, wvl=sc.wvl, side=sc.side, out_f32=True)
    gout = dict(azt=None, rgm=None, azoff=_capi.pinned_empty((lines, width), np.float32), rgoff=_capi.pinned_empty((lines, width), np.float32))
    for rep in range(4):
        t0 = time.perf_counter()
        r = _capi.topo_run(p, dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [[sc.r0, sc.dr]], want_los=True, out=outs)
        t1 = time.perf_counter()
        g = _capi.geo2rdr_run(gp, outs["lat"], outs["lon"], outs["hgt"], sec.orbit_t, sec.orbit_pos, sec.orbit_vel, want=("azoff", "rgoff"), out=gout)
        t2 = time.perf_counter()
        print(f"rep {rep}: topo wall {1e3*(t1-t0):.1f} ms (setup {r['ms_setup']:.1f}, kernels+pipeline {r['ms_kernels']:.1f}, total {r['ms_total']:.1f}); "
              f"geo2rdr wall {1e3*(t2-t1):.1f} ms (kernels+pipeline {g['ms_kernels']:.1f}, total {g['ms_total']:.1f})", flush=True)

if __name__ == "__main__":
    main()
    main(13500, 25000)
