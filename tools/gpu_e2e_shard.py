#!/usr/bin/env python
"""End-to-end (host buffers) time of one rank's line block of the C2 swath, with the library's own breakdown."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isce2_b200 import _capi, synth


def main(world=8, rank=3, lines=13500, width=25000):
    sc = synth.make_scene(lines, width)
    sec = synth.make_scene(lines, width, dem=False, perturb=dict(da=120.0, d_cross=80.0, d_along_s=0.37))
    a, b = lines * rank // world, lines * (rank + 1) // world
    n = b - a
    p = _capi.topo_params(dem_shape=sc.dem.shape, first_lat=sc.first_lat, first_lon=sc.first_lon, delta_lat=sc.delta_lat,
                          delta_lon=sc.delta_lon, length=sc.length, width=sc.width, prf=sc.prf, t0=sc.t0, wvl=sc.wvl,
                          side=sc.side, peg_heading=sc.peg_heading, dem_method="BIQUINTIC", line0=a, nlines=n)
    dem = _capi.pinned_empty(sc.dem.shape, np.float32); dem[...] = sc.dem
    pe = _capi.pinned_empty
    outs = dict(lat=pe((n, width), np.float64), lon=pe((n, width), np.float64), hgt=pe((n, width), np.float64),
                los=pe((n, 2, width), np.float32), inc=pe((n, 2, width), np.float32), mask=pe((n, width), np.int8))
    gp = _capi.geo_params(length=lines - a, width=width, dem_shape=(n, width), r0=sc.r0 - 1.7, dr=sc.dr, prf=sc.prf,
                          t0=sc.t0 - 0.013 + a / sc.prf, wvl=sc.wvl, side=sc.side, out_f32=True)
    gout = dict(azt=None, rgm=None, azoff=pe((n, width), np.float32), rgoff=pe((n, width), np.float32))
    for rep in range(5):
        t0 = time.perf_counter()
        r = _capi.topo_run(p, dem, sc.orbit_t, sc.orbit_pos, sc.orbit_vel, sc.doppler_coeffs, [[sc.r0, sc.dr]], want_los=True,
                           want_inc=True, want_mask=True, out=outs)
        t1 = time.perf_counter()
        g = _capi.geo2rdr_run(gp, outs["lat"], outs["lon"], outs["hgt"], sec.orbit_t, sec.orbit_pos, sec.orbit_vel,
                              want=("azoff", "rgoff"), out=gout)
        t2 = time.perf_counter()
        print(f"world {world} rank {rank} rep {rep}: topo wall {1e3*(t1-t0):.1f} ms (setup {r['ms_setup']:.1f}, pipeline {r['ms_kernels']:.1f}, "
              f"total {r['ms_total']:.1f}); geo2rdr wall {1e3*(t2-t1):.1f} ms (setup {g['ms_setup']:.1f}, pipeline {g['ms_kernels']:.1f}, total {g['ms_total']:.1f})",
              flush=True)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 3)
