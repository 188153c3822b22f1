#!/usr/bin/env python
"""GPU-vs-oracle mismatch statistics on a few synthetic scenes -> gpurun_out/diag.json (run on the GPU box)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isce2_b200 import _capi, synth  # noqa: E402
from tests import parity_util as pu  # noqa: E402


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    rep = {"device": _capi.device_name(0), "fp64_peak_tflops": _capi.fp64_peak(0), "cases": []}
    print(rep, flush=True)
    cases = [("smooth", synth.make_scene, 96, 8192), ("rough", pu.rough_scene, 96, 8192)]
    for tag, mk, L, W in cases:
        sc = mk(L, W)
        for method in ("BILINEAR", "BIQUINTIC", "BICUBIC", "NEAREST"):
            t0 = time.time()
            g = pu.gpu_topo(sc, dem_method=method)
            tg = time.time() - t0
            t0 = time.time()
            c = pu.cpu_topo(sc, dem_method=method)
            tc = time.time() - t0
            st = pu.compare_topo(g, c)
            st.update(case=tag, method=method, ms_kernels=g["ms_kernels"], ms_setup=g["ms_setup"], ms_total=g["ms_total"],
                      gpu_wall_s=tg, cpu_wall_s=tc, pixels=L * W)
            rep["cases"].append(st)
            print(json.dumps(st), flush=True)
            if method == "BILINEAR":
                sec = synth.config_c1_secondary(length=L, width=W)
                kw = pu.secondary_kwargs(sc, sec, recenter=0.37)
                for om in ("HERMITE", "LEGENDRE"):
                    if om == "LEGENDRE" and len(sec.orbit_t) < 9:
                        continue
                    gg = pu.gpu_geo2rdr(c["lat"], c["lon"], c["hgt"], kw, orbit_method=om)
                    cc = pu.orc.geo2rdr(lat=c["lat"], lon=c["lon"], hgt=c["hgt"], orbit_method=om, **kw)
                    s2 = pu.compare_geo(gg, cc)
                    s2.update(case=tag, kind="geo2rdr", orbit=om, ms_kernels=gg["ms_kernels"], pixels=L * W)
                    rep["cases"].append(s2)
                    print(json.dumps(s2), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
        json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
