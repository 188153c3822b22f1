#!/usr/bin/env python
"""The reference's own sequence through the Components, files on tmpfs (run on the GPU box): createTopozero().topo(), then
createGeo2rdr().geo2rdr() on the lat / lon / hgt rasters topo wrote -- seconds per call, with the rasters declared as the
files they are (pwrite / pread, the default) and without (B200_FILE_WRITES=0: stores / loads through the mappings).

    python tools/component_two_call.py [LINES]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import contextlib, json, os, shutil, sys, tempfile
sys.path.insert(0, %r)
import bench
from isce2_b200 import _capi, synth_components as comp
w, sc, sec = bench.build_workload("c2", %s)
base = "/dev/shm" if os.path.isdir("/dev/shm") else None
demdir = tempfile.mkdtemp(prefix="b200_two_dem_", dir=base)
rows = []
try:
    dem_img = comp.prepare_dem(sc, os.path.join(demdir, "dem.dem"))
    for i in range(3):
        d = tempfile.mkdtemp(prefix="b200_two_", dir=base)
        try:
            w0, r0 = _capi.host_file_bytes(), _capi.host_file_bytes_read()
            with contextlib.redirect_stdout(sys.stderr):
                info = comp.run_components_separately(sc, sec, dem_img, d, dem_method=w["dem_method"], orbit_method=w["orbit_method"],
                                                      inc=w["inc"], mask=w["mask"], devices=[0])
            rows.append(dict(topo_s=round(info["seconds_topo"], 3), geo2rdr_s=round(info["seconds_geo2rdr"], 3),
                             GB_pwrite=round((_capi.host_file_bytes() - w0) / 1e9, 2), GB_pread=round((_capi.host_file_bytes_read() - r0) / 1e9, 2)))
        finally:
            shutil.rmtree(d, ignore_errors=True)
finally:
    shutil.rmtree(demdir, ignore_errors=True)
print(json.dumps({"file_writes": os.environ.get("B200_FILE_WRITES", "1"), "pixels": sc.pixels, "calls": rows}))
"""


def main():
    lines = sys.argv[1] if len(sys.argv) > 1 else "None"
    for fw in ("1", "0"):
        env = dict(os.environ, B200_FILE_WRITES=fw)
        out = subprocess.run([sys.executable, "-c", CHILD % (ROOT, lines)], env=env, capture_output=True, text=True)
        print((out.stdout.strip().splitlines() or [out.stderr[-600:]])[-1], flush=True)


if __name__ == "__main__":
    main()
