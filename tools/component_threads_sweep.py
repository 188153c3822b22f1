#!/usr/bin/env python
"""How the Component path (createTopozero().topo() chained with createGeo2rdr(), files on tmpfs) and the file-write floors
scale with the number of host threads (run on the GPU box):

    python tools/component_threads_sweep.py [LINES]

The copier pool is sized once per process (B200_COPY_THREADS), so every setting runs in its own child process."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r"""
import contextlib, json, os, shutil, sys, tempfile, time
sys.path.insert(0, %r)
import bench
from isce2_b200 import synth_components as comp
w, sc, sec = bench.build_workload("c2", %s)
base = "/dev/shm" if os.path.isdir("/dev/shm") else None
demdir = tempfile.mkdtemp(prefix="b200_sweep_dem_", dir=base)
times, lib = [], []
try:
    dem_img = comp.prepare_dem(sc, os.path.join(demdir, "dem.dem"))
    for i in range(3):
        d = tempfile.mkdtemp(prefix="b200_sweep_", dir=base)
        try:
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(sys.stderr):
                info = comp.run_components(sc, sec, dem_img, d, dem_method=w["dem_method"], orbit_method=w["orbit_method"],
                                           inc=w["inc"], mask=w["mask"], devices=[0])
            times.append(time.perf_counter() - t0)
            lib.append([round(float(g["ms_total"]), 1) for g in (info.get("gpu_timings") or [])])
        finally:
            shutil.rmtree(d, ignore_errors=True)
finally:
    shutil.rmtree(demdir, ignore_errors=True)
print(json.dumps({"copy_threads": os.environ.get("B200_COPY_THREADS"), "file_writes": os.environ.get("B200_FILE_WRITES", "0"),
                  "sink_slots": os.environ.get("B200_SINK_SLOTS"), "file_parts": os.environ.get("B200_FILE_PARTS"),
                  "step_s": [round(t, 3) for t in times], "library_ms": lib}))
"""


def main():
    lines = sys.argv[1] if len(sys.argv) > 1 else "None"
    import bench
    w, sc, sec = bench.build_workload("c2", None if lines == "None" else int(lines))
    nbytes = sc.pixels * 49
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    for how in ("mmap", "pwrite"):
        for th in (4, 8, 12, 16):
            r = bench.file_write_floor(base, nbytes, threads=th, how=how)
            print(json.dumps({"floor": how, "threads": th, "seconds": round(r["seconds"], 3), "GBps": round(r["GBps"], 2)}), flush=True)
    combos = [dict(B200_COPY_THREADS=t, B200_FILE_WRITES=0) for t in (4, 8, 12, 16)]
    if os.environ.get("SWEEP") == "file":  # pwrite variants against the default
        combos = [dict(B200_COPY_THREADS=16, B200_FILE_WRITES=0), dict(B200_COPY_THREADS=16, B200_FILE_WRITES=0, B200_SINK_SLOTS=16)]
        combos += [dict(B200_COPY_THREADS=t, B200_FILE_WRITES=1, B200_SINK_SLOTS=sl, B200_FILE_PARTS=pp)
                   for t, sl, pp in ((16, 6, 1), (16, 16, 1), (16, 24, 1), (16, 16, 2), (16, 16, 4), (8, 16, 1))]
    for combo in combos:
        env = dict(os.environ, **{k: str(v) for k, v in combo.items()})
        out = subprocess.run([sys.executable, "-c", CHILD % (ROOT, lines)], env=env, capture_output=True, text=True)
        print((out.stdout.strip().splitlines() or [out.stderr[-400:]])[-1], flush=True)


if __name__ == "__main__":
    main()
