#!/usr/bin/env python
"""Attribute the executed instructions / stall samples of one kernel (ncu --set full --import-source on) to CUDA source
lines, by joining the report's per-SASS-instruction table with nvdisasm's line info of the same build.

    python tools/ncu_by_source_line.py REPORT.ncu-rep OBJECT.o KERNEL_REGEX MANGLED_NAME_SUBSTRING
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(obj, mangled):
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.check_output(["nvdisasm", "-g", os.path.join(d, cubin)], text=True, stderr=subprocess.DEVNULL)
    out, on, cur = [], False, ("?", 0)
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            on = mangled in ln
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((cur, m.group(2).strip()))
    return out


def main(rep, obj, kernel, mangled):
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"], text=True,
                                  stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    idx = {h: i for i, h in enumerate(hdr)}
    data, started = [], False
    for r in rows:
        if r and r[0] == "Address":
            if started:
                break
            started = True
            continue
        if started and len(r) >= len(hdr):
            data.append(r)
    sl = sass_lines(obj, mangled)
    if len(sl) != len(data):
        print(f"warning: {len(sl)} instructions in the object vs {len(data)} in the report (different build?)")
    n = min(len(sl), len(data))
    agg = defaultdict(lambda: [0, 0, 0, 0])  # executed, fp64 executed, samples, long_sb
    tot = [0, 0, 0, 0]
    for i in range(n):
        (f, line), text = sl[i]
        r = data[i]
        ex = int(r[idx["Instructions Executed"]] or 0)
        op = [o for o in text.split() if not o.startswith("@")][0].split(".")[0]
        fp = ex if op in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU") else 0
        v = [ex, fp, int(r[idx["# Samples"]] or 0), int(r[idx["stall_long_sb"]] or 0)]
        for k in range(4):
            agg[(f, line)][k] += v[k]
            tot[k] += v[k]
    print(f"total: {tot[0]} warp instructions, {tot[1]} FP64, {tot[2]} samples")
    for (f, line), v in sorted(agg.items(), key=lambda t: -t[1][2])[:45]:
        print(f"{f}:{line:<5d} exec {100.0 * v[0] / tot[0]:5.1f} %  fp64 {100.0 * v[1] / max(1, tot[1]):5.1f} %  samples {100.0 * v[2] / tot[2]:5.1f} %  "
              f"long_sb {100.0 * v[3] / max(1, tot[3]):5.1f} %")


if __name__ == "__main__":
    main(*sys.argv[1:5])
