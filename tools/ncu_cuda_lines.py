#!/usr/bin/env python
"""Stall samples of one kernel per CUDA source line, from an `ncu --set full --import-source on` report:

    python tools/ncu_cuda_lines.py REPORT.ncu-rep KERNEL_REGEX [TOP_N]

Uses ncu's own source correlation (`--page source --print-source cuda,sass`): every SASS row is attributed to the CUDA
line it was generated from; prints the lines with the most stall samples and their dominant stall reasons.
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main(rep, kernel, top=40):
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}",
                                   "--print-source", "cuda,sass"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(txt)))
    fname, hdr = None, None
    agg = defaultdict(lambda: defaultdict(float))
    src = {}
    cur = None
    for r in rows:
        if len(r) == 2 and r[0] in ("File Name", "File Path"):
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))  # the second "Source" column (SASS text) overwrites the first: keep the CUDA text separately
        if r[0].strip():  # a CUDA line: its own row repeats the sum of the SASS rows below it, which are what is counted
            cur = (fname, int(r[0]))
            src[cur] = r[1].strip()
            continue
        if cur is None:
            continue
        try:
            n = float(d.get("# Samples") or 0)
        except ValueError:
            continue
        a = agg[cur]
        a["samples"] += n
        for k in ("stall_long_sb", "stall_wait", "stall_math", "stall_short_sb", "stall_no_inst", "stall_barrier", "stall_lg",
                  "stall_not_selected", "stall_branch_resolving", "stall_mio"):
            try:
                a[k] += float(d.get(k) or 0)
            except ValueError:
                pass
        try:
            a["inst"] += float(d.get("Instructions Executed") or 0)
        except ValueError:
            pass
    tot = sum(a["samples"] for a in agg.values()) or 1.0
    print(f"kernel {kernel}: {tot:.0f} samples")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:int(top)]:
        reasons = sorted(((k, v) for k, v in a.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:3]
        rs = " ".join(f"{k[6:]}={100 * v / max(a['samples'], 1):.0f}%" for k, v in reasons)
        print(f"{100 * a['samples'] / tot:5.1f}%  {key[0]}:{key[1]:<5d} inst {a['inst']:.3g}  [{rs}]  {src.get(key, '')[:90]}")


if __name__ == "__main__":
    main(*sys.argv[1:4])
