#!/usr/bin/env python
"""Hot spots of one kernel from an `ncu --set full --import-source on` report: stall totals, the SASS instructions with the
most samples of a stall reason, and the dynamic opcode mix.

    python tools/ncu_source_hot.py gpurun_out/full.ncu-rep k_topo_final [stall_long_sb]
"""
import csv
import io
import subprocess
import sys
from collections import Counter


def main(rep, kernel, reason="stall_long_sb"):
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"], text=True,
                                  stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    idx = {h: i for i, h in enumerate(hdr)}
    data = []
    started = False
    for r in rows:
        if r and r[0] == "Address":
            if started:
                break  # only the first captured launch
            started = True
            continue
        if started and len(r) >= len(hdr):
            data.append(r)
    gi = lambda r, k: int(r[idx[k]] or 0)
    tot = sum(gi(r, "# Samples") for r in data)
    texec = sum(gi(r, "Instructions Executed") for r in data)
    print(f"{len(data)} SASS lines, {tot} samples, {texec} warp instructions")
    for c in [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]:
        v = sum(gi(r, c) for r in data)
        if v > 0.01 * tot:
            print(f"  {c:28s} {v:8d} {100.0 * v / tot:5.1f} %")
    print(f"== top by {reason}")
    for i, r in sorted(enumerate(data), key=lambda t: -gi(t[1], reason))[:16]:
        print(f"  {i:5d} {r[idx['Source']].strip()[:72]:72s} {gi(r, reason):7d} exec {gi(r, 'Instructions Executed')}")
    mix = Counter()
    for r in data:
        op = [o for o in r[idx["Source"]].split() if not o.startswith("@")]
        if op:
            mix[op[0].split(".")[0]] += gi(r, "Instructions Executed")
    print("== dynamic opcode mix")
    for k, v in mix.most_common(14):
        print(f"  {k:8s} {100.0 * v / texec:5.1f} %")


if __name__ == "__main__":
    main(*sys.argv[1:4])
