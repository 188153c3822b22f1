#!/usr/bin/env python
"""Build profiles/kernel_counters.json -- what bench.py needs from ncu to state executed FP64 and DRAM traffic per launch.

    python tools/kernel_counters.py TAG OUT.json  c2=SUMMARY.json:BENCH.log [c0c1=SUMMARY.json:BENCH.log ...]

SUMMARY.json: tools/ncu_summary.py output of an `ncu --set full` capture of `bench.py --lines 1500 --workload X`;
BENCH.log: stdout + stderr of that same bench run (its JSON line gives pixels per launch and the iterations per pixel K).
Per kernel instance (named as bench.py names them): executed FP64 flop per pixel (dadd + dmul + 2 dfma thread
instructions / pixels of the launch), DRAM bytes per pixel (dram__bytes_read.sum + dram__bytes_write.sum), and the pipe /
occupancy figures quoted in DESIGN.md.
"""
import json
import re
import sys

DEM = {0: "SINC", 1: "BILINEAR", 2: "BICUBIC", 3: "NEAREST", 4: "AKIMA", 5: "BIQUINTIC"}
ORB = {0: "HERMITE", 1: "SCH", 2: "LEGENDRE"}


def num(s):
    try:
        return float(str(s).split(" ")[0].replace(",", ""))
    except Exception:
        return None


def gb(s):
    v, u = (str(s).split(" ") + [""])[:2]
    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)


def ms(s):
    v, u = (str(s).split(" ") + [""])[:2]
    return float(v) * {"ms": 1.0, "msecond": 1.0, "us": 1e-3, "usecond": 1e-3, "ns": 1e-6, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(u, 1.0)


def name_of(kernel, orbit_method):
    m = re.search(r"(k_\w+)<(\d+)", kernel)
    if not m:
        return None
    base, a = m.group(1), int(m.group(2))
    if base.startswith("k_geo2rdr"):
        return f"{base}<{orbit_method}>"
    if base.startswith("k_topo"):
        return f"{base}<{DEM[a]}>"
    return base


def main(tag, out, pairs):
    res = {"_comment": "ncu --set full per launch of `bench.py --lines 1500 --workload X` (tools/gpu_pass.sh), reduced by "
                       "tools/kernel_counters.py; fp64_flop_per_pixel = (dadd + dmul + 2 dfma thread instructions) / pixels of the "
                       "launch; dram_bytes_per_pixel = (dram__bytes_read.sum + dram__bytes_write.sum) / pixels.  bench.py multiplies "
                       "by the pixels of its own launch (and, for the iterative kernels, by K / per_iteration_K).",
           "tag": tag, "kernels": {}, "by_workload": {}}
    for pair in pairs:
        wkey, pair = pair.split("=", 1)
        summ, log = pair.split(":")
        line = None
        for ln in open(log, errors="replace"):
            if ln.startswith("{") and '"metric"' in ln:
                line = json.loads(ln)
        if line is None:
            print("no JSON line in", log)
            continue
        px = line["config"]["pixels_per_step"]
        K = line["config"]["K_topo_iters_per_pixel"]
        om = line["config"]["orbit_method"]
        seen = set()
        for r in json.load(open(summ)):
            nm = name_of(r["Kernel Name"], om)
            if nm is None or nm in seen or "k_fp64_peak" in r["Kernel Name"]:
                if "k_fp64_peak" in r["Kernel Name"]:
                    res["fp64_peak_kernel"] = {"ms": ms(r["gpu__time_duration.sum"]),
                                               "fp64_pipe_pct": num(r.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")),
                                               "tflops_executed": r.get("derived_fp64_tflops_executed")}
                continue
            seen.add(nm)
            fl = r.get("derived_fp64_flops_executed")
            d = dict(workload=line["config"]["workload"], pixels_of_profiled_launch=px, ms=ms(r["gpu__time_duration.sum"]),
                     registers=num(r.get("launch__registers_per_thread")),
                     fp64_flop_per_pixel=(fl / px) if fl else None, tflops_executed=r.get("derived_fp64_tflops_executed"),
                     dram_bytes_per_pixel=(gb(r["dram__bytes_read.sum"]) + gb(r["dram__bytes_write.sum"])) / px,
                     fp64_pipe_pct=num(r.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")),
                     warps_active_pct=num(r.get("sm__warps_active.avg.pct_of_peak_sustained_active")),
                     l1_hit_pct=num(r.get("l1tex__t_sector_hit_rate.pct")), l2_hit_pct=num(r.get("lts__t_sector_hit_rate.pct")),
                     stall_long_scoreboard=num(r.get("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio")),
                     stall_wait=num(r.get("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio")),
                     stall_math_pipe=num(r.get("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio")),
                     stall_no_instruction=num(r.get("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio")))
            if nm.startswith(("k_topo_solve", "k_topo_fused")):
                d["per_iteration_K"] = K
            res["by_workload"].setdefault(wkey, {})[nm] = d
            res["kernels"].setdefault(nm, d)  # first workload listed wins the un-keyed entry
    json.dump(res, open(out, "w"), indent=1)
    for wk, kk in res["by_workload"].items():
      for k, v in kk.items():
        print(f"{wk:5s} {k:32s} {v['ms']:8.3f} ms  regs {v['registers']}  fp64 {v['fp64_flop_per_pixel'] and round(v['fp64_flop_per_pixel'])} flop/px  "
              f"{v['tflops_executed'] and round(v['tflops_executed'], 2)} TF  dram {v['dram_bytes_per_pixel']:.1f} B/px  pipe {v['fp64_pipe_pct']}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3:])
