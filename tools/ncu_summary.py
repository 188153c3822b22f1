#!/usr/bin/env python
"""Condense an `ncu --set full` report into the JSON summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/full.ncu-rep profiles/rNN_ncu_full_summary.json

One record per captured launch: identification, duration, registers, DRAM bytes, pipe utilisation, the warp
stall breakdown and the FP64 instruction counts (dadd / dmul / dfma -> executed FP64 FLOP/s).
"""
import csv
import io
import json
import subprocess
import sys

KEEP_PREFIX = ("smsp__average_warps_issue_stalled_",)
KEEP = {
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "local_load_requests", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__inst_executed_per_warp.ratio", "launch__grid_size", "launch__block_size",
}


def fnum(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def main(rep, out):
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        rec = {}
        for h, u, v in zip(hdr, units, r):
            if h in ("Kernel Name", "Block Size", "Grid Size"):
                rec[h] = v
            elif h in KEEP or h.startswith(KEEP_PREFIX):
                rec[h] = f"{v} {u}".strip()
        da = fnum(rec.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "").split(" ")[0])
        dm = fnum(rec.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "").split(" ")[0])
        df = fnum(rec.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "").split(" ")[0])
        tv, tu = (rec.get("gpu__time_duration.sum", "0 ns").split(" ") + ["ns"])[:2]
        t = fnum(tv)
        scale = {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "s": 1.0, "second": 1.0}.get(tu, 1e-9)
        if None not in (da, dm, df, t) and t > 0:
            rec["derived_fp64_flops_executed"] = da + dm + 2 * df
            rec["derived_fp64_tflops_executed"] = (da + dm + 2 * df) / (t * scale) / 1e12
        recs.append(rec)
    json.dump(recs, open(out, "w"), indent=1)
    for rec in recs:
        print(rec["Kernel Name"][:70], rec.get("gpu__time_duration.sum"), rec.get("launch__registers_per_thread"),
              "fp64pipe", rec.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
              "TF", rec.get("derived_fp64_tflops_executed"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
