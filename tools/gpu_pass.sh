#!/bin/bash
# One GPU pass of a round: parity tests, both bench arms, ncu launch list + full captures of every kernel instance of the
# step.  Usage (from the repo root, under gpurun):  bash tools/gpu_pass.sh TAG [quick]
TAG=${1:-r02}
QUICK=${2:-}
O=gpurun_out
mkdir -p $O
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/box_$TAG.txt; nproc >> $O/box_$TAG.txt; free -g >> $O/box_$TAG.txt
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu_full_$TAG.log 2>&1; tail -40 $O/pytest_gpu_full_$TAG.log > $O/pytest_gpu_$TAG.log; cat $O/pytest_gpu_$TAG.log
timeout 300 python tools/gpu_perf.py 2>&1 | tail -1 > $O/parity_rough_$TAG.json
[ -n "$AB_MASK" ] && { timeout 600 bash tools/ab_mask_modes.sh > $O/ab_mask_$TAG.log 2>&1; cat $O/ab_mask_$TAG.log; }
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; tail -25 $O/bench_$TAG.err
[ -n "$QUICK" ] && exit 0
timeout 300 python bench.py --impl reference --steps 6 --warmup 1 --ref-step-seconds 4 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err
NCU_COMMON="--clock-control none"
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 1 --component 0 --other-configs none --lines 1500"
timeout 600 ncu --metrics gpu__time_duration.sum $NCU_COMMON -c 120 --csv --log-file $O/launches_$TAG.csv python bench.py --lines 1500 --other-lines 1500 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --component 0 > $O/launch_bench_$TAG.log 2>&1
FP="--metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"
K='regex:k_topo_solve|k_topo_final|k_topo_fused|k_topo_mask|k_geo2rdr'
timeout 300 ncu --set full $FP $NCU_COMMON -k regex:k_fp64_peak -c 1 -f -o /tmp/full_peak_$TAG $B --workload c0c1 > $O/full_peak_$TAG.log 2>&1
python tools/ncu_summary.py /tmp/full_peak_$TAG.ncu-rep $O/ncu_fp64_peak_$TAG.json > $O/ncu_fp64_peak_$TAG.txt 2>&1
# gpurun brings back at most 64 MiB per call: two reports with source come home (the three topo kernels; geo2rdr), every
# other capture is summarised here and dropped
timeout 900 ncu --set full $FP $NCU_COMMON --import-source on -k 'regex:k_topo_solve|k_topo_final|k_topo_mask' -c 3 -f -o $O/full_c2_$TAG $B --workload c2 > $O/full_c2_src_$TAG.log 2>&1
timeout 900 ncu --set full $FP $NCU_COMMON --import-source on -k 'regex:k_geo2rdr' -c 1 -f -o $O/full_geo_$TAG $B --workload c2 > $O/full_geo_src_$TAG.log 2>&1
for W in c2 c0c1 c3; do
  timeout 900 ncu --set full $FP $NCU_COMMON -k "$K" -c 8 -f -o /tmp/full_${W}_$TAG $B --workload $W > $O/full_${W}_$TAG.log 2>&1
  python tools/ncu_summary.py /tmp/full_${W}_$TAG.ncu-rep $O/ncu_full_${W}_$TAG.json > $O/ncu_full_${W}_$TAG.txt 2>&1
done
ls -la $O/*.ncu-rep; du -sh $O
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
