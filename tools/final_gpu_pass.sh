set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_r01d.log; cat gpurun_out/pytest_gpu_r01d.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; tail -3 gpurun_out/bench_r01d.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01d.json 2> gpurun_out/bench_ref_r01d.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --lines 1500 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launch_bench_r01d.log 2>&1
ncu --set full --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum --clock-control none --import-source on -k regex:"k_topo_solve|k_topo_final|k_topo_mask|k_geo2rdr_poly" -c 4 -f -o gpurun_out/full_r01d python bench.py --lines 1500 --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 1 > gpurun_out/full_bench_r01d.log 2>&1
ls -la gpurun_out/full_r01d.ncu-rep
python __graft_entry__.py smoke 2>&1 | tail -2
