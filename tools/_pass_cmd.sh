O=gpurun_out; mkdir -p $O
cat /sys/kernel/mm/transparent_hugepage/shmem_enabled > $O/box_thp_r02m.txt 2>&1; uname -r >> $O/box_thp_r02m.txt; nproc >> $O/box_thp_r02m.txt
timeout 300 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_n1_paths.py -m gpu -q > $O/pytest_gpu_part_r02m.log 2>&1; tail -4 $O/pytest_gpu_part_r02m.log
SWEEP=file timeout 900 python tools/component_threads_sweep.py 4000 > $O/component_file_sweep_4000lines_r02m.log 2>$O/component_file_sweep_r02m.err; cat $O/component_file_sweep_4000lines_r02m.log
SWEEP=file timeout 900 python tools/component_threads_sweep.py > $O/component_file_sweep_r02m.log 2>>$O/component_file_sweep_r02m.err; cat $O/component_file_sweep_r02m.log
