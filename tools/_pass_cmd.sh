O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu_full_r02l.log 2>&1; tail -8 $O/pytest_gpu_full_r02l.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_r02l.json 2> $O/bench_r02l.err; tail -5 $O/bench_r02l.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02l.json')); print(d['value'], d['ms_per_step'], d['breakdown_ms'], d['e2e']['value'], d.get('bench_seconds')); print(json.dumps(d['e2e_component'])[:1800])"
