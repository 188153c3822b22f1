"""N > 1 host logic on CPU (gloo, world_size 2): the azimuth line-block sharding of bench.py covers the swath exactly
once, the timing reduction is a max over ranks, byte counters are summed, and the `--impl reference` arm runs on rank
0 only.  No data-path collective exists on this path (lines are independent), so this is all the multi-rank code."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, {root!r})
import bench
r = bench.Ranks()
a, n = bench.shard(13501, r.rank, r.world)
tot = r.reduce_sum(n)
mx = r.reduce_max(10.0 * (r.rank + 1))
r.barrier()
open(os.path.join({out!r}, "rank%d.json" % r.rank), "w").write(
    json.dumps(dict(rank=r.rank, world=r.world, line0=a, nlines=n, total=tot, max=mx)))
r.close()
'''


def _torchrun(args, timeout=240):
    env = dict(os.environ)
    env["MASTER_ADDR"] = "127.0.0.1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533"] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_line_block_sharding_and_reductions_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=str(tmp_path)))
    p = _torchrun([str(script)])
    assert p.returncode == 0, p.stderr[-2000:]
    rows = [json.loads((tmp_path / f"rank{r}.json").read_text()) for r in (0, 1)]
    assert [r["rank"] for r in rows] == [0, 1] and all(r["world"] == 2 for r in rows)
    assert rows[0]["line0"] == 0 and rows[0]["line0"] + rows[0]["nlines"] == rows[1]["line0"]
    assert rows[1]["line0"] + rows[1]["nlines"] == 13501
    assert all(r["total"] == 13501 and r["max"] == 20.0 for r in rows)


def test_shard_partitions_any_world():
    sys.path.insert(0, ROOT)
    import bench
    for length in (1, 7, 1500, 13500, 60001):
        for world in (1, 2, 3, 4, 8):
            blocks = [bench.shard(length, r, world) for r in range(world)]
            pos = 0
            for a, n in blocks:
                assert a == pos and n >= 0
                pos += n
            assert pos == length
            assert max(n for _, n in blocks) - min(n for _, n in blocks) <= 1


@pytest.mark.timeout(600)
def test_reference_arm_runs_on_rank0_only_world2():
    p = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "c0c1",
                   "--lines", "48", "--ref-step-seconds", "1"], timeout=500)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["metric"] == "topo+geo2rdr Mpixels/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


BENCH_WORKER = r'''
import sys
sys.path.insert(0, {root!r})
import isce2_b200
from tests import fake_capi_for_bench as fake
sys.modules["isce2_b200._capi"] = fake
isce2_b200._capi = fake
import bench
{patch}
sys.argv = ["bench.py", "--gpus", "2", "--steps", "2", "--warmup", "1", "--workload", {workload!r}, "--lines", "37", "--e2e-steps", "2",
            "--other-lines", "29", "--component", "0", "--no-cpu-baseline", "--max-seconds", "300"]
bench.main()
'''


@pytest.mark.timeout(400)
def test_bench_control_flow_world2_with_a_stand_in_library(tmp_path):
    """The whole of bench.run_b200 under torchrun with two ranks (tests/fake_capi_for_bench.py stands in for the CUDA
    library): every rank reaches every barrier / reduction, also when the two end-to-end arms of a rank that does not
    start at line 0 differ in the last bits; rank 0 alone prints the JSON line, with the contract's keys."""
    script = tmp_path / "bench_worker.py"
    script.write_text(BENCH_WORKER.format(root=ROOT, workload="c0c1", patch=""))
    p = _torchrun([str(script)], timeout=300)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = lines[0]
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "e2e_two_calls", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 2 and d["config"]["pixels_per_step"] == 37 * 21000 and d["cpu_baseline"] is None
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] < d["e2e_two_calls"]["h2d_bytes_per_step"]
    v = d["e2e"]["vs_two_calls"]
    assert v["compared"] and v["validity_equal"] and v["max_abs_offset_diff_px"] < 1e-3
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert d["e2e"]["statistic"].startswith("median") and d["e2e"]["d2h_floor"]["ms"] > 0 and 0 < d["e2e"]["frac_of_d2h_floor"]
    assert v["rows_compared_per_rank"] > 0
    # the other BASELINE configs ride in the same line: C3 (NISAR, Legendre, native Doppler) and the 29-orbit stack batch
    oc = d["other_configs"]
    assert set(oc) == {"c3", "c4"} and all("error" not in oc[k] for k in oc), oc
    assert oc["c3"]["config"]["pixels_per_step"] == 29 * 25000 and oc["c3"]["config"]["orbit_method"] == "LEGENDRE"
    assert oc["c3"]["value"] > 0 and oc["c3"]["e2e"]["value"] > 0 and "e2e_two_calls" not in oc["c3"]
    assert oc["c4"]["config"]["jobs_on_rank0"] == 15 and oc["c4"]["value"] > 0


@pytest.mark.timeout(400)
def test_stack_batch_control_flow_world2_with_a_stand_in_library(tmp_path):
    """configs[4] (29 secondary orbits dealt round-robin to the ranks) through the same stand-in: one JSON line, from rank 0."""
    script = tmp_path / "bench_worker_c4.py"
    script.write_text(BENCH_WORKER.format(root=ROOT, workload="c4", patch=""))
    p = _torchrun([str(script)], timeout=300)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and lines[0]["n_gpus"] == 2 and lines[0]["config"]["jobs_on_rank0"] == 15
    assert lines[0]["metric"].startswith("geo2rdr Mpixels/s") and lines[0]["value"] > 0


@pytest.mark.timeout(400)
def test_a_failing_secondary_config_is_reported_not_fatal_world2(tmp_path):
    """A secondary config that fails on one rank only (here: rank 1 cannot build its scene) must neither hang the ranks
    nor cost the main line: its local work runs under a guard, every collective is still reached by every rank."""
    patch = ("import os\n"
             "if os.environ.get('RANK') == '1':\n"
             "    bench.WORKLOADS['c3']['sensor'] = 'no-such-sensor'\n")
    script = tmp_path / "bench_worker_fail.py"
    script.write_text(BENCH_WORKER.format(root=ROOT, workload="c0c1", patch=patch))
    p = _torchrun([str(script)], timeout=300)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and lines[0]["value"] > 0 and lines[0]["e2e"]["value"] > 0
    oc = lines[0]["other_configs"]
    assert "error" in oc["c3"] and "error" not in oc["c4"] and oc["c4"]["value"] > 0
