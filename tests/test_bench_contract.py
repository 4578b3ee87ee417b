"""CPU: the driver's contract for bench.py's reference arm (`--impl reference`): one JSON line with the bench's metric,
unit and config, `impl: reference`, a cpu_baseline describing the run and an e2e block; under torchrun only rank 0 runs.
(The default arm needs a GPU: it is exercised by the driver and by `tests -m gpu`.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], capture_output=True, text=True,
                          env=env, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run({"OMP_NUM_THREADS": "1"}, "--steps", "2", "--warmup", "1", "--ref-sample-ecs", "20000", "--gpus", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "VI throughput (EC-iterations/s)" and d["unit"] == "EC-iter/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and abs(d["value"] - d["cpu_baseline"]["value"]) < 1e-9 * d["value"] and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["sample"]
    # every core this process may run on, NOT the OMP_NUM_THREADS=1 torchrun hands its children
    assert cb["cores"] == max(1, len(os.sched_getaffinity(0)))
    cfg = d["config"]
    assert "workload" in cfg and "model" not in cfg and cfg["ecs_per_gpu"] == 12_500_000 and cfg["n_groups"] == 2000
    assert d["extrapolated"] is False and d["vs_baseline"] is None


def test_reference_arm_runs_on_rank_zero_only():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--steps", "1", "--warmup", "0", "--ref-sample-ecs", "20000", "--gpus", "2")
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
