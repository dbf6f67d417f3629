"""bench.py --impl reference runs on the host alone (the CPU arm the driver times beside the GPU arm): the JSON line must
carry the contract's keys with the same metric / unit / config as the GPU arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--epoch-cap", "1", "--ref-min-candidates", "2"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT,
                         env=dict(os.environ, OMP_NUM_THREADS="1"))      # what torch.distributed.run exports
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    assert line["metric"] == "candidate_evals_per_sec" and line["unit"] == "evals/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and line["scaling"] == "strong"
    assert "workload" in line["config"] and "model" not in line["config"]
    cpu = line["cpu_baseline"]
    assert cpu["kind"] in ("port", "reference") and cpu["value"] == line["value"] and cpu["sample"]
    # the launcher's OMP_NUM_THREADS=1 must not shrink the CPU arm (round-1 SCALE bug): all host threads at every N
    assert cpu["cores"] == len(os.sched_getaffinity(0)) and len(cpu["seconds_per_candidate"]) == 2
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
