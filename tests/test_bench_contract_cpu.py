"""bench.py contract on the CPU side: the reference arm (the oracle's C/OpenMP port timed on the host cores) prints
exactly one JSON line with the agreed keys; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.strip()]


def test_reference_arm_line():
    lines = _run({"OMP_NUM_THREADS": "1"})          # what torchrun exports; the arm must still use every core
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "NFFT+adjoint nonuniform pts/sec" and d["unit"] == "pts/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and abs(d["value"] - 2 * 2 ** 21 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    assert d["config"]["workload"].startswith("C2")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"]
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": "pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
