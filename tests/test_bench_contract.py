"""bench.py's output contract, on the CPU-runnable arm (`--impl reference` times the oracle port on host cores)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_exactly_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-photons", "4000",
                        "--workload", "c2_bunny_chess"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "photons_per_s" and d["unit"] == "photons/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0
    assert d["config"]["workload"] == "c2_bunny_chess"
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_uses_every_host_core_under_torchrun_and_times_the_reference_as_shipped():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the CPU arm must still use the cores the process may run on (affinity
    mask), and carry the unmodified reference's own photon loop (oracle/_ref, main.cpp:222-249) beside the port."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-photons", "4000",
                        "--shipped-photons", "2000", "--workload", "c2_bunny_chess"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][-1])
    cb = d["cpu_baseline"]
    assert cb["cores"] == len(os.sched_getaffinity(0)) and cb["omp_num_threads_env"] == "1"
    assert cb["kind"] == "port" and cb["deposits_per_hit"] > 0
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libcgref.so")):
        sh = cb["as_shipped"]
        assert sh["kind"] == "reference" and sh["value"] > 0 and sh["cores"] == cb["cores"] and sh["hitpoints"] > 0


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0           # no CPU fallback on the product path
    assert not r.stdout.strip()        # and no bench line
