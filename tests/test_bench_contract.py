"""bench.py's contract with the driver, as far as it can be exercised without a GPU: the reference arm (the CPU oracle on
the host cores) prints ONE JSON line with the agreed keys, sets its thread count itself (torchrun exports
OMP_NUM_THREADS=1), and the CUDA arm fails loudly on a box without a GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(args, env_extra=None, timeout=300):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=env, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line_and_sets_its_threads():
    out = run_bench(["--impl", "reference", "--config", "cfg1", "--steps", "2", "--warmup", "1"], {"OMP_NUM_THREADS": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["unit"] == "ray-samples/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("cfg1")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] > 0 and "sample" in cb
    usable = len(os.sched_getaffinity(0))
    assert cb["cores"] >= 1 and (cb["cores"] > 1 or usable == 1), "the oracle must not inherit OMP_NUM_THREADS=1"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] > 0
    assert d["gpu_launches"] == 0


def test_cuda_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return  # on a GPU box the arm runs (the driver does that)
    out = run_bench(["--config", "cfg1", "--steps", "1", "--warmup", "0", "--no-cpu-baseline"])
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.strip().startswith("{")], "no bench line may be printed without a GPU"
