"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`, the CPU oracle on the host
cores) prints exactly one JSON line with the keys the driver reads, alone and under torchrun (rank 0 prints, the other rank exits 0 silently);
the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
            "cpu_baseline", "e2e"]


def _check_line(out, n_gpus):
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    r = json.loads(lines[0])
    for k in REQUIRED:
        assert k in r, k
    assert r["impl"] == "reference" and r["unit"] == "Mpaths/s" and r["n_gpus"] == n_gpus and r["steps"] == 1 and r["warmup"] >= 3
    assert r["value"] > 0 and r["higher_is_better"] is True and r["vs_baseline"] is None and r["dtype"] == "f32"
    from oracle import pyrefglsl
    # the reference's own shaders where oracle/_ref/libref_glsl_*.so exist (built from /root/reference), else the oracle port
    assert r["cpu_baseline"]["kind"] == ("reference" if pyrefglsl.available() else "port")
    assert r["cpu_baseline"]["cores"] >= 1 and r["cpu_baseline"]["value"] == r["value"]
    assert r["e2e"] == {"value": r["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in r["config"]
    return r


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3", "--texture-size", "16"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    _check_line(p.stdout, 1)


def test_reference_arm_under_torchrun_rank0_only():
    env = dict(os.environ, OMP_NUM_THREADS="1")  # what torchrun exports to every rank
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29731",
                        os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3", "--texture-size", "16"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    r = _check_line(p.stdout, 2)
    assert r["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))  # all host cores, not torchrun's OMP_NUM_THREADS=1


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present: the bench would run")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--texture-size", "16"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode != 0
    assert "no CPU path" in (p.stderr + p.stdout)
    assert not [l for l in p.stdout.splitlines() if l.startswith("{")]
