"""bench.py's output contract where it can run without a GPU: the reference arm falls back to the CPU oracle port
(the reference itself has no CPU path), and the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(*args, **env):
    e = dict(os.environ, NBODY_BENCH_CPU_BUDGET_S="1.5", **env)
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=300, env=e)


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--n", "16384", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["metric"] == "pairwise_interactions_per_sec" and out["unit"] == "interactions/s"
    assert out["higher_is_better"] is True and out["vs_baseline"] is None and out["n_gpus"] == 1
    assert out["value"] > 0 and out["e2e"]["value"] == out["value"]
    assert out["e2e"]["h2d_bytes_per_step"] == 0 and out["e2e"]["d2h_bytes_per_step"] == 0
    cb = out["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] > 0 and "sample" in cb
    assert "workload" in out["config"]
    assert out["loads_product_library"] is False, "the reference arm must not map the product's library"


def test_reference_arm_takes_every_baseline_config():
    for name in ("shipped", "disc16k", "cluster"):
        r = _run("--impl", "reference", "--config", name, "--n", "4096", "--batch", "1", "--steps", "1", "--warmup", "0")
        assert r.returncode == 0, r.stderr[-2000:]
        out = json.loads(r.stdout.strip().splitlines()[-1])
        assert name in out["config"]["workload"] and out["value"] > 0


def test_reference_arm_other_ranks_stay_silent():
    r = _run("--impl", "reference", "--n", "16384", "--steps", "1", "--warmup", "0", RANK="1", WORLD_SIZE="2")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    if _has_gpu():
        return                                    # on a GPU box the arm runs for real (driver, bench)
    r = _run("--n", "16384", "--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CPU fallback" in r.stderr and r.stdout.strip() == ""
