"""bench.py's output contract, on the arm that runs without a GPU: `--impl reference`
prints exactly ONE line on stdout, a JSON object with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    from oracle import cpu_checkers as cc
    assert cc.available("ref") or cc.available("orc")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--events", "20000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"].startswith("MH steps/sec") and d["unit"] == "steps/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["steps"] == 1 and d["warmup"] == 0 and d["data"] == "synthetic" and d["dtype"] == "f64"


def test_reference_arm_of_the_other_configs():
    """`bench.py --impl reference --config c1|c3|c4`: one JSON line each with the contract's keys."""
    for config, metric in (("c1", "MH steps/sec"), ("c3", "MH steps/sec"), ("c4", "HMC steps/sec")):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", config,
                            "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=900, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-2000:]
        lines = [l for l in r.stdout.splitlines() if l.strip()]
        assert len(lines) == 1, r.stdout
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["metric"].startswith(metric) and d["value"] > 0
        assert d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["d2h_bytes_per_step"] == 0
        assert "workload" in d["config"] and d["vs_baseline"] is None


def test_bench_refuses_to_run_the_gpu_arm_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    for extra in ([], ["--config", "c3"]):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"] + extra,
                           capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
