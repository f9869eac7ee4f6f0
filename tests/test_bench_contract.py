"""The bench.py contract that can be checked without a GPU: the reference arm (`--impl reference`, the torch-CPU port of
the reference timed on the host cores) prints exactly one JSON line with the keys the driver reads, on the same
metric / unit / config as the CUDA arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"].startswith("histogram-loss fwd+bwd images/s") and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["config"]["bins"] == 64 and d["config"]["global_batch"] == 4096 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_bench_declares_the_contract_flags():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in src
    for key in ('"roofline"', '"cpu_baseline"', '"e2e"', '"gpu_launches"', '"clocks"', '"scaling"'):
        assert key in src, key
