#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2k_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-generator-step --no-scale-sweep --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2k_bench.json")); print(d["ms_per_step"], d["roofline"]["phase_ms"], d["loss"], d["grad0_checksum"])
PY
