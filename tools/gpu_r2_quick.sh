#!/bin/bash
# quick check of a kernel change: histogram GPU tests + the cfgC bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hist.py tests/test_gpu_fuzz.py -m gpu -q -x > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2q_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-generator-step --no-scale-sweep --no-cpu-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2q_bench.json")); print(d["ms_per_step"], d["roofline"]["phase_ms"], d["loss"], d["grad0_checksum"])
PY
