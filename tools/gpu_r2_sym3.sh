#!/bin/bash
# mirrored-tile forward: all histogram GPU tests, the bench line, launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_hist.py tests/test_gpu_fuzz.py tests/test_gpu_generator_step.py -m gpu -q -x > gpurun_out/sym_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/sym_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-generator-step --no-scale-sweep --no-cpu-baseline > gpurun_out/sym_bench.json 2> gpurun_out/sym_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/sym_bench.json")); print(d["ms_per_step"], d["value"], d["roofline"].get("phase_ms"), d["loss"], d["grad0_checksum"], d.get("e2e",{}).get("value"))
PY
