#!/bin/bash
# mirrored-tile forward: GPU tests, bench line, ncu capture of the kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hist.py tests/test_gpu_fuzz.py -m gpu -q > gpurun_out/sym_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/sym_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-generator-step --no-scale-sweep --no-cpu-baseline > gpurun_out/sym_bench.json 2> gpurun_out/sym_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/sym_bench.json")); print(d["ms_per_step"], d["roofline"].get("phase_ms"), d["loss"], d.get("e2e"))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"hist_fwd_sym" -s 3 -c 1 -o gpurun_out/sym_prof -f python tools/step_only.py > gpurun_out/sym_ncu.log 2>&1; echo "ncu rc=$?"
