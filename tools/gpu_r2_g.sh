#!/bin/bash
# Round-2 GPU pass G (1 GPU): forward kernel with the A operand in tensor memory — parity tests, then both variants timed.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hist.py -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -12 gpurun_out/r2g_pytest.log
for v in tmem smem; do
  PH_FWD_A=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-generator-step --no-scale-sweep --no-cpu-baseline > gpurun_out/r2g_bench_$v.json 2> gpurun_out/r2g_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2g_bench_$v.json")); print("$v", d["ms_per_step"], d["roofline"]["phase_ms"], d["loss"])
PY
done
