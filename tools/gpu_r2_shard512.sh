#!/bin/bash
# the 8-GPU shard (512 pairs) on one GPU: phases without any exchange
mkdir -p gpurun_out
PH_BENCH_BATCH=512 timeout 300 python bench.py --steps 50 --warmup 10 --no-generator-step --no-scale-sweep --no-cpu-baseline > gpurun_out/shard512.json 2> gpurun_out/shard512.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/shard512.json")); print(d["ms_per_step"], d["value"], d["roofline"]["phase_ms"])
PY
