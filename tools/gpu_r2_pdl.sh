#!/bin/bash
# programmatic dependent launch on / off: parity tests, then the step at 4096 and at the 512-pair shard of the 8-GPU run
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hist.py tests/test_gpu_fuzz.py -m gpu -q -x > gpurun_out/r2pdl_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2pdl_pytest.log
for pdl in 1 0 1 0; do for b in 4096 512; do
  PH_PDL=$pdl PH_BENCH_BATCH=$b timeout 300 python bench.py --steps 30 --warmup 5 --no-generator-step --no-scale-sweep --no-cpu-baseline > gpurun_out/r2pdl_$pdl_$b.json 2> gpurun_out/r2pdl.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2pdl_$pdl_$b.json")); print("PDL=$pdl batch=$b", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["phase_ms"].items()}, d["loss"])
PY
done; done
