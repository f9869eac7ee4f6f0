#!/bin/bash
# quick check of a backward-kernel change: accuracy against both references, phase times (twice), histogram tests
mkdir -p gpurun_out
{
  timeout 150 python tools/tc_check_bwd.py 2>&1 | tail -9 | cut -c1-230
  timeout 120 python tools/time_bwd.py 2>&1 | tail -1
  timeout 120 python tools/time_bwd.py 2>&1 | tail -1
  if [ -n "$RUN_TESTS" ]; then timeout 900 python -m pytest tests/test_gpu_hist.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -3; fi
} > gpurun_out/quickbwd.log 2>&1
cat gpurun_out/quickbwd.log
