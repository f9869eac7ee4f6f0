#!/bin/bash
# table-driven float64 log in the pixel terms: accuracy of both engines (tc_check_bwd), phase times, histogram tests
mkdir -p gpurun_out
{
  timeout 150 python tools/tc_check_bwd.py 2>&1 | tail -10
  timeout 120 python tools/time_bwd.py 2>&1 | tail -1
  timeout 120 python tools/time_bwd.py 2>&1 | tail -1
  timeout 900 python -m pytest tests/test_gpu_hist.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -3
} > gpurun_out/log.log 2>&1
cat gpurun_out/log.log
