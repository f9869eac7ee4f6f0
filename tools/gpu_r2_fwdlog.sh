#!/bin/bash
# table-driven float32 log differences in the mirrored forward's pixel pass: accuracy, phase times, histogram tests
mkdir -p gpurun_out
{
  timeout 150 python tools/tc_check_sym.py 2>&1 | tail -12
  timeout 120 python tools/time_bwd.py 2>&1 | tail -1
  timeout 120 python tools/time_bwd.py 2>&1 | tail -1
  timeout 900 python -m pytest tests/test_gpu_hist.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -3
} > gpurun_out/fwdlog.log 2>&1
cat gpurun_out/fwdlog.log
