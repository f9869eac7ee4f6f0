#!/bin/bash
# backward prologue staged in shared memory (40 instead of 128 registers): histogram tests, cfgC bench line (loss and
# gradient checksum must not move), kernel durations under ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hist.py tests/test_gpu_fuzz.py -m gpu -q -x > gpurun_out/prep_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/prep_pytest.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-generator-step --no-scale-sweep --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['phase_ms'], d['loss'], d['grad0_checksum'], d['e2e']['value'])"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hist_bwd_prep|hist_bwd_tc" -s 4 -c 4 python tools/step_only.py 2>&1 | grep -E "gpu__time" | tr '\n' ' '; echo
PH_BENCH_BATCH=512 timeout 120 python tools/time_bwd.py
