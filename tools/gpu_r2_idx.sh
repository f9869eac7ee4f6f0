#!/bin/bash
# finish_item with three base pointers instead of per-element divisions: histogram tests, A/B against the previous
# object (results must be bit-identical: same loss digits, same gradient checksum), kernel durations under ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hist.py tests/test_gpu_fuzz.py -m gpu -q -x > gpurun_out/idx_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/idx_pytest.log
REPS=3 bash tools/gpu_r2_variants.sh > /dev/null 2>&1; cat gpurun_out/variants.log
for v in a_old b_new; do
PALHIST_LIB=$PWD/palette_and_histo_gan_b200/csrc/build/variants/$v.so timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hist_fwd_tca|hist_fwd_sym" -s 4 -c 4 python tools/step_only.py 2>&1 | grep -E "gpu__time" | tr '\n' ' '; echo " <- $v"
PALHIST_LIB=$PWD/palette_and_histo_gan_b200/csrc/build/variants/$v.so timeout 200 python bench.py --steps 20 --warmup 5 --no-generator-step --no-scale-sweep --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['roofline']['phase_ms'], d['loss'], d['grad0_checksum'])"
done
