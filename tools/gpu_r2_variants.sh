#!/bin/bash
# A/B of library builds on one box: every csrc/build/variants/*.so through tools/time_bwd.py, twice, interleaved
mkdir -p gpurun_out
: > gpurun_out/variants.log
for rep in $(seq 1 ${REPS:-2}); do
  for v in palette_and_histo_gan_b200/csrc/build/variants/*.so; do
    PALHIST_LIB=$PWD/$v timeout 200 python tools/time_bwd.py >> gpurun_out/variants.log 2>&1 || echo "$v FAILED rc=$?" >> gpurun_out/variants.log
  done
done
cat gpurun_out/variants.log
