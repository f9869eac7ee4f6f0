#!/bin/bash
# ncu capture of the forward kernel variants (A in TMEM / in shared memory) and the backward
mkdir -p gpurun_out
python tools/step_only.py > gpurun_out/r2h_step.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"hist_fwd_tca|hist_bwd_tc" -s 4 -c 2 -o gpurun_out/r2h_prof_hist -f python tools/step_only.py > gpurun_out/r2h_ncu_hist.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches_step.csv python tools/step_only.py > gpurun_out/r2h_ncu_launch.log 2>&1
