#!/bin/bash
# Round-2 GPU pass A (1 GPU): tests, bench lines, launch list and full captures of the hot kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench rc=$?"
PH_BENCH_BATCH=512 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-generator-step --no-scale-sweep > gpurun_out/r2a_bench_b512.json 2> gpurun_out/r2a_bench_b512.err; echo "bench512 rc=$?"
python tools/step_only.py > gpurun_out/r2a_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_step.csv python tools/step_only.py > gpurun_out/r2a_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hist_fwd_tc|hist_bwd_tc" -s 4 -c 2 -o gpurun_out/r2a_prof_hist -f python tools/step_only.py > gpurun_out/r2a_ncu_hist.log 2>&1
python tools/prof_palette.py > gpurun_out/r2a_pal.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"extract_palette" -s 2 -c 1 -o gpurun_out/r2a_prof_palette -f python tools/prof_palette.py > gpurun_out/r2a_ncu_pal.log 2>&1
ls -la gpurun_out | tail -15
