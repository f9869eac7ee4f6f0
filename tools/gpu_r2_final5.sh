#!/bin/bash
# records of the round's final code (1 GPU): all GPU tests, smoke, driver-style bench line, reference arm,
# ncu launch list of the step and full captures of the two dominant kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/f5_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/f5_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f5_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/f5_smoke.log
timeout 900 python bench.py > gpurun_out/f5_bench_n1.json 2> gpurun_out/f5_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/f5_bench_ref.json 2> gpurun_out/f5_bench_ref.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f5_launches_step.csv python tools/step_only.py > gpurun_out/f5_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"hist_fwd_sym|hist_bwd_tc_kernel" -s 4 -c 2 -o gpurun_out/f5_prof_hist -f python tools/step_only.py > gpurun_out/f5_ncu_hist.log 2>&1; echo "ncu full rc=$?"
