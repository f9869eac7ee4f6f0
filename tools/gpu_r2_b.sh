#!/bin/bash
# Round-2 GPU pass B (1 GPU): tests, bench, the A-tile-store experiment, palette capture.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -8 gpurun_out/r2b_pytest.log
python bench.py --steps 20 --warmup 5 --no-generator-step --no-scale-sweep --no-cpu-baseline > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench rc=$?"
PH_FWD_EXP=1 python bench.py --steps 10 --warmup 3 --no-generator-step --no-scale-sweep --no-cpu-baseline > gpurun_out/r2b_bench_exp1.json 2> gpurun_out/r2b_bench_exp1.err; echo "exp1 rc=$?"
python tools/prof_palette.py > gpurun_out/r2b_pal.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"extract_palette" -s 2 -c 1 -o gpurun_out/r2b_prof_palette -f python tools/prof_palette.py > gpurun_out/r2b_ncu_pal.log 2>&1
