#!/bin/bash
# Round-2 GPU pass C (1 GPU): palette tests + bench palette numbers + capture of the new extraction kernel.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_palette.py -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
python bench.py --steps 10 --warmup 3 --no-generator-step --no-scale-sweep --no-cpu-baseline > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; echo "bench rc=$?"
python tools/prof_palette.py > gpurun_out/r2c_pal.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"extract_palette" -s 2 -c 1 -o gpurun_out/r2c_prof_palette -f python tools/prof_palette.py > gpurun_out/r2c_ncu_pal.log 2>&1
