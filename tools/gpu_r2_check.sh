#!/bin/bash
# Round-2 GPU check (1 GPU): the whole GPU suite, bench at 4096 pairs and at the 512-pair shard of the 8-GPU run, launch list of that shard.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2chk_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2chk_pytest.log
tail -8 gpurun_out/r2chk_pytest.log
python bench.py --steps 20 --warmup 5 --no-generator-step --no-scale-sweep --no-cpu-baseline > gpurun_out/r2chk_bench_n1.json 2> gpurun_out/r2chk_bench_n1.err; echo "bench rc=$?"
PH_BENCH_BATCH=512 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-generator-step --no-scale-sweep > gpurun_out/r2chk_bench_b512.json 2> gpurun_out/r2chk_bench_b512.err; echo "bench512 rc=$?"
PH_BENCH_BATCH=512 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2chk_launches_b512.csv python tools/step_only.py > gpurun_out/r2chk_ncu_launch.log 2>&1
