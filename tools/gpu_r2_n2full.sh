#!/bin/bash
# the driver's 2-GPU command, all legs (cfgC, e2e, palette, cfgD with DDP, cfgE)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_n2full_bench.json 2> gpurun_out/r2_n2full_bench.err; echo "bench2 rc=$?"
tail -3 gpurun_out/r2_n2full_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_n2full_ref.json 2> gpurun_out/r2_n2full_ref.err; echo "ref2 rc=$?"
timeout 600 python -m pytest tests/test_gpu_hist.py -m gpu -q -k "two_rank" > gpurun_out/r2_n2full_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_n2full_pytest.log
