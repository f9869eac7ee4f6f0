#!/bin/bash
# Round-2 GPU pass on 2 GPUs: the 2-rank parity tests (peer-memory all-reduce and NCCL) and a short 2-GPU bench line.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_n2_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_hist.py -m gpu -q -k "two_rank or single_rank" > gpurun_out/r2_n2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_n2_pytest.log
tail -15 gpurun_out/r2_n2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-generator-step --no-scale-sweep > gpurun_out/r2_n2_bench.json 2> gpurun_out/r2_n2_bench.err; echo "bench2 rc=$?"
PH_COLLECTIVE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --no-generator-step --no-scale-sweep > gpurun_out/r2_n2_bench_nccl.json 2> gpurun_out/r2_n2_bench_nccl.err; echo "bench2 nccl rc=$?"
tail -3 gpurun_out/r2_n2_bench.err
