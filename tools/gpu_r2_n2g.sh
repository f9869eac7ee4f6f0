#!/bin/bash
# 2 GPUs: set-up of the peer-memory communicator after the collective-safe rewrite of _comm.py (two-process parity tests,
# peer and NCCL variants; driver-style 2-GPU bench line)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_hist.py -m gpu -q -k "two_rank or single_rank" > gpurun_out/r2g_n2_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2g_n2_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 2 --steps 20 --warmup 5 --no-generator-step --no-scale-sweep > gpurun_out/r2g_n2_bench.json 2> gpurun_out/r2g_n2_bench.err; echo "bench2 rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2g_n2_bench.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["phase_ms"], d["loss"], d["grad0_checksum"], d["e2e"]["value"], d["config"]["parallelism"])
PY
