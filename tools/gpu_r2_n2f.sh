#!/bin/bash
# final code on 2 GPUs: the whole GPU suite (the 2-rank parity tests run here) and the driver-style 2-GPU bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hist.py -m gpu -q > gpurun_out/r2f_n2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_n2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2f_n2_bench.json 2> gpurun_out/r2f_n2_bench.err; echo "bench2 rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2f_n2_bench.json")); print(d["value"], d["ms_per_step"], d["roofline"]["phase_ms"], d["loss"], d["grad0_checksum"], d["e2e"]["value"])
PY
