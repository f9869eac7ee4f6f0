#!/bin/bash
# Round-2 pass on 8 GPUs: scaling lines with the peer-memory all-reduce (N = 8, 4) and with NCCL (N = 8)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_n8_topo.txt 2>&1
run() {  # name, nproc, port, extra env
  env $4 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $2 --steps 30 --warmup 5 --no-generator-step --no-scale-sweep > gpurun_out/$1.json 2> gpurun_out/$1.err; echo "$1 rc=$?"
}
run r2_n8_bench 8 29551 PH_COLLECTIVE=peer
run r2_n8_bench_nccl 8 29552 PH_COLLECTIVE=nccl
run r2_n4_bench 4 29553 PH_COLLECTIVE=peer
python - <<PY
import json
for f in ("r2_n8_bench","r2_n8_bench_nccl","r2_n4_bench"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d["roofline"]["phase_ms"], d["loss"], d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
