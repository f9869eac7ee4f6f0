#!/bin/bash
# final scaling lines of the round (mirrored-tile forward): N = 8, 4, 2 on one 8-GPU box, peer-memory sum over ranks
mkdir -p gpurun_out
run() {  # name, nproc, port
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $2 --steps 30 --warmup 5 --no-generator-step --no-scale-sweep > gpurun_out/$1.json 2> gpurun_out/$1.err; echo "$1 rc=$?"
}
run r2d_n8_bench 8 29551
run r2d_n4_bench 4 29553
run r2d_n2_bench 2 29554
timeout 200 python bench.py --steps 30 --warmup 5 --no-generator-step --no-scale-sweep --no-cpu-baseline > gpurun_out/r2d_n1_bench.json 2> gpurun_out/r2d_n1_bench.err; echo "n1 rc=$?"
python - <<PY
import json
for f in ("r2d_n1_bench","r2d_n2_bench","r2d_n4_bench","r2d_n8_bench"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d["roofline"]["phase_ms"], d["loss"], d["grad0_checksum"], d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
