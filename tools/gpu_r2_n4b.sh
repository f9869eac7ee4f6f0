#!/bin/bash
mkdir -p gpurun_out
for c in 0 148; do
  if [ $c = 0 ]; then E=""; else E="PH_HOST_CHUNK=$c"; fi
  env $E timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2958$((c % 10)) tools/e2e_shard.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tee -a gpurun_out/r2_n4_e2e_shard.txt
done
