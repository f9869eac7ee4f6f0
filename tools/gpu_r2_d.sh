#!/bin/bash
# Round-2 GPU pass D (1 GPU): cluster variants of the palette extraction.
mkdir -p gpurun_out
for cl in 1 2 4; do
  PH_PALETTE_CLUSTER=$cl python -m pytest tests/test_gpu_palette.py -m gpu -q -x > gpurun_out/r2d_pytest_cl$cl.log 2>&1; echo "cluster $cl pytest rc=$?"
  tail -3 gpurun_out/r2d_pytest_cl$cl.log
done
python -m pytest tests/test_gpu_palette.py -m gpu -q > gpurun_out/r2d_pytest_auto.log 2>&1; echo "auto pytest rc=$?"
cat > /tmp/pal_time.py <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import bench
from palette_and_histo_gan_b200 import dataset_utils
dev = torch.device("cuda:0")
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)
for batch in (256, 4096):
    src_np, tgt_np = bench.make_palette_inputs(batch, 47)
    for dt in (torch.int32, torch.uint8):
        src, tgt = torch.from_numpy(src_np).to(dev).to(dt), torch.from_numpy(tgt_np).to(dev).to(dt)
        for _ in range(5):
            dataset_utils.load_indexed_images(src, tgt, "grayness", check=False)
        ts = []
        for _ in range(20):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); dataset_utils.load_indexed_images(src, tgt, "grayness", check=False); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts)); npx = 2 * batch * 64 * 64
        print(f"cluster={os.environ.get('PH_PALETTE_CLUSTER','auto')} batch={batch} dtype={dt} median {ms*1e3:.1f} us  {npx/ms/1e6:.1f} Gpix/s", flush=True)
PY
for cl in 1 2 4 0; do PH_PALETTE_CLUSTER=$cl python /tmp/pal_time.py 2>&1 | grep cluster= ; done | tee gpurun_out/r2d_pal_times.txt
PH_PALETTE_CLUSTER=2 ncu --set full --clock-control none --import-source on -k regex:"extract_palette" -s 2 -c 1 -o gpurun_out/r2d_prof_palette_cl2 -f python tools/prof_palette.py > gpurun_out/r2d_ncu_pal.log 2>&1
