import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import bench
from palette_and_histo_gan_b200 import histogram as H
dev = torch.device("cuda:0")
dom = H.histogram_domain(64, dev); s2 = H._sigma_sqr(0.02)
real_np, fake_np = bench.make_hist_inputs(4096, 47)
real = torch.from_numpy(real_np).to(dev); fake = torch.from_numpy(fake_np).to(dev)
for _ in range(2):
    H._forward(real, dom, 0, s2, 2 | 8)      # dedup
    H._forward(fake[:37], dom, 0, s2, 2)     # 37 images sliced by 8
    H._forward(fake[:148], dom, 0, s2, 2)    # 148 whole
    H._forward(fake[:512], dom, 0, s2, 2)
torch.cuda.synchronize(); print("ok")
