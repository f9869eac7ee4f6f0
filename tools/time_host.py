import sys, os, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import bench
from palette_and_histo_gan_b200 import hostapi
B = 4096
real_np, fake_np, real_u8 = bench.make_hist_inputs(B, 47, with_u8=True)
rf = torch.from_numpy(real_np).pin_memory(); ff = torch.from_numpy(fake_np).pin_memory(); ru = torch.from_numpy(real_u8).pin_memory()
gd = torch.empty((B, 64, 64, 4), dtype=torch.float32, device="cuda:0")
ctx = hostapi.HostContext(0)
for name, r in (("float real", rf), ("u8 real", ru), ("float real", rf), ("u8 real", ru)):
    for k in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s = hostapi.histogram_loss_begin(r, ff, ctx=ctx)
        t1 = time.perf_counter()
        l, _ = hostapi.histogram_loss_finish(s, B, None, out_grad_device=gd, ctx=ctx)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"{name} iter {k}: begin {1e3*(t1-t0):.2f} ms finish {1e3*(t2-t1):.2f} ms loss {l:.6f}", flush=True)
