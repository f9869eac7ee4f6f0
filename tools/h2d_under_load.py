"""Does the pinned H2D copy of one step's inputs slow down while the loss kernels run?  Times 336 MB of uploads in
296-image chunks on a copy stream, alone and with the device-resident loss step looping on the compute stream."""
import sys, os, time, threading
sys.path.insert(0, os.getcwd())
import torch, numpy as np, bench
from palette_and_histo_gan_b200 import histogram as H
dev = torch.device("cuda:0")
real_np, fake_np, real_u8 = bench.make_hist_inputs(4096, 47, with_u8=True)
real_h, fake_h = torch.from_numpy(real_u8).pin_memory(), torch.from_numpy(fake_np).pin_memory()
d_real, d_fake = torch.empty_like(real_h, device=dev), torch.empty_like(fake_h, device=dev)
real = torch.from_numpy(real_np).to(dev); fake = torch.from_numpy(fake_np).to(dev).requires_grad_(True)
copy_stream = torch.cuda.Stream()
def upload():
    with torch.cuda.stream(copy_stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for lo in range(0, 4096, 296):
            hi = min(lo + 296, 4096)
            d_real[lo:hi].copy_(real_h[lo:hi], non_blocking=True); d_fake[lo:hi].copy_(fake_h[lo:hi], non_blocking=True)
        e1.record()
    return e0, e1
for _ in range(2): upload()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = upload(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"H2D alone: {np.median(ts):.2f} ms = {(real_h.numel() + fake_h.numel() * 4) / np.median(ts) / 1e6:.1f} GB/s")
def step():
    fake.grad = None
    H.histogram_loss(real, fake).backward()
for _ in range(3): step()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    for _ in range(3): step()          # ~18 ms of kernels queued on the compute stream
    e0, e1 = upload()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"H2D while the loss kernels run: {np.median(ts):.2f} ms = {(real_h.numel() + fake_h.numel() * 4) / np.median(ts) / 1e6:.1f} GB/s")
