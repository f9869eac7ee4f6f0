"""Kernel-only timing of augment_pair_kernel (4096 RGBA pairs of 64x64, draws resident on the device)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from palette_and_histo_gan_b200 import dataset_utils as D
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(47)
a = (torch.rand(4096, 64, 64, 4, generator=g) * 255).round().to(dev)
b = a.flip(0).contiguous()
delta = (torch.rand(4096, generator=g) - 0.5).to(dev)
tr = D._draw_translations(4096, 64, 64, g).to(dev)
on = (torch.rand(4096, generator=g) < 0.8).to(dev)
for name, kw in (("hue+translate+normalize", dict(hue_delta=delta, translations=tr, apply=on, should_normalize=True)),):
    for _ in range(3):
        D.augment_two(a, b, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        D.augment_two(a, b, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(name, f"{ms:.3f} ms", f"{2 * a.numel() * 4 * 2 / ms / 1e6:.0f} GB/s")
