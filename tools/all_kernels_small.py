"""Small end-to-end pass over every kernel of the library (a quick functional check on a fresh box)."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from palette_and_histo_gan_b200 import histogram as H, io_utils, dataset_utils, hostapi
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
def loss_step(b, hw, bins, impl, dedup_like=False):
    fake = torch.tanh(torch.randn((b, hw, hw, 4), device=dev, generator=g)).requires_grad_(True)
    if dedup_like:
        real = (torch.randint(0, 4, (b, hw, hw, 4), device=dev, generator=g).float() / 1.5 - 1.0).contiguous()
    else:
        real = torch.tanh(torch.randn((b, hw, hw, 4), device=dev, generator=g))
    l = H.histogram_loss(real, fake, size=bins, impl=impl); l.backward(); torch.cuda.synchronize()
    return float(l.detach())
print("tc sliced   ", loss_step(3, 16, 64, "tc"))
print("tc whole+tail", loss_step(150, 8, 64, "tc", dedup_like=True))
print("tc 128 bins ", loss_step(2, 16, 128, "tc"))
print("simt 48 bins", loss_step(2, 8, 48, "simt"))
up = torch.randn((2, 64, 64, 3), device=dev, generator=g)
x = torch.tanh(torch.randn((2, 16, 16, 4), device=dev, generator=g)).requires_grad_(True)
H.calculate_rgbuv_histogram(x, impl="tc").backward(up); torch.cuda.synchronize()
rng = np.random.default_rng(0)
real = np.tanh(rng.standard_normal((300, 8, 8, 4))).astype(np.float32); fake = np.tanh(rng.standard_normal((300, 8, 8, 4))).astype(np.float32)
print("host pipeline", hostapi.histogram_loss(real, fake)[0])
src = torch.from_numpy(rng.integers(0, 3, (4, 16, 16, 4)).astype(np.int32) * 100).to(dev); tgt = src.flip(1).contiguous()
s_idx, t_idx, pal = dataset_utils.load_indexed_images(src, tgt, "grayness")
oh = io_utils.one_hot(t_idx)
idx2, rgba = io_utils.probabilities_to_indexed(oh, pal)
assert torch.equal(rgba, tgt) and torch.equal(idx2, t_idx)
u8 = torch.from_numpy(rng.integers(0, 256, (2, 8, 8, 4)).astype(np.uint8)).to(dev)
dataset_utils.load_image(u8); torch.cuda.synchronize()
print("ok")
