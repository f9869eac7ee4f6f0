"""Accuracy at large pixel counts: tcgen05 and CUDA-core engines against the float64 oracle (one image)."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from palette_and_histo_gan_b200 import histogram as H
from oracle import histogram_oracle as ho
dev = torch.device("cuda:0")
torch.manual_seed(1)
for side, bins in ((128, 256), (256, 256), (256, 64)):
    real = torch.tanh(torch.randn(1, side, side, 4, device=dev)); fake = torch.tanh(torch.randn(1, side, side, 4, device=dev))
    ref = ho.hist_loss_and_grad_f64(real.cpu().numpy(), fake.cpu().numpy(), size=bins)
    for impl in ("simt", "tc"):
        f = fake.clone().requires_grad_(True)
        loss = H.histogram_loss(real, f, size=bins, impl=impl); loss.backward()
        h = H.calculate_rgbuv_histogram(fake, size=bins, impl=impl).cpu().numpy()
        print(f"{side}x{side} bins {bins} {impl}: hist {ho.rel_l2(h, ref['hist_fake']):.2e} loss {abs(float(loss.detach())-ref['loss'])/ref['loss']:.1e} grad {ho.rel_l2(f.grad.cpu().numpy(), ref['grad']):.2e}", flush=True)
from tests.conftest import sprite_like_batch
rng = np.random.default_rng(9)
for side, bins in ((256, 256), (256, 64)):
    real = torch.from_numpy(sprite_like_batch(rng, 1, hw=side).astype(np.float32) / np.float32(127.5) - 1).to(dev)
    fake = torch.tanh(torch.randn(1, side, side, 4, device=dev))
    ref = ho.hist_loss_and_grad_f64(real.cpu().numpy(), fake.cpu().numpy(), size=bins)
    for impl in ("simt", "tc"):
        f = fake.clone().requires_grad_(True)
        loss = H.histogram_loss(real, f, size=bins, impl=impl); loss.backward()
        print(f"sprite real {side}x{side} bins {bins} {impl}: loss {abs(float(loss.detach())-ref['loss'])/ref['loss']:.1e} grad {ho.rel_l2(f.grad.cpu().numpy(), ref['grad']):.2e}", flush=True)
