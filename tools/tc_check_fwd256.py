"""Dedicated 256-bin forward kernel (hist_tc_fwd256.cu) against the CUDA-core engine and the float64 oracle, and
forward-only timing at the cfgE shape (256 x 256 pixels, 256 bins).  PH_FWD256=0 selects the block path."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from palette_and_histo_gan_b200 import histogram as H
from oracle import histogram_oracle as ho
from tests.conftest import sprite_like_batch
dev = torch.device("cuda:0")
torch.manual_seed(0)
for shape, oracle in [((3, 24, 24, 4), True), ((2, 40, 40, 3), True), ((150, 32, 32, 4), False), ((5, 128, 128, 4), False)]:
    fake = torch.tanh(torch.randn(*shape, device=dev))
    hs = H.calculate_rgbuv_histogram(fake, size=256, impl="simt")
    ht = H.calculate_rgbuv_histogram(fake, size=256, impl="tc")
    msg = f"{shape}: hist tc-vs-simt {ho.rel_l2(ht.cpu().numpy(), hs.cpu().numpy()):.2e} sum-1 {float((ht.sum((1,2,3))-1).abs().max()):.1e}"
    if oracle:
        ref = ho.rgbuv_histogram_f64(fake.cpu().numpy(), size=256)[0]
        msg += f" | tc-vs-f64 {ho.rel_l2(ht.cpu().numpy(), ref):.2e} simt-vs-f64 {ho.rel_l2(hs.cpu().numpy(), ref):.2e}"
    print(msg, flush=True)
# de-duplicated sprites (the real side of the loss) through histogram_loss
rng = np.random.default_rng(3)
spr = torch.from_numpy(sprite_like_batch(rng, 6, hw=64).astype(np.float32) / 127.5 - 1).to(dev)
fk = torch.tanh(torch.randn(6, 64, 64, 4, device=dev))
for impl in ("simt", "tc"):
    f = fk.clone().requires_grad_(True)
    l = H.histogram_loss(spr, f, size=256, impl=impl); l.backward()
    print(impl, "sprite loss", float(l), "grad norm", float(f.grad.norm()), flush=True)
B = int(os.environ.get("PH_E_BATCH", "148"))
for b in (B, 20):
    fake = torch.tanh(torch.randn(b, 256, 256, 4, device=dev))
    H.calculate_rgbuv_histogram(fake, size=256, impl="tc"); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): h = H.calculate_rgbuv_histogram(fake, size=256, impl="tc")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"cfgE forward, batch {b}: {ms:.2f} ms -> {6*256*256*65536*b/ms/1e9:.1f} TFLOP/s algorithmic", flush=True)
    if b == 20:
        hs = H.calculate_rgbuv_histogram(fake[:2], size=256, impl="simt")
        print("sliced images tc-vs-simt", ho.rel_l2(h[:2].cpu().numpy(), hs.cpu().numpy()), flush=True)
