"""Block-decomposed tensor-core path (bins = 128 / 256) against the CUDA-core engine and the float64 oracle, and
timing at a cfgE-like shape (256 x 256 pixels, 256 bins)."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from palette_and_histo_gan_b200 import histogram as H
from oracle import histogram_oracle as ho
dev = torch.device("cuda:0")
torch.manual_seed(0)
def run(real, fake, impl, bins):
    f = fake.clone().requires_grad_(True)
    loss = H.histogram_loss(real, f, size=bins, impl=impl)
    loss.backward(); torch.cuda.synchronize()
    return float(loss.detach()), f.grad
for shape, bins, oracle in [((2, 16, 16, 4), 128, True), ((3, 24, 24, 4), 256, True), ((150, 32, 32, 4), 128, False), ((4, 64, 64, 4), 192, False)]:
    real = torch.tanh(torch.randn(*shape, device=dev)); fake = torch.tanh(torch.randn(*shape, device=dev))
    ls, gs = run(real, fake, "simt", bins); lt, gt = run(real, fake, "tc", bins)
    hs = H.calculate_rgbuv_histogram(fake, size=bins, impl="simt"); ht = H.calculate_rgbuv_histogram(fake, size=bins, impl="tc")
    msg = f"{shape} bins {bins}: loss simt {ls:.8f} tc {lt:.8f} | hist tc-vs-simt {ho.rel_l2(ht.cpu().numpy(), hs.cpu().numpy()):.2e} | grad tc-vs-simt {ho.rel_l2(gt.cpu().numpy(), gs.cpu().numpy()):.2e}"
    if oracle:
        ref = ho.hist_loss_and_grad_f64(real.cpu().numpy(), fake.cpu().numpy(), size=bins)
        msg += f" | tc-vs-f64 hist {ho.rel_l2(ht.cpu().numpy(), ref['hist_fake']):.2e} loss {abs(lt-ref['loss'])/ref['loss']:.1e} grad {ho.rel_l2(gt.cpu().numpy(), ref['grad']):.2e}"
    print(msg, flush=True)
B = int(os.environ.get("PH_E_BATCH", "16"))
real = torch.tanh(torch.randn(B, 256, 256, 4, device=dev)); fake = torch.tanh(torch.randn(B, 256, 256, 4, device=dev))
for impl in ("tc", "simt"):
    n = 3 if impl == "tc" else 1
    run(real, fake, impl, 256)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): run(real, fake, impl, 256)
    dt = (time.perf_counter() - t0) / n
    print(f"cfgE shape, batch {B}, {impl}: {dt*1e3:.1f} ms per step -> {B/dt:.1f} pairs/s, {24*256*256*65536*B/dt/1e12:.1f} TFLOP/s", flush=True)
