"""Dedicated 256-bin kernels (hist_tc_fwd256.cu, hist_tc_bwd256.cu) against the CUDA-core engine and the float64
oracle, and timing at the cfgE shape (256 x 256 pixels, 256 bins).  PH_FWD256=0 / PH_BWD256=0 select the block path."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from palette_and_histo_gan_b200 import histogram as H
from oracle import histogram_oracle as ho
dev = torch.device("cuda:0")
torch.manual_seed(0)
def run(real, fake, impl, bins=256):
    f = fake.clone().requires_grad_(True)
    loss = H.histogram_loss(real, f, size=bins, impl=impl)
    loss.backward(); torch.cuda.synchronize()
    return float(loss.detach()), f.grad
for shape, oracle in [((3, 24, 24, 4), True), ((2, 40, 40, 3), True), ((150, 32, 32, 4), False)]:
    real = torch.tanh(torch.randn(*shape, device=dev)); fake = torch.tanh(torch.randn(*shape, device=dev))
    ls, gs = run(real, fake, "simt"); lt, gt = run(real, fake, "tc")
    msg = f"{shape}: loss simt {ls:.8f} tc {lt:.8f} | grad tc-vs-simt {ho.rel_l2(gt.cpu().numpy(), gs.cpu().numpy()):.2e}"
    if oracle:
        ref = ho.hist_loss_and_grad_f64(real.cpu().numpy(), fake.cpu().numpy(), size=256)
        msg += f" | tc-vs-f64 loss {abs(lt-ref['loss'])/ref['loss']:.1e} grad {ho.rel_l2(gt.cpu().numpy(), ref['grad']):.2e} (simt {ho.rel_l2(gs.cpu().numpy(), ref['grad']):.2e})"
    print(msg, flush=True)
for B in (int(os.environ.get("PH_E_BATCH", "148")), 20):
    real = torch.tanh(torch.randn(B, 256, 256, 4, device=dev)); fake = torch.tanh(torch.randn(B, 256, 256, 4, device=dev))
    run(real, fake, "tc")
    f = fake.clone().requires_grad_(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    loss = H.histogram_loss(real, f, size=256, impl="tc")
    ev[1].record()
    loss.backward()
    ev[2].record(); torch.cuda.synchronize()
    fwd, bwd = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    unit = 256 * 256 * 65536 * B / 1e9
    print(f"cfgE shape, batch {B}: fwd(real)+fwd(fake) {fwd:.2f} ms ({12*unit/fwd:.0f} TFLOP/s), bwd {bwd:.2f} ms ({12*unit/bwd:.0f} TFLOP/s), "
          f"step {fwd+bwd:.2f} ms -> {B/(fwd+bwd)*1e3:.0f} pairs/s, {24*unit/(fwd+bwd):.0f} TFLOP/s", flush=True)
