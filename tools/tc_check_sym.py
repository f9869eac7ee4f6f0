"""Mirrored-tile forward (hist_fwd_sym_kernel, PH_IMPL_MIRROR) against the float64 oracle, the exact-centre tensor-core
kernel and the CUDA-core engine: whole-image items, sliced tails, 3-channel input, RBF, tiny images; then timing."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from palette_and_histo_gan_b200 import histogram as H
from oracle import histogram_oracle as ho
dev = torch.device("cuda:0")
torch.manual_seed(0)
worst = 0.0
def check(shape, oracle=True, method="inverse-quadratic", sigma=0.02, gen="tanh"):
    global worst
    if gen == "tanh": x = torch.tanh(torch.randn(*shape, device=dev))
    else: x = torch.rand(*shape, device=dev) * 2 - 1
    s = H.calculate_rgbuv_histogram(x, method=method, sigma=sigma, impl="tc", mirror=True)
    torch.cuda.synchronize()
    e = H.calculate_rgbuv_histogram(x, method=method, sigma=sigma, impl="tc", mirror=False)
    torch.cuda.synchronize()
    sn, en = s.cpu().numpy(), e.cpu().numpy()
    sums = float((s.sum((1, 2, 3)) - 1).abs().max())
    msg = f"{shape} {method} {gen}: sym-vs-exact relL2 {ho.rel_l2(sn, en):.3e} relmax {ho.rel_max(sn, en):.3e} |sum-1| {sums:.1e}"
    if oracle:
        ref, _ = ho.rgbuv_histogram_f64(x.cpu().numpy(), method=method, sigma=sigma)
        es = ho.rel_l2(sn, ref); worst = max(worst, es)
        msg += f" | sym-vs-f64 {es:.3e}/{ho.rel_max(sn, ref):.3e}  exact-vs-f64 {ho.rel_l2(en, ref):.3e}"
        # per-channel error (index reversals)
        msg += "  per-channel " + " ".join(f"{ho.rel_l2(sn[..., c], ref[..., c]):.1e}" for c in range(3))
    print(msg, flush=True)
check((2, 32, 32, 4))
check((5, 64, 64, 4))
check((3, 20, 12, 4))
check((2, 16, 16, 3))
check((4, 8, 8, 4))
check((1, 64, 64, 4), gen="uniform")
check((2, 32, 32, 4), method="RBF", sigma=0.5)
check((150, 64, 64, 4), oracle=False)
check((300, 64, 64, 4), oracle=False)
check((1, 256, 256, 4), oracle=False)
x = torch.tanh(torch.randn(160, 64, 64, 4, device=dev))
s = H.calculate_rgbuv_histogram(x, impl="tc", mirror=True); torch.cuda.synchronize()
ref, _ = ho.rgbuv_histogram_f64(x[148:].cpu().numpy())
print("tail images of a 160 batch vs f64:", ho.rel_l2(s[148:].cpu().numpy(), ref), flush=True)
ref, _ = ho.rgbuv_histogram_f64(x[:4].cpu().numpy())
print("whole images of a 160 batch vs f64:", ho.rel_l2(s[:4].cpu().numpy(), ref), flush=True)
s2 = H.calculate_rgbuv_histogram(x, impl="tc", mirror=True); torch.cuda.synchronize()
print("deterministic:", bool(torch.equal(s, s2)))
# loss + gradient through the fused forward
real = torch.tanh(torch.randn(6, 32, 32, 4, device=dev)); fake = torch.tanh(torch.randn(6, 32, 32, 4, device=dev)).requires_grad_(True)
R = ho.hist_loss_and_grad_f64(real.cpu().numpy(), fake.detach().cpu().numpy())
for m in (True, False):
    fake.grad = None
    l = H.histogram_loss(real, fake, mirror=m); l.backward(); torch.cuda.synchronize()
    print(f"loss/grad mirror={m}: loss rel {abs(float(l) - R['loss']) / R['loss']:.2e} grad relL2 {ho.rel_l2(fake.grad.cpu().numpy(), R['grad']):.3e}", flush=True)
print("worst sym-vs-f64", worst)
from palette_and_histo_gan_b200 import _lib
print("async status", _lib.async_status(clear=False))
# timing
x = torch.tanh(torch.randn(4096, 64, 64, 4, device=dev))
for m in (False, True):
    for _ in range(3): H.calculate_rgbuv_histogram(x, impl="tc", mirror=m)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): H.calculate_rgbuv_histogram(x, impl="tc", mirror=m)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"fwd mirror={m}: {dt*1e3:.3f} ms / 4096 images, {6*64*64*4096*4096/dt/1e12:.1f} TFLOP/s algorithmic", flush=True)
x = torch.tanh(torch.randn(512, 64, 64, 4, device=dev))
for m in (False, True):
    for _ in range(3): H.calculate_rgbuv_histogram(x, impl="tc", mirror=m)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): H.calculate_rgbuv_histogram(x, impl="tc", mirror=m)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print(f"fwd mirror={m}: {dt*1e3:.3f} ms / 512 images", flush=True)
