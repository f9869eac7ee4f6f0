"""Device-side cross-check of the tensor-core backward against the CUDA-core backward and the fp64 oracle."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from palette_and_histo_gan_b200 import histogram as H
from oracle import histogram_oracle as ho
dev = torch.device("cuda:0")
torch.manual_seed(0)
def grads(real, fake, impl, method="inverse-quadratic", sigma=0.02):
    f = fake.clone().requires_grad_(True)
    loss = H.histogram_loss(real, f, method=method, sigma=sigma, impl=impl)
    loss.backward()
    torch.cuda.synchronize()
    return float(loss.detach()), f.grad
def check(shape, oracle=True, method="inverse-quadratic", sigma=0.02):
    real = torch.tanh(torch.randn(*shape, device=dev)); fake = torch.tanh(torch.randn(*shape, device=dev))
    ls, gs = grads(real, fake, "simt", method, sigma)
    lt, gt = grads(real, fake, "tc", method, sigma)
    msg = f"{shape} {method}: loss simt {ls:.8f} tc {lt:.8f} | grad tc-vs-simt relL2 {ho.rel_l2(gt.cpu().numpy(), gs.cpu().numpy()):.3e} relmax {ho.rel_max(gt.cpu().numpy(), gs.cpu().numpy()):.3e}"
    if oracle:
        ref = ho.hist_loss_and_grad_f64(real.cpu().numpy(), fake.cpu().numpy(), method=method, sigma=sigma)
        msg += f" | tc-vs-f64 {ho.rel_l2(gt.cpu().numpy(), ref['grad']):.3e} simt-vs-f64 {ho.rel_l2(gs.cpu().numpy(), ref['grad']):.3e} loss rel {abs(lt-ref['loss'])/ref['loss']:.2e}"
    print(msg, flush=True)
check((2, 32, 32, 4))
check((3, 20, 12, 4))
check((2, 16, 16, 3))
check((5, 64, 64, 4))
check((2, 32, 32, 4), method="RBF", sigma=0.5)
check((300, 64, 64, 4), oracle=False)
check((2, 128, 128, 4), oracle=False)
B = 4096
real = torch.tanh(torch.randn(B, 64, 64, 4, device=dev)); fake = torch.tanh(torch.randn(B, 64, 64, 4, device=dev))
for impl in ("simt", "tc"):
    for _ in range(2): grads(real, fake, impl)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): grads(real, fake, impl)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"fwd+bwd step {impl}: {dt*1e3:.2f} ms / {B} pairs -> {B/dt:.0f} pairs/s, {24*64*64*4096*B/dt/1e12:.1f} TFLOP/s", flush=True)
