"""Device-side cross-check of the tensor-core engine against the CUDA-core engine and the fp64 oracle."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from palette_and_histo_gan_b200 import histogram as H
from oracle import histogram_oracle as ho
dev = torch.device("cuda:0")
torch.manual_seed(0)
def check(shape, oracle=True, method="inverse-quadratic", sigma=0.02):
    x = torch.tanh(torch.randn(*shape, device=dev))
    a = H.calculate_rgbuv_histogram(x, method=method, sigma=sigma, impl="simt")
    torch.cuda.synchronize()
    b = H.calculate_rgbuv_histogram(x, method=method, sigma=sigma, impl="tc")
    torch.cuda.synchronize()
    an, bn = a.cpu().numpy(), b.cpu().numpy()
    msg = f"{shape} {method}: tc-vs-simt relL2 {ho.rel_l2(bn, an):.3e} relmax {ho.rel_max(bn, an):.3e}"
    if oracle:
        ref, _ = ho.rgbuv_histogram_f64(x.cpu().numpy(), method=method, sigma=sigma)
        msg += f" | tc-vs-f64 {ho.rel_l2(bn, ref):.3e}/{ho.rel_max(bn, ref):.3e} simt-vs-f64 {ho.rel_l2(an, ref):.3e}"
    print(msg, flush=True)
check((2, 32, 32, 4))
check((5, 64, 64, 4))
check((3, 20, 12, 4))
check((2, 16, 16, 3))
check((2, 32, 32, 4), method="RBF", sigma=0.5)
check((300, 64, 64, 4), oracle=False)
check((1, 256, 256, 4), oracle=False)
# timing
x = torch.tanh(torch.randn(4096, 64, 64, 4, device=dev))
for impl in ("simt", "tc"):
    for _ in range(2): H.calculate_rgbuv_histogram(x, impl=impl)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): H.calculate_rgbuv_histogram(x, impl=impl)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"fwd {impl}: {dt*1e3:.2f} ms / 4096 images -> {4096/dt:.0f} img/s, {6*64*64*4096*4096/dt/1e12:.1f} TFLOP/s", flush=True)
