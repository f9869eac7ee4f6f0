import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
from palette_and_histo_gan_b200 import histogram as H, _lib
dev = torch.device("cuda:0")
def ev(): return torch.cuda.Event(enable_timing=True)
for B in (296, 512, 1024, 4096):
    real = torch.tanh(torch.randn(B, 64, 64, 4, device=dev)); fake = torch.tanh(torch.randn(B, 64, 64, 4, device=dev))
    dom = H.histogram_domain(64, dev); s2 = H._sigma_sqr(0.02)
    for impl in (1, 2):
        hr, _ = H._forward(real, dom, 0, s2, impl); hf, df = H._forward(fake, dom, 0, s2, impl)
        ssum = H._ssum(hr, hf)
        for _ in range(2): H._backward(fake, dom, 0, s2, impl, hf, df, hist_true=hr, ssum=ssum, global_batch=B)
        torch.cuda.synchronize()
        e = [ev() for _ in range(3)]
        e[0].record(); H._forward(fake, dom, 0, s2, impl); e[1].record()
        H._backward(fake, dom, 0, s2, impl, hf, df, hist_true=hr, ssum=ssum, global_batch=B); e[2].record()
        torch.cuda.synchronize()
        print(f"B={B} impl={impl}: fwd {e[0].elapsed_time(e[1]):.3f} ms  bwd {e[1].elapsed_time(e[2]):.3f} ms", flush=True)
B = 4096
real = torch.tanh(torch.randn(B, 64, 64, 4, device=dev)); fake = torch.tanh(torch.randn(B, 64, 64, 4, device=dev)).requires_grad_(True)
for impl in ("simt", "tc"):
    for k in range(5):
        fake.grad = None
        torch.cuda.synchronize(); t0 = time.perf_counter()
        loss = H.histogram_loss(real, fake, impl=impl)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        loss.backward()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"{impl} iter {k}: fwd {1e3*(t1-t0):.2f} ms, bwd {1e3*(t2-t1):.2f} ms", flush=True)
