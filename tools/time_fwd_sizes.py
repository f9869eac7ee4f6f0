import sys, os
sys.path.insert(0, os.getcwd())
import torch
from palette_and_histo_gan_b200 import histogram as H
dev = torch.device("cuda:0")
dom = H.histogram_domain(64, dev); s2 = H._sigma_sqr(0.02)
def ev(): return torch.cuda.Event(enable_timing=True)
for B in (37, 74, 148, 222, 296, 444, 512, 592):
    x = torch.tanh(torch.randn(B, 64, 64, 4, device=dev))
    for _ in range(3): H._forward(x, dom, 0, s2, 2)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(20): H._forward(x, dom, 0, s2, 2)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B}: fwd {e0.elapsed_time(e1)/20*1e3:.1f} us", flush=True)
