import faulthandler, sys, os
faulthandler.enable()
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import palette_and_histo_gan_b200 as pkg
from palette_and_histo_gan_b200 import _lib, histogram as H
print("loaded", flush=True)
dev = torch.device("cuda:0")
lib = _lib.load()
import ctypes
sm = ctypes.c_int(); print("devinfo", lib.ph_device_info(0, ctypes.byref(sm), None, None), sm.value, flush=True)
print("ws", lib.ph_hist_workspace_bytes(4, 1024, 64, 1), flush=True)
x = torch.tanh(torch.randn(4, 32, 32, 4, device=dev))
dom = H.histogram_domain(64, dev)
print("dom", dom[:3], flush=True)
h, d = H._forward(x, dom, 0, H._sigma_sqr(0.02), 1)
torch.cuda.synchronize(); print("fwd ok", h.sum().item(), d, flush=True)
