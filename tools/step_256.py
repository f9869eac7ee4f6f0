"""One device-resident loss step at the cfgE shape (256 x 256 pixels, 256 bins) on the dedicated 256-bin kernels:
the command the ncu capture profiles/r1_prof_hist256_raw.csv is taken from."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from palette_and_histo_gan_b200 import histogram as H
dev = torch.device("cuda:0")
B = int(os.environ.get("PH_E_BATCH", "148"))
torch.manual_seed(0)
real = torch.tanh(torch.randn(B, 256, 256, 4, device=dev)); fake = torch.tanh(torch.randn(B, 256, 256, 4, device=dev))
for _ in range(2):
    f = fake.clone().requires_grad_(True)
    H.histogram_loss(real, f, size=256, impl="tc").backward()
torch.cuda.synchronize()
print("ok")
