import sys, os
sys.path.insert(0, os.getcwd())
import torch
from palette_and_histo_gan_b200 import histogram as H
dev = torch.device("cuda:0")
B = 296
real = torch.tanh(torch.randn(B, 64, 64, 4, device=dev)); fake = torch.tanh(torch.randn(B, 64, 64, 4, device=dev))
for _ in range(2):
    f = fake.clone().requires_grad_(True)
    H.histogram_loss(real, f, impl="tc").backward()
torch.cuda.synchronize()
print("ok")
