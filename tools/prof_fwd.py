import sys, os
sys.path.insert(0, os.getcwd())
import torch
from palette_and_histo_gan_b200 import histogram as H
dev = torch.device("cuda:0")
x = torch.tanh(torch.randn(1024, 64, 64, 4, device=dev))
for _ in range(3):
    H.calculate_rgbuv_histogram(x, impl="tc")
torch.cuda.synchronize()
print("ok")
