"""One device-resident step loop of the cfgC workload (no e2e / palette / CPU legs): the command the ncu
launch list in profiles/ is taken from."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
import bench
from palette_and_histo_gan_b200 import histogram as H
dev = torch.device("cuda:0")
B = int(os.environ.get("PH_BENCH_BATCH", "4096"))
real_np, fake_np = bench.make_hist_inputs(B, 47)
real = torch.from_numpy(real_np).to(dev); fake = torch.from_numpy(fake_np).to(dev).requires_grad_(True)
for _ in range(5):
    fake.grad = None
    H.histogram_loss(real, fake).backward()
torch.cuda.synchronize()
print("ok")
