"""Times the phases of the cfgC loss step (forward of the real images, forward of the fake images + Hellinger sum,
backward) with CUDA events and checks the gradient of a few images against the float64 oracle — the quick A/B of a
kernel change.  PALHIST_LIB selects another build of the library (tools/gpu_r2_variants.sh)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from palette_and_histo_gan_b200 import histogram as H
from oracle import histogram_oracle as ho

dev = torch.device("cuda:0")
B = int(os.environ.get("PH_BENCH_BATCH", "4096"))
real_np, fake_np = bench.make_hist_inputs(B, 47)
real = torch.from_numpy(real_np).to(dev)
fake = torch.from_numpy(fake_np).to(dev).requires_grad_(True)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
fw, bw = [], []
for it in range(25):
    fake.grad = None
    ev[0].record()
    loss = H.histogram_loss(real, fake)
    ev[1].record()
    loss.backward()
    ev[2].record()
    torch.cuda.synchronize()
    if it >= 5:
        fw.append(ev[0].elapsed_time(ev[1])); bw.append(ev[1].elapsed_time(ev[2]))
pick = [0, 1, 777 % B, B - 1]
ssum = (float(loss.detach()) * (2.0 ** 0.5) * B) ** 2   # S of the whole batch (the float32 loss carries it to 1e-7)
ref = ho.hist_loss_and_grad_f64(real_np[pick], fake_np[pick], global_batch=B, global_ssum=ssum)
g, gref = fake.grad[pick].cpu().numpy(), ref["grad"]
print("lib %s  fwd %.3f ms  bwd %.3f ms  step %.3f ms  loss %.10f  grad rel-L2 vs f64 %.2e" % (
    os.path.basename(os.environ.get("PALHIST_LIB", "default")), np.median(fw), np.median(bw), np.median(fw) + np.median(bw),
    float(loss.detach()), ho.rel_l2(g, gref)), flush=True)
