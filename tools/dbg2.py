import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from palette_and_histo_gan_b200 import histogram as H
from oracle import histogram_oracle as ho
dev = torch.device("cuda:0")
for shape, bins in [((2, 32, 32, 4), 128), ((2, 32, 32, 4), 64), ((1, 16, 16, 4), 256), ((2,32,32,4), 96)]:
    rng = np.random.default_rng(47)
    real = np.tanh(rng.standard_normal(shape)).astype(np.float32)
    fake = np.tanh(rng.standard_normal(shape)).astype(np.float32)
    ref = ho.hist_loss_and_grad_f64(real, fake, size=bins)
    f = torch.from_numpy(fake).to(dev).requires_grad_(True)
    hr = H.calculate_rgbuv_histogram(torch.from_numpy(real).to(dev), size=bins, impl="simt")
    hf = H.calculate_rgbuv_histogram(f, size=bins, impl="simt")
    loss = H.hellinger_loss(hr, hf)
    loss.backward()
    print(shape, bins, "hist", ho.rel_l2(hf.detach().cpu().numpy(), ref["hist_fake"]), ho.rel_max(hf.detach().cpu().numpy(), ref["hist_fake"]),
          "loss", abs(float(loss.detach()) - ref["loss"]) / ref["loss"], "grad", ho.rel_l2(f.grad.cpu().numpy(), ref["grad"]), flush=True)
    # fp32 torch reference error for comparison
    from oracle import torch_port as tp
    l32, g32 = tp.hist_loss_fwd_bwd(torch.from_numpy(real), torch.from_numpy(fake), bins)
    print("   torch-f32 port: loss", abs(float(l32) - ref["loss"]) / ref["loss"], "grad", ho.rel_l2(g32.numpy(), ref["grad"]))
