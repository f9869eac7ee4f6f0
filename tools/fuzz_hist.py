"""Randomised cross-check of the tcgen05 engine against the CUDA-core engine (both on the device) over batch sizes
around the work-plan boundaries (SM count, partial waves, slices), odd image sizes, 3- and 4-channel pixels, bin counts
64 / 128 / 256, both bin kernels and a range of sigmas; loss, histograms and gradients must agree to 1e-5 (norm-relative).
A sample of the cases is also checked against the float64 oracle.  `python tools/fuzz_hist.py [cases] [seed]`."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from palette_and_histo_gan_b200 import histogram as H
from oracle import histogram_oracle as ho



def run(cases=60, seed=0, verbose=True):
    """Returns the worst relative differences; raises AssertionError on the first case outside 1e-5."""
    _print = print if verbose else (lambda *a, **k: None)
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(seed)
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    worst = {"loss": 0.0, "hist": 0.0, "grad": 0.0, "oracle_grad": 0.0}
    for k in range(cases):
        bins = int(rng.choice([64, 64, 64, 128, 256]))
        batch = int(rng.choice([1, 2, 3, 7, 40, 147, 148, 149, 150, 200, 295, 296, 297, 300, 333])) if bins == 64 else int(rng.choice([1, 2, 5, 150]))
        h, w = (int(rng.integers(1, 40)), int(rng.integers(1, 40))) if batch < 100 else (int(rng.integers(2, 14)), int(rng.integers(2, 14)))
        ch = int(rng.choice([3, 4]))
        method, sigma = ("inverse-quadratic", float(rng.choice([0.02, 0.02, 0.05, 0.2]))) if rng.random() < 0.8 else ("RBF", float(rng.choice([1.0, 2.0])))  # smaller RBF sigmas underflow far bins to exact zeros: 0/0 in the reference's loss too
        dedup = bool(rng.random() < 0.5)
        if dedup:  # palette-like real images
            real = (torch.randint(0, 5, (batch, h, w, ch), device=dev, generator=g).float() / 2.0 - 1.0).contiguous()
        else:
            real = torch.tanh(torch.randn((batch, h, w, ch), device=dev, generator=g))
        fake = torch.tanh(torch.randn((batch, h, w, ch), device=dev, generator=g))
        out = {}
        for impl in ("simt", "tc"):
            f = fake.clone().requires_grad_(True)
            loss = H.histogram_loss(real, f, size=bins, method=method, sigma=sigma, impl=impl, dedup_real=dedup)
            loss.backward()
            out[impl] = (float(loss.detach()), H.calculate_rgbuv_histogram(fake, size=bins, method=method, sigma=sigma, impl=impl), f.grad)
        tag0 = f"case {k}: bins {bins} batch {batch} {h}x{w}x{ch} {method} sigma {sigma} dedup {dedup}"
        if not (np.isfinite(out["simt"][0]) and bool(torch.isfinite(out["simt"][2]).all()) and out["simt"][0] > 0):
            # e.g. RBF with every bin underflowing to zero for a one-pixel image: 0/0 in the reference as well
            _print("skip " + tag0 + "  (the CUDA-core engine's result is not finite: degenerate input)", flush=True)
            continue
        e_loss = abs(out["tc"][0] - out["simt"][0]) / abs(out["simt"][0])
        e_hist, e_grad = rel(out["tc"][1], out["simt"][1]), rel(out["tc"][2], out["simt"][2])
        tag = f"case {k}: bins {bins} batch {batch} {h}x{w}x{ch} {method} sigma {sigma} dedup {dedup}"
        e_or = None
        if k % 6 == 0 and batch * h * w * bins <= 4e6:
            ref = ho.hist_loss_and_grad_f64(real.cpu().numpy(), fake.cpu().numpy(), size=bins, method=method, sigma=sigma)
            e_or = ho.rel_l2(out["tc"][2].cpu().numpy(), ref["grad"])
            worst["oracle_grad"] = max(worst["oracle_grad"], e_or)
        worst["loss"], worst["hist"], worst["grad"] = max(worst["loss"], e_loss), max(worst["hist"], e_hist), max(worst["grad"], e_grad)
        ok = e_loss < 1e-5 and e_hist < 1e-5 and e_grad < 1e-5 and (e_or is None or e_or < 1e-5) and bool(torch.isfinite(out["tc"][2]).all())
        _print(("ok   " if ok else "FAIL ") + tag + f"  loss {e_loss:.1e} hist {e_hist:.1e} grad {e_grad:.1e}" + (f" oracle-grad {e_or:.1e}" if e_or is not None else ""), flush=True)
        if not ok:
            raise AssertionError(tag)
    _print("worst:", {k: f"{v:.2e}" for k, v in worst.items()})
    return worst


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 60, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
