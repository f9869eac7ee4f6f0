"""torch-CPU op-for-op port of `histogram.py` with autograd (TEST INFRASTRUCTURE, see oracle/__init__.py).

Used for (1) an independent check of the analytic gradient of `histogram_oracle.hist_loss_and_grad_f64`
and (2) the `cpu_baseline` / `--impl reference` arm of bench.py: it performs the same tensor ops,
in the same order and with the same materialised (B,N,S) temporaries, that TensorFlow executes for
`histogram.py:5-89` + `tape.gradient` (pix2pix_model.py:78), on all host threads.
It is a PORT (TensorFlow is not installable here), labelled `"kind": "port"` wherever it is timed.
"""
from __future__ import annotations

import numpy as np
import torch

from .histogram_oracle import tf_linspace_f32, sigma_sqr_f32


def calculate_component_histogram(component, projection1, projection2, color_intensities,
                                  histogram_domain, method, sigma_sqr, epsilon):
    # histogram.py:13-30
    iu = torch.log(component + epsilon) - torch.log(projection1 + epsilon)
    iu = iu.unsqueeze(-1)
    iv = torch.log(component + epsilon) - torch.log(projection2 + epsilon)
    iv = iv.unsqueeze(-1)
    diff_u = torch.pow(iu - histogram_domain, 2.0) / sigma_sqr
    diff_v = torch.pow(iv - histogram_domain, 2.0) / sigma_sqr
    if method == "RBF":
        diff_u = torch.exp(-diff_u)
        diff_v = torch.exp(-diff_v)
    elif method == "inverse-quadratic":
        diff_u = 1.0 / (1.0 + diff_u)
        diff_v = 1.0 / (1.0 + diff_v)
    else:
        raise ValueError(f"unknown histogram method {method!r}")
    a = (color_intensities * diff_u).transpose(1, 2)
    return torch.matmul(a, diff_v)


def calculate_rgbuv_histogram(image_batch, size=64, method="inverse-quadratic", sigma=0.02):
    # histogram.py:53-81
    dtype = image_batch.dtype
    epsilon = 1e-6
    sigma_sqr = float(sigma_sqr_f32(sigma))
    dom = torch.from_numpy(tf_linspace_f32(-3.0, 3.0, size)).to(dtype).unsqueeze(0)
    image_batch = image_batch * 0.5 + 0.5
    b = image_batch.shape[0]
    image_batch = image_batch[:, :, :, :3]
    i_ = image_batch.reshape(b, -1, 3)
    ii = torch.pow(i_, 2)
    iy = torch.sqrt(ii[..., 0] + ii[..., 1] + ii[..., 2] + epsilon).unsqueeze(-1)
    r, g, bl = i_[..., 0], i_[..., 1], i_[..., 2]
    hr = calculate_component_histogram(r, g, bl, iy, dom, method, sigma_sqr, epsilon)
    hg = calculate_component_histogram(g, r, bl, iy, dom, method, sigma_sqr, epsilon)
    hb = calculate_component_histogram(bl, r, g, iy, dom, method, sigma_sqr, epsilon)
    h = torch.stack([hr, hg, hb], -1)
    denom = h.sum(dim=(1, 2, 3), keepdim=True)
    return h / denom


def hellinger_loss(y_true, y_pred):
    # histogram.py:84-89
    b = float(y_true.shape[0])
    return (1.0 / np.sqrt(2.0) * torch.sqrt(torch.sum(torch.pow(torch.sqrt(y_pred) - torch.sqrt(y_true), 2.0)))) / b


def hist_loss_fwd_bwd(real, fake, size=64, method="inverse-quadratic", sigma=0.02):
    """One generator-loss evaluation of the histogram term: fwd(real), fwd(fake), Hellinger, backward
    to the fake image (pix2pix_model.py:243-245, :78).  Returns (loss, grad_fake)."""
    fake = fake.detach().clone().requires_grad_(True)
    with torch.no_grad():
        ht = calculate_rgbuv_histogram(real, size, method, sigma)
    hp = calculate_rgbuv_histogram(fake, size, method, sigma)
    loss = hellinger_loss(ht, hp)
    loss.backward()
    return loss.detach(), fake.grad
