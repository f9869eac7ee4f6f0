"""numpy oracle for `histogram.py` of the reference (TEST INFRASTRUCTURE, see oracle/__init__.py).

Two restatements of the same algorithm:

* ``*_f64``  – the arithmetic of the reference carried out in float64 on the float32 inputs
  (float32 image, float32 bin centres, float32 sigma^2): the arbiter for the 1e-5 bar.
* ``*_f32``  – op-for-op float32 restatement in the order TensorFlow evaluates
  `histogram.py:13-30,53-81,84-89` (every intermediate rounded to float32).

plus the analytic backward of the whole loss (the reference relies on TF autodiff,
`pix2pix_model.py:78`; formulas in SURVEY.md §8a row H7), verified against central finite
differences and torch autograd in tests/test_oracle.py.

PARITY UNPINNED (no TF in this image, the reference has no golden vectors).
"""
from __future__ import annotations

import numpy as np

EPSILON = 1e-6  # histogram.py:53
SQRT2 = np.sqrt(2.0)


def tf_linspace_f32(start: float, stop: float, num: int) -> np.ndarray:
    """`tf.linspace(start, stop, num)` in float32 (histogram.py:55).

    TF 2.9 `linspace_nd`: delta = (stop-start)/(num-1); result = concat(start,
    start + delta*range(1, num-1), stop), all in the dtype of start (float32).
    """
    start = np.float32(start)
    stop = np.float32(stop)
    if num == 1:
        return np.array([start], dtype=np.float32)
    delta = np.float32((stop - start) / np.float32(num - 1))
    mid = (start + delta * np.arange(1, num - 1, dtype=np.float32)).astype(np.float32)
    return np.concatenate([[start], mid, [stop]]).astype(np.float32)


def sigma_sqr_f32(sigma: float = 0.02) -> np.float32:
    """`tf.pow(sigma, 2)` on a python float -> float32 tensor (histogram.py:54)."""
    s = np.float32(sigma)
    return np.float32(s * s)


def _bin_kernel(t, method):
    if method == "inverse-quadratic":  # histogram.py:25-27
        return 1.0 / (1.0 + t)
    if method == "RBF":  # histogram.py:22-24
        return np.exp(-t)
    raise ValueError(f"unknown histogram method {method!r}")


# --------------------------------------------------------------------------------------
# float64 ground truth
# --------------------------------------------------------------------------------------
def _pixel_terms_f64(img):
    """histogram.py:58-69 in float64. img (B,H,W,>=3) float32 in [-1,1]."""
    x = img.astype(np.float64)[..., :3] * 0.5 + 0.5
    b = x.shape[0]
    x = x.reshape(b, -1, 3)
    iy = np.sqrt((x * x).sum(-1) + EPSILON)
    lg = np.log(x + EPSILON)
    return x, iy, lg


# (component, projection1, projection2) per output channel, histogram.py:72-74
_CHANNEL_TRIPLES = ((0, 1, 2), (1, 0, 2), (2, 0, 1))


def raw_histogram_f64(img, dom, sigma_sqr, method="inverse-quadratic"):
    """Un-normalised (B,S,S,3) histogram, float64 (histogram.py:13-30, 72-75)."""
    x, iy, lg = _pixel_terms_f64(np.asarray(img, dtype=np.float32))
    dom = np.asarray(dom, dtype=np.float32).astype(np.float64).reshape(-1)
    s2 = float(np.float32(sigma_sqr))
    bsz, n = iy.shape
    size = dom.shape[0]
    out = np.empty((bsz, size, size, 3), dtype=np.float64)
    for b in range(bsz):
        for c, (cc, p1, p2) in enumerate(_CHANNEL_TRIPLES):
            u = lg[b, :, cc] - lg[b, :, p1]
            v = lg[b, :, cc] - lg[b, :, p2]
            ku = _bin_kernel((u[:, None] - dom[None, :]) ** 2 / s2, method)
            kv = _bin_kernel((v[:, None] - dom[None, :]) ** 2 / s2, method)
            out[b, :, :, c] = (iy[b, :, None] * ku).T @ kv
    return out


def rgbuv_histogram_f64(img, size=64, method="inverse-quadratic", sigma=0.02, dom=None):
    """`calculate_rgbuv_histogram` (histogram.py:36-81) in float64; returns (hist, denom)."""
    if dom is None:
        dom = tf_linspace_f32(-3.0, 3.0, size)
    raw = raw_histogram_f64(img, dom, sigma_sqr_f32(sigma), method)
    denom = raw.sum(axis=(1, 2, 3), keepdims=True)
    return raw / denom, denom.reshape(-1)


def hellinger_loss_f64(y_true, y_pred):
    """histogram.py:84-89 in float64."""
    y_true = np.asarray(y_true, dtype=np.float64)
    y_pred = np.asarray(y_pred, dtype=np.float64)
    bsz = y_true.shape[0]
    ssum = ((np.sqrt(y_pred) - np.sqrt(y_true)) ** 2).sum()
    return (1.0 / SQRT2) * np.sqrt(ssum) / bsz


def l1_loss_f64(y_true, y_pred):  # histogram.py:92-93
    return np.abs(np.asarray(y_true, np.float64) - np.asarray(y_pred, np.float64)).mean()


def l2_loss_f64(y_true, y_pred):  # histogram.py:96-97
    return ((np.asarray(y_true, np.float64) - np.asarray(y_pred, np.float64)) ** 2).mean()


def hist_loss_and_grad_f64(real, fake, size=64, method="inverse-quadratic", sigma=0.02, dom=None,
                           global_batch=None, global_ssum=None):
    """Loss of `Pix2PixHistogramModel.generator_loss` (pix2pix_model.py:243-245) and its gradient
    with respect to the fake image, float64, analytic.

    `global_batch` / `global_ssum` let a shard of a larger batch be evaluated with the whole-batch
    scalars (the loss couples images only through sum-of-squares S and the batch size).
    Returns dict(loss, grad (B,H,W,C) float64, ssum, hist_real, hist_fake, denom_fake).
    """
    real = np.asarray(real, dtype=np.float32)
    fake = np.asarray(fake, dtype=np.float32)
    if dom is None:
        dom = tf_linspace_f32(-3.0, 3.0, size)
    dom64 = np.asarray(dom, np.float32).astype(np.float64).reshape(-1)
    s2 = float(sigma_sqr_f32(sigma))
    ht, _ = rgbuv_histogram_f64(real, size, method, sigma, dom)
    hp, denom = rgbuv_histogram_f64(fake, size, method, sigma, dom)
    bsz = fake.shape[0]
    ssum_local = ((np.sqrt(hp) - np.sqrt(ht)) ** 2).sum()
    ssum = ssum_local if global_ssum is None else float(global_ssum)
    gb = bsz if global_batch is None else int(global_batch)
    loss = (1.0 / SQRT2) * np.sqrt(ssum) / gb

    # dL/dHp, then through the per-image normalisation Hp = Hraw / D
    g = (1.0 - np.sqrt(ht / hp)) / (2.0 * SQRT2 * gb * np.sqrt(ssum))
    ghat = (g - (g * hp).sum(axis=(1, 2, 3), keepdims=True)) / denom.reshape(-1, 1, 1, 1)

    x, iy, lg = _pixel_terms_f64(fake)
    n = iy.shape[1]
    grad_x = np.zeros((bsz, n, 3), dtype=np.float64)
    for b in range(bsz):
        d_iy = np.zeros(n)
        d_l = np.zeros((n, 3))
        for c, (cc, p1, p2) in enumerate(_CHANNEL_TRIPLES):
            gm = ghat[b, :, :, c]
            u = lg[b, :, cc] - lg[b, :, p1]
            v = lg[b, :, cc] - lg[b, :, p2]
            du = u[:, None] - dom64[None, :]
            dv = v[:, None] - dom64[None, :]
            ku = _bin_kernel(du ** 2 / s2, method)
            kv = _bin_kernel(dv ** 2 / s2, method)
            if method == "inverse-quadratic":
                dku = -2.0 * du / s2 * ku * ku
                dkv = -2.0 * dv / s2 * kv * kv
            else:  # RBF
                dku = -2.0 * du / s2 * ku
                dkv = -2.0 * dv / s2 * kv
            p = kv @ gm.T  # (n, S): sum_j G[i,j] Kv[n,j]
            q = ku @ gm  # (n, S): sum_i G[i,j] Ku[n,i]
            d_iy += (ku * p).sum(-1)
            g_u = iy[b] * (dku * p).sum(-1)
            g_v = iy[b] * (dkv * q).sum(-1)
            d_l[:, cc] += g_u + g_v
            d_l[:, p1] -= g_u
            d_l[:, p2] -= g_v
        grad_x[b] = d_l / (x[b] + EPSILON) + d_iy[:, None] * x[b] / iy[b][:, None]
    grad = np.zeros(fake.shape, dtype=np.float64)
    grad[..., :3] = 0.5 * grad_x.reshape(fake.shape[:-1] + (3,))
    return dict(loss=loss, grad=grad, ssum=ssum_local, hist_real=ht, hist_fake=hp, denom_fake=denom)


# --------------------------------------------------------------------------------------
# float32 op-for-op restatement
# --------------------------------------------------------------------------------------
def component_histogram_f32(component, projection1, projection2, color_intensities, histogram_domain,
                            method, sigma_sqr, epsilon):
    """`calculate_component_histogram` (histogram.py:5-32), every op in float32."""
    f = np.float32
    component = np.asarray(component, f)
    eps = f(epsilon)
    s2 = f(sigma_sqr)
    lc = np.log(component + eps, dtype=f)
    iu = (lc - np.log(np.asarray(projection1, f) + eps, dtype=f))[..., None]
    iv = (lc - np.log(np.asarray(projection2, f) + eps, dtype=f))[..., None]
    dom = np.asarray(histogram_domain, f)
    du = (iu - dom).astype(f)
    dv = (iv - dom).astype(f)
    diff_u = ((du * du).astype(f) / s2).astype(f)
    diff_v = ((dv * dv).astype(f) / s2).astype(f)
    if method == "RBF":
        diff_u = np.exp(-diff_u, dtype=f)
        diff_v = np.exp(-diff_v, dtype=f)
    elif method == "inverse-quadratic":
        diff_u = (f(1.0) / (f(1.0) + diff_u)).astype(f)
        diff_v = (f(1.0) / (f(1.0) + diff_v)).astype(f)
    else:
        raise ValueError(f"unknown histogram method {method!r}")
    a = np.transpose((np.asarray(color_intensities, f) * diff_u).astype(f), (0, 2, 1))
    return np.matmul(a, diff_v).astype(f)


def rgbuv_histogram_f32(img, size=64, method="inverse-quadratic", sigma=0.02):
    """`calculate_rgbuv_histogram` (histogram.py:36-81), every op in float32."""
    f = np.float32
    eps = f(EPSILON)
    s2 = sigma_sqr_f32(sigma)
    dom = tf_linspace_f32(-3.0, 3.0, size)[None, :]
    x = (np.asarray(img, f) * f(0.5) + f(0.5)).astype(f)[:, :, :, :3]
    b = x.shape[0]
    i_ = x.reshape(b, -1, 3)
    ii = (i_ * i_).astype(f)
    iy = np.sqrt(((ii[..., 0] + ii[..., 1]).astype(f) + ii[..., 2]).astype(f) + eps, dtype=f)[..., None]
    r, g, bl = i_[..., 0], i_[..., 1], i_[..., 2]
    hr = component_histogram_f32(r, g, bl, iy, dom, method, s2, eps)
    hg = component_histogram_f32(g, r, bl, iy, dom, method, s2, eps)
    hb = component_histogram_f32(bl, r, g, iy, dom, method, s2, eps)
    h = np.stack([hr, hg, hb], -1)
    denom = h.sum(axis=(1, 2, 3), keepdims=True, dtype=f)
    return (h / denom).astype(f)


def hellinger_loss_f32(y_true, y_pred):
    """histogram.py:84-89 in float32."""
    f = np.float32
    y_true = np.asarray(y_true, f)
    y_pred = np.asarray(y_pred, f)
    bsz = f(y_true.shape[0])
    d = (np.sqrt(y_pred, dtype=f) - np.sqrt(y_true, dtype=f)).astype(f)
    ssum = (d * d).astype(f).sum(dtype=f)
    return f(f(f(1.0) / np.sqrt(f(2.0))) * np.sqrt(ssum, dtype=f)) / bsz


# --------------------------------------------------------------------------------------
# error metrics used by every parity test (the bar is norm-relative)
# --------------------------------------------------------------------------------------
def rel_l2(a, ref):
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.linalg.norm((a - ref).ravel()) / max(np.linalg.norm(ref.ravel()), 1e-300))


def rel_max(a, ref):
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-300))
