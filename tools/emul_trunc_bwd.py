"""CPU model of the backward product chains P = Kv.G^T, P' = dKv.G^T on the tensor cores (hist_tc_bwd.cu: K = 64 bins
= 12 accumulating MMAs per chain; hist_tc_bwd256.cu: K = 256 = 48): fp16 hi + lo operands, three products per 16-bin
K step, every tcgen05.mma adding its exact product sum into the fp32 accumulator with TRUNCATION.  Everything else
(G^, the epilogue dot products) is float64, so the printed gradient error is the chains' own.  `splits` cuts the
K = 256 chain into that many independently accumulated parts added in float64 — what an extra epilogue per part would buy.

    python tools/emul_trunc_bwd.py        # two dense 48 x 48 images of the same distribution (the ill-conditioned case)
"""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np

from oracle import histogram_oracle as ho
from tools.emul_trunc_chain import trunc32


def split_f16(x64, scale):
    xs = (x64 * scale).astype(np.float32)
    hi = xs.astype(np.float16)
    lo = (xs - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64)


def chain(a, g, sa, sg, splits, truncate=True):
    """sum_j a[n, j] g[i, j] -> (n, i), accumulated per 16-bin K step like the MMAs do."""
    ah, al = split_f16(a, sa)
    gh, gl = split_f16(g, sg)
    nb = a.shape[1]
    out = np.zeros((a.shape[0], g.shape[0]))
    per = nb // splits
    for s0 in range(0, nb, per):
        acc = np.zeros((a.shape[0], g.shape[0]), np.float32)
        for k in range(s0, s0 + per, 16):
            ks = slice(k, k + 16)
            for x, y in ((ah, gh), (ah, gl), (al, gh)):
                t = acc.astype(np.float64) + x[:, ks] @ y[:, ks].T
                acc = trunc32(t) if truncate else t.astype(np.float32)
        out += acc.astype(np.float64)
    return out / (sa * sg)


def run(real, fake, bins, splits, truncate=True, sigma=0.02):
    dom = ho.tf_linspace_f32(-3.0, 3.0, bins).astype(np.float64)
    s2 = float(ho.sigma_sqr_f32(sigma))
    ref = ho.hist_loss_and_grad_f64(real, fake, size=bins, sigma=sigma)
    x, iy, lg = ho._pixel_terms_f64(fake)
    bsz, n = iy.shape
    ht, hp, ssum = ref["hist_real"], ref["hist_fake"], ref["ssum"]
    g = (1.0 - np.sqrt(ht / hp)) / (2.0 * ho.SQRT2 * bsz * np.sqrt(ssum))
    ghat = (g - (g * hp).sum(axis=(1, 2, 3), keepdims=True)) / ref["denom_fake"].reshape(-1, 1, 1, 1)
    # operand scales as the kernels choose them (hist_tc_gen.cuh: bwd_scales)
    kexp = int(round(-5.25 - 0.5 * np.log2(s2)))
    sc = 2.0 ** kexp
    w = sc * sc * s2
    s_k, s_dk = 1.0 / w, sc / (w * w)
    grad_x = np.zeros((bsz, n, 3))
    for b in range(bsz):
        sg = 2.0 ** (13 - np.floor(np.log2(np.abs(ghat[b]).max())))
        d_iy = np.zeros(n)
        d_l = np.zeros((n, 3))
        for c, (cc, p1, p2) in enumerate(ho._CHANNEL_TRIPLES):
            gm = ghat[b, :, :, c]
            u = lg[b, :, cc] - lg[b, :, p1]
            v = lg[b, :, cc] - lg[b, :, p2]
            du, dv = u[:, None] - dom, v[:, None] - dom
            ku, kv = 1.0 / (1.0 + du * du / s2), 1.0 / (1.0 + dv * dv / s2)
            dku_, dkv_ = du * ku * ku, dv * kv * kv
            P = chain(kv, gm, s_k, sg, splits, truncate)
            Pp = chain(dkv_, gm, s_dk / sc * sc, sg, splits, truncate)   # |dk'| s_dk stays below fp16's maximum
            d_iy += (ku * P).sum(-1)
            g_u = iy[b] * (-2.0 / s2) * (dku_ * P).sum(-1)
            g_v = iy[b] * (-2.0 / s2) * (ku * Pp).sum(-1)
            d_l[:, cc] += g_u + g_v
            d_l[:, p1] -= g_u
            d_l[:, p2] -= g_v
        grad_x[b] = d_l / (x[b] + ho.EPSILON) + d_iy[:, None] * x[b] / iy[b][:, None]
    grad = np.zeros(fake.shape)
    grad[..., :3] = 0.5 * grad_x.reshape(fake.shape[:-1] + (3,))
    return ho.rel_l2(grad, ref["grad"])


def main():
    rng = np.random.default_rng(3)
    side = 48
    real = np.tanh(rng.standard_normal((1, side, side, 4))).astype(np.float32)
    fake = np.tanh(rng.standard_normal((1, side, side, 4))).astype(np.float32)
    print(f"dense {side}x{side} real and fake of the same distribution: gradient rel-L2 of the product chains alone")
    print(f"   64 bins, one chain of 12 MMAs:                      {run(real, fake, 64, 1):.2e}")
    for splits in (1, 2, 4):
        print(f"  256 bins, {splits} chain(s) of {48 // splits:2d} MMAs:                    {run(real, fake, 256, splits):.2e}")
    print(f"  256 bins, one chain, round-to-nearest accumulation:  {run(real, fake, 256, 1, truncate=False):.2e}")


if __name__ == "__main__":
    main()
