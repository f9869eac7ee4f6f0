"""CPU emulation of the fp16 hi+lo operand split (three products hi.hi + hi.lo + lo.hi, exact products,
float64 accumulation) against the float64 oracle: isolates the error of the operand representation that
the tensor-core kernels use.  Compares with the tf32 split (10-bit truncation hi, fp32 remainder)."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from oracle import histogram_oracle as ho

def split_f16(x32, scale):
    xs = (x32.astype(np.float32) * np.float32(scale)).astype(np.float32)
    hi = xs.astype(np.float16)                                  # cvt.rn.f16x2.f32
    lo = (xs - hi.astype(np.float32)).astype(np.float16)        # FHFMA (exact difference) + cvt.rn
    return hi.astype(np.float64) / scale, lo.astype(np.float64) / scale

def split_tf32(x32, scale=1.0):
    xs = x32.astype(np.float32)
    hi = (xs.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = (xs - hi)
    lo = (lo.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    return hi.astype(np.float64), lo.astype(np.float64)

def mm3(a, b, split, sa, sb, four=False):
    ah, al = split(a, sa); bh, bl = split(b, sb)
    r = ah @ bh + ah @ bl + al @ bh
    if four: r = r + al @ bl
    return r

def run(real, fake, split, name, sa=2.0 ** 14, sg_target=2.0 ** 14, sigma=0.02, method="inverse-quadratic"):
    dom = ho.tf_linspace_f32(-3.0, 3.0, 64); dom64 = dom.astype(np.float64)
    s2 = float(ho.sigma_sqr_f32(sigma))
    ref = ho.hist_loss_and_grad_f64(real, fake, method=method, sigma=sigma)
    x, iy, lg = ho._pixel_terms_f64(fake)
    bsz, n = iy.shape
    # forward
    raw = np.empty((bsz, 64, 64, 3))
    for b in range(bsz):
        for c, (cc, p1, p2) in enumerate(ho._CHANNEL_TRIPLES):
            u = lg[b, :, cc] - lg[b, :, p1]; v = lg[b, :, cc] - lg[b, :, p2]
            ku = ho._bin_kernel((u[:, None] - dom64) ** 2 / s2, method); kv = ho._bin_kernel((v[:, None] - dom64) ** 2 / s2, method)
            A = (iy[b, :, None] * ku).astype(np.float32); Bm = kv.astype(np.float32)
            raw[b, :, :, c] = mm3(A.T.copy(), Bm, split, sa, sa, four=True)
    den = raw.sum(axis=(1, 2, 3), keepdims=True); hp = raw / den
    e_h = ho.rel_l2(hp, ref["hist_fake"])
    # backward with exact ghat (from f64) to isolate the contraction error
    ht = ref["hist_real"]; hp64 = ref["hist_fake"]; ssum = ref["ssum"]
    g = (1.0 - np.sqrt(ht / hp64)) / (2.0 * ho.SQRT2 * bsz * np.sqrt(ssum))
    ghat = (g - (g * hp64).sum(axis=(1, 2, 3), keepdims=True)) / ref["denom_fake"].reshape(-1, 1, 1, 1)
    grad_x = np.zeros((bsz, n, 3))
    for b in range(bsz):
        gmax = np.abs(ghat[b]).max()
        sg = 2.0 ** np.floor(np.log2(sg_target / gmax))
        d_iy = np.zeros(n); d_l = np.zeros((n, 3))
        for c, (cc, p1, p2) in enumerate(ho._CHANNEL_TRIPLES):
            gm = ghat[b, :, :, c].astype(np.float32)
            u = lg[b, :, cc] - lg[b, :, p1]; v = lg[b, :, cc] - lg[b, :, p2]
            du = u[:, None] - dom64; dv = v[:, None] - dom64
            ku = ho._bin_kernel(du ** 2 / s2, method); kv = ho._bin_kernel(dv ** 2 / s2, method)
            if method == "inverse-quadratic":
                dku_ = du * ku * ku; dkv_ = dv * kv * kv
            else:
                dku_ = du * ku; dkv_ = dv * kv
            sdk = 2.0 ** np.floor(np.log2(2.0 ** 14 / (0.45 * sigma)))
            P = mm3(kv.astype(np.float32), gm.T.copy(), split, sa, sg)
            Pp = mm3(dkv_.astype(np.float32), gm.T.copy(), split, sdk, sg)
            d_iy += (ku * P).sum(-1)
            g_u = iy[b] * (-2.0 / s2) * (dku_ * P).sum(-1)
            g_v = iy[b] * (-2.0 / s2) * (ku * Pp).sum(-1)
            d_l[:, cc] += g_u + g_v; d_l[:, p1] -= g_u; d_l[:, p2] -= g_v
        grad_x[b] = d_l / (x[b] + ho.EPSILON) + d_iy[:, None] * x[b] / iy[b][:, None]
    grad = np.zeros(fake.shape); grad[..., :3] = 0.5 * grad_x.reshape(fake.shape[:-1] + (3,))
    print(f"{name:28s} hist relL2 {e_h:.2e}   grad relL2 {ho.rel_l2(grad, ref['grad']):.2e} relmax {ho.rel_max(grad, ref['grad']):.2e}")

rng = np.random.default_rng(0)
dense_r = np.tanh(rng.standard_normal((3, 32, 32, 4))).astype(np.float32); dense_f = np.tanh(rng.standard_normal((3, 32, 32, 4))).astype(np.float32)
sp = np.load("tests/golden/sprites.npz")
keys = list(sp.keys()); print(keys)
for nm, (r, f) in {"dense": (dense_r, dense_f)}.items():
    run(r, f, split_tf32, nm + " tf32x3")
    run(r, f, split_f16, nm + " f16x3 s=2^14")
    run(r, f, split_f16, nm + " f16x3 s=2^8", sa=2.0 ** 8, sg_target=2.0 ** 8)
    run(r, f, split_f16, nm + " f16x3 s=1", sa=1.0, sg_target=1.0)
def norm(u8):
    a = u8.astype(np.float32)
    a = np.where(a[..., 3:4] == 0, 0.0, a) if a.shape[-1] == 4 else a
    return (a / 127.5 - 1.0).astype(np.float32)
fr = norm(sp["front"][:3]); ri = norm(sp["right"][:3])
noisy = np.clip(ri + 0.05 * rng.standard_normal(ri.shape).astype(np.float32), -1, 1).astype(np.float32)
run(fr, ri, split_tf32, "sprite tf32x3"); run(fr, ri, split_f16, "sprite f16x3 2^14")
run(fr, noisy, split_tf32, "noisy sprite tf32x3"); run(fr, noisy, split_f16, "noisy sprite f16x3 2^14")
run(dense_r, dense_f, split_tf32, "dense RBF .5 tf32x3", sigma=0.5, method="RBF"); run(dense_r, dense_f, split_f16, "dense RBF .5 f16x3", sigma=0.5, method="RBF")
run(dense_r, dense_f, split_tf32, "dense s=.2 tf32x3", sigma=0.2); run(dense_r, dense_f, split_f16, "dense s=.2 f16x3", sigma=0.2)
run(dense_r, dense_f, split_tf32, "dense s=.002 tf32x3", sigma=0.002); run(dense_r, dense_f, split_f16, "dense s=.002 f16x3", sigma=0.002)
