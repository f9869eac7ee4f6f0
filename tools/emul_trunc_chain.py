"""CPU model of the 256-bin forward's accumulation (hist_tc_fwd256.cu): fp16 hi + lo operands, three products per
16-pixel K step, every tcgen05.mma adding its (exact) 16-term product sum into the fp32 accumulator with TRUNCATION,
chains of `chain_px` pixels summed in fp32 round-to-nearest.  Shows how the histogram error against the float64
oracle depends on the chain length and on float32 vs float64 log-chroma — the two knobs of the kernel.

    python tools/emul_trunc_chain.py            # one dense 64 x 64 image, 256 bins (a minute on the CPU)
"""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np

from oracle import histogram_oracle as ho


def trunc32(x64):
    """float64 -> float32, rounded toward zero (what the tensor core does when it aligns the addends)."""
    f = x64.astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(x64)
    return np.where(over, np.nextafter(f, np.float32(0)), f).astype(np.float32)


def split_f16(x32, scale):
    xs = (x32.astype(np.float32) * np.float32(scale)).astype(np.float32)
    hi = xs.astype(np.float16)
    lo = (xs - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64)


def forward(img, bins, chain_px, logs64, sigma=0.02):
    dom = ho.tf_linspace_f32(-3.0, 3.0, bins)
    s2 = float(ho.sigma_sqr_f32(sigma))
    x = img.astype(np.float64)[..., :3].reshape(-1, 3) * 0.5 + 0.5
    iy = np.sqrt((x ** 2).sum(-1) + 1e-6)
    if logs64:
        lg = np.log(x + 1e-6)
        diff = lambda a, b: lg[:, a] - lg[:, b]
    else:  # one float32 log of the ratio, as the 64-bin kernel takes it
        x32 = (x + 1e-6).astype(np.float32)
        diff = lambda a, b: np.log((x32[:, a] / x32[:, b]).astype(np.float32)).astype(np.float32).astype(np.float64)
    n = x.shape[0]
    k = int(round(-6.75 - 0.5 * np.log2(s2)))
    sc = 2.0 ** k
    w = np.float32(sc * sc * s2)
    out = np.zeros((bins, bins, 3))
    for c, (cc, p1, p2) in enumerate(ho._CHANNEL_TRIPLES):
        u, v = diff(cc, p1), diff(cc, p2)
        # weights as the kernel generates them: K / w = 1 / (d d + w), d = s (x - c), float32
        du = ((u[:, None] - dom.astype(np.float64)) * sc).astype(np.float32)
        dv = ((v[:, None] - dom.astype(np.float64)) * sc).astype(np.float32)
        ku = (np.float32(1) / (du * du + w)).astype(np.float32)
        kv = (np.float32(1) / (dv * dv + w)).astype(np.float32)
        a_hi, a_lo = split_f16((ku * iy[:, None].astype(np.float32)).astype(np.float32), 1.0)
        b_hi, b_lo = split_f16(kv, 1.0)
        total = np.zeros((bins, bins), np.float32)
        for p0 in range(0, n, chain_px):
            acc = np.zeros((bins, bins), np.float32)
            for q in range(p0, min(p0 + chain_px, n), 16):
                s = slice(q, q + 16)
                for a, b in ((a_hi, b_hi), (a_hi, b_lo), (a_lo, b_hi)):
                    acc = trunc32(acc.astype(np.float64) + a[s].T @ b[s])
            total = (total + acc).astype(np.float32)   # the bulk reduction at the L2: round to nearest
        out[:, :, c] = total
    return out / out.sum()


def main():
    rng = np.random.default_rng(47)
    side, bins = 64, 256
    img = np.tanh(rng.standard_normal((1, side, side, 4))).astype(np.float32)
    ref, _ = ho.rgbuv_histogram_f64(img, size=bins)
    print(f"one dense {side}x{side} image, {bins} bins: histogram rel-L2 against the float64 oracle")
    for logs64 in (False, True):
        for chain in (256, 512, 1024, 4096):
            h = forward(img[0], bins, chain, logs64)
            print(f"  logs {'float64 hi+lo' if logs64 else 'float32      '}  chain {chain:5d} px: {ho.rel_l2(h, ref[0]):.2e}", flush=True)


if __name__ == "__main__":
    main()
