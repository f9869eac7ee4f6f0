"""Generator and CPU check of the 128-entry table behind `log_pos` (palette_and_histo_gan_b200/csrc/common.cuh), the
float64 logarithm the pixel terms of the histogram kernels take (histogram.py:58-66: `tf.math.log(x + eps)`).

Test infrastructure (oracle/): nothing on the product path imports this.  `python -m oracle.make_log_table` prints the
table as C initialisers; tests/test_oracle.py::test_log_table checks that common.cuh holds exactly these values and that
the kernel's formula, replayed in numpy float64, stays within 3e-15 of long-double logs on [1e-6, 1].

x = 2^k m, m in [1, 2); the top 7 mantissa bits pick the interval with centre c = 1 + (i + 1/2) / 128;
entry = (RN(1 / c), -log(RN(1 / c))): with the ROUNDED reciprocal the reduction r = m RN(1 / c) - 1 is exact up to
the one rounding of the FMA, and log x = k ln 2 + entry[1] + log1p(r), |r| < 2^-8, six series terms.
"""
import numpy as np

LD = np.longdouble


def table():
    rows = []
    for i in range(128):
        c = LD(1) + (LD(i) + LD(0.5)) / LD(128)
        inv = float(LD(1) / c)
        rows.append((inv, float(-np.log(LD(inv)))))
    return rows


def log_pos_numpy(x, rows=None):
    """The arithmetic of log_pos in numpy float64 (no FMA: the product m * inv is rounded once more than on the GPU)."""
    rows = rows or table()
    x = np.asarray(x, np.float64)
    bits = x.view(np.int64)
    hi = bits >> 32
    k = (hi >> 20) - 1023
    idx = (hi >> 13) & 127
    m = ((bits & 0x000FFFFFFFFFFFFF) | (0x3FF << 52)).view(np.float64)
    inv = np.array([r[0] for r in rows])[idx]
    nl = np.array([r[1] for r in rows])[idx]
    r = m * inv - 1.0
    q = (-1.0 / 6.0) * r + 0.2
    q = q * r - 0.25
    q = q * r + 1.0 / 3.0
    q = q * r - 0.5
    q = q * r + 1.0
    return (k * 0.6931471805599453 + nl) + r * q


def as_c_initialisers(rows=None):
    rows = rows or table()
    out = []
    for j, (a, b) in enumerate(rows):
        out.append("    {%s, %s},%s" % (float(a).hex(), float(b).hex(), "\n" if j % 2 == 1 else " "))
    return "".join(out).rstrip()


if __name__ == "__main__":
    print(as_c_initialisers())


# ---- float32 variant: the pixel pass of the mirrored-tile forward (hist_tc.cu: logdiff_f32) ---------------------------
# log x = k ln2 + L_i + log1p(r), r = fma(m, RN32(1 / c_i), -1).  L_i = -log(RN32(1 / c_i)) is stored as
# L_hi (a multiple of 2^-23, so differences of two L_hi are exact in float32) + L_lo; ln2 = LN2_HI (a multiple of 2^-18:
# k LN2_HI is exact for |k| <= 32) + LN2_LO.  The DIFFERENCE of two logs then has an exactly known large part
# (dk LN2_HI + dL_hi, one FMA whose rounding error is recovered) and a small part (< 0.01) evaluated in float32.
LN2_HI = float(np.float32(round(float(np.log(LD(2))) * 2 ** 18) / 2 ** 18))
LN2_LO = float(np.float32(np.log(LD(2)) - LD(LN2_HI)))


def table_f32():
    rows = []
    for i in range(128):
        c = LD(1) + (LD(i) + LD(0.5)) / LD(128)
        inv = np.float32(LD(1) / c)
        big = -np.log(LD(inv))
        l_hi = np.float32(round(float(big) * 2 ** 23) / 2 ** 23)
        l_lo = np.float32(big - LD(l_hi))
        rows.append((float(inv), float(l_hi), float(l_lo)))
    return rows


def logdiff_f32_numpy(x0, x1, rows=None):
    """log(x0) - log(x1) with the kernel's float32 arithmetic (numpy float32; FMAs emulated in float64, which is exact
    for these operand sizes)."""
    rows = rows or table_f32()
    f32 = np.float32
    tab = np.array(rows, dtype=np.float64)

    def parts(x):
        x = np.asarray(x, f32)
        bits = x.view(np.int32)
        k = ((bits >> 23) - 127).astype(f32)
        idx = (bits >> 16) & 127
        m = ((bits & 0x007FFFFF) | 0x3F800000).astype(np.int32).view(f32)
        inv, l_hi, l_lo = tab[idx, 0], tab[idx, 1].astype(f32), tab[idx, 2].astype(f32)
        r = (m.astype(np.float64) * inv - 1.0).astype(f32)                      # fma
        q = (f32(1.0 / 3.0) + (r.astype(np.float64) * -0.25)).astype(f32)       # fma(r, -1/4, 1/3)
        q = (f32(-0.5) + r.astype(np.float64) * q).astype(f32)                  # fma(r, q, -1/2)
        s = ((r * r).astype(f32) * q).astype(f32)
        return k, l_hi, ((l_lo + r).astype(f32) + s).astype(f32)

    k0, h0, s0 = parts(x0)
    k1, h1, s1 = parts(x1)
    dk, dh = (k0 - k1).astype(f32), (h0 - h1).astype(f32)
    big = (dk.astype(np.float64) * LN2_HI + dh).astype(f32)                     # fma
    err = ((dk.astype(np.float64) * LN2_HI - big).astype(f32) + dh).astype(f32)
    small = ((s0 - s1).astype(f32) + (dk * f32(LN2_LO)).astype(f32)).astype(f32)
    return (big + (err + small).astype(f32)).astype(f32)


def as_c_initialisers_f32(rows=None):
    rows = rows or table_f32()
    out = []
    for j, (a, b, c) in enumerate(rows):
        out.append("    {%s, %s, %s, 0.0f},%s" % (float(a).hex() + "f", float(b).hex() + "f", float(c).hex() + "f",
                                                 "\n" if j % 2 == 1 else " "))
    return "".join(out).rstrip()
