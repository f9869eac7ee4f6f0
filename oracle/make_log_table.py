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
