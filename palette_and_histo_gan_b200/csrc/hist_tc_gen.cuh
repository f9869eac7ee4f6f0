// Pieces shared by the tensor-core forward kernels (hist_tc.cu: 64-bin tile, hist_tc_fwd256.cu: 256-bin tile):
// generation of the scaled bin weights and the named barrier of the producer warps.
#pragma once

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ph {
namespace tcgen {

using namespace tc;

// largest intensity (x multiplicity x intensity scale) whose product with a scaled weight (<= 2^14) stays below
// fp16's 65504; images in [-1, 1] reach sqrt(3)
constexpr float IY_OPERAND_LIMIT = 3.99f;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// two scaled bin weights at once (two pixels, one bin): d = x + (-c)
template <int METHOD>
__device__ __forceinline__ f32x2 weight2(f32x2 x, f32x2 negc, f32x2 wa2, f32x2 wb2) {
  const f32x2 d = add2(x, negc);
  if (METHOD == PH_METHOD_INVERSE_QUADRATIC) {
    const f32x2 e = fma2(d, d, wb2);
    // one MUFU.RCP for the two weights (1/e0 = e1 / (e0 e1), 1/e1 = e0 / (e0 e1)): with one reciprocal per weight
    // the forward is bound by the MUFU pipe (16/clk/SM: 768 cycles per 32-pixel stage), measured 5 % slower
    const float e0 = lo_of(e), e1 = hi_of(e);
    const float r = fast_rcp(e0 * e1);
    return pack2(r * e1, r * e0);
  } else {
    const f32x2 e = fma2(mul2(d, d), wa2, wb2);
    return pack2(fast_ex2(lo_of(e)), fast_ex2(hi_of(e)));
  }
}

// four scaled bin weights at once (four pixels, one bin).  Inverse-quadratic: the two packed denominators are
// multiplied lane-wise, p = e_a (.) e_b, one reciprocal per lane of p serves two weights
// (1 / e_a = e_b / p, 1 / e_b = e_a / p) and both back-multiplications are packed: 9 instructions for four weights
// (2 FADD2, 2 FFMA2, 3 FMUL2, 2 MUFU) against 12 with the products taken inside one packed pair
// (the two lanes of one register are then multiplied by each other, which the packed multiply cannot do)
template <int METHOD>
__device__ __forceinline__ void weight4(f32x2 xa, f32x2 xb, f32x2 ya, f32x2 yb, f32x2 wa2, f32x2 wb2, f32x2& ka, f32x2& kb) {
  // d = x + y with separate second terms for the two pairs (hi + lo coordinates of the 256-bin kernel)
  if (METHOD == PH_METHOD_INVERSE_QUADRATIC) {
    const f32x2 da = add2(xa, ya), db = add2(xb, yb);
    const f32x2 ea = fma2(da, da, wb2), eb = fma2(db, db, wb2);
    const f32x2 p = mul2(ea, eb);
    const f32x2 r = pack2(fast_rcp(lo_of(p)), fast_rcp(hi_of(p)));
    ka = mul2(r, eb);
    kb = mul2(r, ea);
  } else {
    ka = weight2<METHOD>(xa, ya, wa2, wb2);
    kb = weight2<METHOD>(xb, yb, wa2, wb2);
  }
}
template <int METHOD>
__device__ __forceinline__ void weight4(f32x2 xa, f32x2 xb, f32x2 negc, f32x2 wa2, f32x2 wb2, f32x2& ka, f32x2& kb) {
  if (METHOD == PH_METHOD_INVERSE_QUADRATIC) {
    const f32x2 da = add2(xa, negc), db = add2(xb, negc);
    const f32x2 ea = fma2(da, da, wb2), eb = fma2(db, db, wb2);
    const f32x2 p = mul2(ea, eb);
    const f32x2 r = pack2(fast_rcp(lo_of(p)), fast_rcp(hi_of(p)));
    ka = mul2(r, eb);
    kb = mul2(r, ea);
  } else {
    ka = weight2<METHOD>(xa, negc, wa2, wb2);
    kb = weight2<METHOD>(xb, negc, wa2, wb2);
  }
}

// ---- backward: weight and derivative of two bins at once -----------------------------------------------
// two bins at once (packed fp32x2): d = (x_hi - c) + x_lo
template <int METHOD>
__device__ __forceinline__ void weight_pair2(f32x2 x_hi, f32x2 x_lo, f32x2 c, f32x2 wa2, f32x2 wb2, f32x2 wc2, f32x2 mone2,
                                             f32x2& k, f32x2& dk) {
  const f32x2 d = add2(fma2(c, mone2, x_hi), x_lo);
  if (METHOD == PH_METHOD_INVERSE_QUADRATIC) {
    const f32x2 e = fma2(d, d, wb2);
    // one MUFU.RCP per weight: here the issue slots are scarcer than the MUFU pipe (the shared-reciprocal form
    // of the forward kernel, 1 MUFU + 3 FMUL per pair, measured 6 % slower)
    k = pack2(fast_rcp(lo_of(e)), fast_rcp(hi_of(e)));
    dk = mul2(mul2(d, k), k);
  } else {
    const f32x2 e = fma2(mul2(d, d), wa2, wb2);
    k = pack2(fast_ex2(lo_of(e)), fast_ex2(hi_of(e)));
    dk = mul2(mul2(d, wc2), k);
  }
}

// Scales of the backward operands (Params of hist_tc_bwd.cu): k_s = S_k k, dk_s = S_dk dk' with dk' the derivative
// without the common factor -2 / sigma^2; powers of two that keep the largest weight below fp16's 65504:
//   IQ:  k <= 1, |dk'| = |d| k^2 <= 0.325 sigma;   RBF: k <= 1, |dk'| = |d| k <= 0.43 sigma
struct BwdScales {
  float wa, wb, wc, coord_scale, inv_sk2, inv_sk_sdk;
};
static inline BwdScales bwd_scales(int method, float sigma_sqr) {
  BwdScales r;
  const double sigma = sqrt((double)sigma_sqr), inv_s2 = 1.0 / (double)sigma_sqr;
  double s_k, s_dk;
  if (method == PH_METHOD_INVERSE_QUADRATIC) {
    // s = the power of two nearest to 2^-5.25 / sigma: w = s^2 sigma^2 in [2^-11, 2^-10];  k_s = k / w <= 2^11,
    // dk_s = s d k^2 / w^2 <= 0.325 w^-1.5 <= 30 100
    const int kexp = (int)lrint(-5.25 - 0.5 * log2((double)sigma_sqr));
    const double sc = ldexp(1.0, kexp);
    r.coord_scale = (float)sc;
    r.wa = 0.f;
    r.wb = (float)(sc * sc * (double)sigma_sqr);
    r.wc = 1.0f;
    const double w = (double)r.wb;
    s_k = 1.0 / w;
    s_dk = sc / (w * w);
  } else {
    int e = (int)floor(log2(60000.0 / (0.43 * sigma)));
    e = e > 60 ? 60 : (e < -60 ? -60 : e);
    r.coord_scale = 1.0f;
    s_k = 16384.0;
    s_dk = ldexp(1.0, e);
    r.wa = (float)(-inv_s2 * 1.4426950408889634);
    r.wb = 14.0f;
    r.wc = (float)(s_dk / s_k);
  }
  r.inv_sk2 = (float)(1.0 / (s_k * s_k));
  r.inv_sk_sdk = (float)(-2.0 * inv_s2 / (s_k * s_dk));
  return r;
}

// Scales of the generated operands (see Params of the kernels): the inverse-quadratic weight is produced as
// K / w = 1 / (d d + w) with coordinates pre-multiplied by the power of two s, w = s^2 sigma^2 in [2^-14, 2^-13];
// the RBF weight as 2^14 K = 2^(wa d d + 14).
struct WeightScales {
  float wa, wb, coord_scale;
  double weight_scale;  // generated weight = weight_scale * K
};
static inline WeightScales weight_scales(int method, float sigma_sqr) {
  WeightScales ws;
  if (method == PH_METHOD_INVERSE_QUADRATIC) {
    const int k = (int)lrint(-6.75 - 0.5 * log2((double)sigma_sqr));
    const double sc = ldexp(1.0, k), w = sc * sc * (double)sigma_sqr;
    ws.coord_scale = (float)sc;
    ws.wa = 0.f;
    ws.wb = (float)w;
    ws.weight_scale = 1.0 / (double)ws.wb;
  } else {
    ws.coord_scale = 1.0f;
    ws.wa = (float)(-1.4426950408889634 / (double)sigma_sqr);
    ws.wb = 14.0f;
    ws.weight_scale = 16384.0;
  }
  return ws;
}

}  // namespace tcgen
}  // namespace ph
