// Pieces shared by the tensor-core forward kernels (hist_tc.cu: 64-bin tile, hist_tc_fwd256.cu: 256-bin tile):
// generation of the scaled bin weights and the named barrier of the producer warps.
#pragma once

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ph {
namespace tcgen {

using namespace tc;

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
