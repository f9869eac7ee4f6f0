// Shared helpers of libpalhist: error reporting, launch accounting, per-pixel colour terms.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/palhist.h"

namespace ph {

// ---- thread-local error message / launch counter (defined in abi.cu) -------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define PH_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ph::set_error(__VA_ARGS__);          \
      return PH_ERR_INVALID;               \
    }                                      \
  } while (0)

#define PH_CUDA_OK(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ph::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                    __LINE__);                                                        \
      return PH_ERR_CUDA;                                                             \
    }                                                                                 \
  } while (0)

#define PH_LAUNCH_OK(name)                                                          \
  do {                                                                              \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      ph::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));       \
      return PH_ERR_CUDA;                                                           \
    }                                                                               \
    ph::count_launch();                                                             \
  } while (0)

// ---- programmatic dependent launch (PDL): a kernel launched with launch_pdl may become resident while the kernel before
// it in the stream is still draining (its CTAs start as SMs free up); everything that reads what earlier kernels wrote
// must come after pdl_wait(), which returns once ALL earlier work of the stream has completed and is visible.  The
// kernels of the loss step keep only their set-up (shared-memory barriers, tensor-memory allocation) in front of it,
// and release their own dependents right behind it, so launch latency and set-up overlap the previous kernel's tail.
// Launched without the attribute both instructions are no-ops.  PH_PDL=0 launches everything the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();  // abi.cu
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---- per-pixel colour terms (histogram.py:58-66 and :13-17) ----------------------------------
// The three log-chroma differences generate all six (u,v) coordinates of the three channel
// histograms (histogram.py:72-74):  R:(d_rg,d_rb)  G:(-d_rg,d_gb)  B:(-d_rb,-d_gb).
struct PixelTerms {
  float x0, x1, x2;  // image*0.5+0.5
  float iy;          // sqrt(x0^2+x1^2+x2^2+eps)
  // log(x_a+eps)-log(x_b+eps) for (a,b) = (r,g), (r,b), (g,b), each as an unevaluated float pair
  // hi+lo.  The bin kernel is 1/(1+(u-c)^2/sigma^2) with sigma = 0.02: an absolute error of 1e-7 in
  // u (what float32 logs give, and what the reference's own arithmetic has) is a 1e-5 relative
  // error in dK/du near the bin centre.  The three logs are therefore taken in float64 (3 per pixel
  // against >= 384 bin weights per pixel) and the residual is carried into every (u-c).
  float d_rg, d_rb, d_gb;
  float l_rg, l_rb, l_gb;
};

__device__ __forceinline__ void split_double(double v, float& hi, float& lo) {
  hi = (float)v;
  lo = (float)(v - (double)hi);
}

__device__ __forceinline__ PixelTerms pixel_terms(float r, float g, float b, float eps) {
  PixelTerms t;
  t.x0 = fmaf(r, 0.5f, 0.5f);
  t.x1 = fmaf(g, 0.5f, 0.5f);
  t.x2 = fmaf(b, 0.5f, 0.5f);
  t.iy = sqrtf(t.x0 * t.x0 + t.x1 * t.x1 + t.x2 * t.x2 + eps);
  const double e = (double)eps;
  const double l0 = log(fma((double)r, 0.5, 0.5) + e);
  const double l1 = log(fma((double)g, 0.5, 0.5) + e);
  const double l2 = log(fma((double)b, 0.5, 0.5) + e);
  split_double(l0 - l1, t.d_rg, t.l_rg);
  split_double(l0 - l2, t.d_rb, t.l_rb);
  split_double(l1 - l2, t.d_gb, t.l_gb);
  return t;
}

// (u,v) of output channel c (0=R,1=G,2=B) as hi+lo pairs — histogram.py:72-74.
__device__ __forceinline__ void channel_uv(const PixelTerms& t, int c, float& u, float& ul, float& v, float& vl) {
  if (c == 0) { u = t.d_rg; ul = t.l_rg; v = t.d_rb; vl = t.l_rb; }
  else if (c == 1) { u = -t.d_rg; ul = -t.l_rg; v = t.d_gb; vl = t.l_gb; }
  else { u = -t.d_rb; ul = -t.l_rb; v = -t.d_gb; vl = -t.l_gb; }
}

// Bin kernel (histogram.py:20-27): t = (x-c)^2/sigma^2; IQ: 1/(1+t); RBF: exp(-t).
template <int METHOD>
__device__ __forceinline__ float bin_weight(float d, float inv_sigma_sqr) {
  const float t = d * d * inv_sigma_sqr;
  if (METHOD == PH_METHOD_INVERSE_QUADRATIC) {
    return __frcp_rn(1.0f + t);
  } else {
    return expf(-t);
  }
}

// d weight / d x at distance d = x - c, given the weight w.
template <int METHOD>
__device__ __forceinline__ float bin_weight_grad(float d, float w, float inv_sigma_sqr) {
  const float s = -2.0f * d * inv_sigma_sqr;
  if (METHOD == PH_METHOD_INVERSE_QUADRATIC) {
    return s * w * w;
  } else {
    return s * w;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum, result valid in every thread. `scratch` >= 32 elements of T in shared memory.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  T r = (lane < nwarps) ? scratch[lane] : T(0);
  r = warp_sum(r);
  return r;
}

}  // namespace ph
