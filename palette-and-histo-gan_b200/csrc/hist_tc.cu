// RGB-uv histogram, tensor-core engine (tcgen05, sm_100a): the Ku^T.Kv contraction over pixels
// (histogram.py:29-30) as kind::tf32 MMAs with fp32 emulation by operand splitting
// (w = hi + lo, both tf32; hi.hi + hi.lo + lo.hi + lo.lo accumulated in fp32 in TMEM).
//
// Forward, 64 bins, one persistent CTA per SM, 21 warps:
//   warps 17-20 pixel terms: RGBA load, float64 log-chroma (hi+lo), intensity -> smem ring
//               (round-robin over 32-pixel rounds)
//   warps 0-7   A operand (u side, Iy-weighted) written straight into TMEM: lane = bin, the two
//               half-warps of a TMEM sub-partition hold the hi and the lo rows of the same 16 bins,
//               so M = 128 = 64 bins x {hi, lo}; warps w and w+4 share a sub-partition and take 16 of
//               the 32 pixels of a stage each;   also the epilogue (TMEM -> global) warps
//   warps 8-15  B operand (v side) hi|lo into shared memory, K-major no-swizzle core matrices
//   warp 16     one thread issues tcgen05.mma (M=128, N=64, K=8) twice per k-step: B_hi and B_lo,
//               accumulating all four cross terms into the same 64 TMEM columns per channel
// The operands never exist in global memory: they are generated from 16 B per pixel.
#include "common.cuh"
#include "hist_internal.cuh"
#include "tc_ptx.cuh"

namespace ph {

using namespace tc;

namespace fwdtc {

constexpr int BINS = 64;
constexpr int KB = 32;        // pixels per pipeline stage
constexpr int NS = 3;         // A/B operand stages
constexpr int A_WARPS = 8, B_WARPS = 8, PXW = 4;
constexpr int MMA_WARP = A_WARPS + B_WARPS;  // 16
constexpr int PX_WARP0 = MMA_WARP + 1;       // 17
constexpr int PR = 8;         // pixel-term ring slots
constexpr int THREADS = (PX_WARP0 + PXW) * 32;  // 21 warps
constexpr int TMEM_COLS = 512;
constexpr int D_COLS = 64;                 // per channel
constexpr int A_COL0 = 3 * D_COLS;         // 192
constexpr int A_STAGE_COLS = 3 * KB;       // 96
constexpr int B_CH_BYTES = KB * 128 * 4;   // 16384: [kq 0..7][n-group 0..15][n%8][k%4]
constexpr int B_STAGE_BYTES = 3 * B_CH_BYTES;
constexpr int B_KQ_BYTES = 16 * 128;       // 2048: one 4-pixel quad for all 128 rows
static_assert(A_COL0 + NS * A_STAGE_COLS <= TMEM_COLS, "TMEM budget");

struct PxSlot {
  float u_hi[3][KB], u_lo[3][KB], v_hi[3][KB], v_lo[3][KB], iy[KB];
};

struct Smem {
  alignas(128) unsigned char b[NS][B_STAGE_BYTES];
  PxSlot px[PR];
  float dom[BINS];
  alignas(8) uint64_t px_full[PR], px_empty[PR], ab_full[NS], ab_empty[NS], d_full, d_empty;
  uint32_t tmem_base;
};

struct Params {
  const float* image;
  const float* dom;
  float* partial;  // (B, splits, 3, 64, 64)
  int64_t npix;
  int channels;
  int splits;
  int64_t px_per_split;
  int64_t items;  // B * splits
  float inv_sigma_sqr;
  float eps;
};

template <int METHOD>
__device__ __forceinline__ float weight(float d, float inv_s2) {
  const float t = d * d;
  if (METHOD == PH_METHOD_INVERSE_QUADRATIC) return fast_rcp(fmaf(t, inv_s2, 1.0f));
  return __expf(-t * inv_s2);
}

template <int METHOD>
__global__ void __launch_bounds__(THREADS, 1) hist_fwd_tc_kernel(Params p) {
  // no-swizzle operand tiles need only 16 B alignment; keeping the pointer derived from the
  // __shared__ symbol (no integer round trip) lets ptxas emit LDS/STS instead of generic LD/ST
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < PR; ++i) { mbar_init(&S.px_full[i], 32); mbar_init(&S.px_empty[i], (A_WARPS + B_WARPS) * 32); }
    for (int i = 0; i < NS; ++i) { mbar_init(&S.ab_full[i], (A_WARPS + B_WARPS) * 32); mbar_init(&S.ab_empty[i], 1); }
    mbar_init(&S.d_full, 1);
    mbar_init(&S.d_empty, A_WARPS * 32);
    fence_mbar_init();
  }
  if (tid < BINS) S.dom[tid] = p.dom[tid];
  if (warp == MMA_WARP) tmem_alloc(&S.tmem_base, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = S.tmem_base;

  const int64_t first = blockIdx.x, step = gridDim.x;

  if (warp >= PX_WARP0) {
    // ===================== pixel terms (PXW warps, round-robin over rounds) =====================
    const int me = warp - PX_WARP0;
    uint32_t it = 0;
    for (int64_t w = first; w < p.items; w += step) {
      const int64_t b = w / p.splits, split = w % p.splits;
      const int64_t px0 = split * p.px_per_split;
      const int64_t px1 = min(px0 + p.px_per_split, p.npix);
      for (int64_t base = px0; base < px1; base += KB, ++it) {
        if ((int)(it % PXW) != me) continue;
        const int slot = it % PR;
        const int64_t px = base + lane;
        float r = 0.f, g = 0.f, bl = 0.f;
        const bool valid = px < px1;
        if (valid) {
          const float* src = p.image + (b * p.npix + px) * p.channels;
          if (p.channels == 4) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(src));
            r = q.x; g = q.y; bl = q.z;
          } else {
            r = __ldg(src); g = __ldg(src + 1); bl = __ldg(src + 2);
          }
        }
        const PixelTerms t = pixel_terms(r, g, bl, p.eps);
        mbar_wait(&S.px_empty[slot], ((it / PR) & 1) ^ 1);
        PxSlot& o = S.px[slot];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float u, ul, v, vl;
          channel_uv(t, c, u, ul, v, vl);
          o.u_hi[c][lane] = u; o.u_lo[c][lane] = ul; o.v_hi[c][lane] = v; o.v_lo[c][lane] = vl;
        }
        o.iy[lane] = valid ? t.iy : 0.f;  // masked pixels contribute nothing (A operand = 0)
        mbar_arrive(&S.px_full[slot]);
      }
    }
  } else if (warp < A_WARPS) {
    // ===================== A operand (TMEM) + epilogue =====================
    const int quad = warp & 3;                  // TMEM sub-partition of this warp
    const int sub = warp >> 2;                  // which 16 of the 32 pixels of a stage
    const int half = lane >> 4;                 // 0: hi rows, 1: lo rows
    const int bin = quad * 16 + (lane & 15);
    const float c_bin = S.dom[bin];
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int px_own = sub * 16 + half * 8;     // the 8 pixels this thread evaluates
    uint32_t it = 0, item_idx = 0;
    for (int64_t w = first; w < p.items; w += step, ++item_idx) {
      const int64_t b = w / p.splits, split = w % p.splits;
      const int64_t px0 = split * p.px_per_split;
      const int64_t px1 = min(px0 + p.px_per_split, p.npix);
      for (int64_t base = px0; base < px1; base += KB, ++it) {
        const int slot = it % PR, stage = it % NS;
        mbar_wait(&S.px_full[slot], (it / PR) & 1);
        mbar_wait(&S.ab_empty[stage], ((it / NS) & 1) ^ 1);
        tc_fence_after_sync();
        const PxSlot& in = S.px[slot];
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
          uint32_t out[16];
#pragma unroll
          for (int i4 = 0; i4 < 2; ++i4) {
            const float4 uh = *reinterpret_cast<const float4*>(&in.u_hi[c][px_own + i4 * 4]);
            const float4 ul = *reinterpret_cast<const float4*>(&in.u_lo[c][px_own + i4 * 4]);
            const float4 iy = *reinterpret_cast<const float4*>(&in.iy[px_own + i4 * 4]);
            const float wv[4] = {iy.x * weight<METHOD>((uh.x - c_bin) + ul.x, p.inv_sigma_sqr),
                                 iy.y * weight<METHOD>((uh.y - c_bin) + ul.y, p.inv_sigma_sqr),
                                 iy.z * weight<METHOD>((uh.z - c_bin) + ul.z, p.inv_sigma_sqr),
                                 iy.w * weight<METHOD>((uh.w - c_bin) + ul.w, p.inv_sigma_sqr)};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // split once, keep the part this row needs, send the other part to the partner row
              const uint32_t hi = tf32_hi(wv[e]);
              const uint32_t lo = __float_as_uint(wv[e] - __uint_as_float(hi));
              const uint32_t keep = half ? lo : hi;
              const uint32_t recv = __shfl_xor_sync(0xffffffffu, half ? hi : lo, 16);
              const int i = i4 * 4 + e;
              out[i] = half ? recv : keep;      // pixels sub*16 + 0..7  (evaluated by the hi half-warp)
              out[8 + i] = half ? keep : recv;  // pixels sub*16 + 8..15 (evaluated by the lo half-warp)
            }
          }
          tmem_st16(tmem + lane_addr + A_COL0 + stage * A_STAGE_COLS + c * KB + sub * 16, out);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        mbar_arrive(&S.px_empty[slot]);
        mbar_arrive(&S.ab_full[stage]);
      }
      // ---- epilogue: D (TMEM) -> partial histogram of this work item; warps w / w+4 take 32 columns each ----
      mbar_wait(&S.d_full, item_idx & 1);
      tc_fence_after_sync();
      float* dst = p.partial + ((b * p.splits + split) * 3) * (int64_t)(BINS * BINS);
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + c * D_COLS + sub * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float mine = __uint_as_float(v[i]);
          f[i] = mine + __shfl_xor_sync(0xffffffffu, mine, 16);  // hi-row sum + lo-row sum
        }
        if (half == 0) {
          float4* row = reinterpret_cast<float4*>(dst + (int64_t)c * BINS * BINS + bin * BINS + sub * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) row[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&S.d_empty);
    }
  } else if (warp < MMA_WARP) {
    // ===================== B operand (shared memory) =====================
    const int t = tid - A_WARPS * 32;
    const int j = t & 63, part = t >> 6;  // part: which 8 of the 32 pixels (two 4-pixel quads)
    const float c_bin = S.dom[j];
    const uint32_t row_off = (uint32_t)((j >> 3) * 128 + (j & 7) * 16);  // hi row j; lo row j + 64 is +1024
    uint32_t it = 0;
    for (int64_t w = first; w < p.items; w += step) {
      const int64_t split = w % p.splits;
      const int64_t px0 = split * p.px_per_split;
      const int64_t px1 = min(px0 + p.px_per_split, p.npix);
      for (int64_t base = px0; base < px1; base += KB, ++it) {
        const int slot = it % PR, stage = it % NS;
        mbar_wait(&S.px_full[slot], (it / PR) & 1);
        mbar_wait(&S.ab_empty[stage], ((it / NS) & 1) ^ 1);
        const PxSlot& in = S.px[slot];
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
          unsigned char* tile = &S.b[stage][c * B_CH_BYTES];
#pragma unroll
          for (int q4 = 0; q4 < 2; ++q4) {
            const int kq = part * 2 + q4;
            const float4 vh = *reinterpret_cast<const float4*>(&in.v_hi[c][kq * 4]);
            const float4 vl = *reinterpret_cast<const float4*>(&in.v_lo[c][kq * 4]);
            const float w0 = weight<METHOD>((vh.x - c_bin) + vl.x, p.inv_sigma_sqr);
            const float w1 = weight<METHOD>((vh.y - c_bin) + vl.y, p.inv_sigma_sqr);
            const float w2 = weight<METHOD>((vh.z - c_bin) + vl.z, p.inv_sigma_sqr);
            const float w3 = weight<METHOD>((vh.w - c_bin) + vl.w, p.inv_sigma_sqr);
            uint4 hi, lo;
            hi.x = tf32_hi(w0); hi.y = tf32_hi(w1); hi.z = tf32_hi(w2); hi.w = tf32_hi(w3);
            lo.x = tf32_lo(w0); lo.y = tf32_lo(w1); lo.z = tf32_lo(w2); lo.w = tf32_lo(w3);
            *reinterpret_cast<uint4*>(tile + kq * B_KQ_BYTES + row_off) = hi;
            *reinterpret_cast<uint4*>(tile + kq * B_KQ_BYTES + row_off + 1024) = lo;
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&S.px_empty[slot]);
        mbar_arrive(&S.ab_full[stage]);
      }
    }
  } else if (warp == MMA_WARP && lane == 0) {
    // ===================== MMA issue =====================
    constexpr uint32_t IDESC = idesc_tf32(128, 64);
    const uint32_t b_base = smem_u32(&S.b[0][0]);
    uint32_t it = 0, item_idx = 0;
    for (int64_t w = first; w < p.items; w += step, ++item_idx) {
      const int64_t split = w % p.splits;
      const int64_t px0 = split * p.px_per_split;
      const int64_t px1 = min(px0 + p.px_per_split, p.npix);
      if (item_idx > 0) {
        mbar_wait(&S.d_empty, (item_idx - 1) & 1);
        tc_fence_after_sync();
      }
      bool first_kb = true;
      for (int64_t base = px0; base < px1; base += KB, ++it) {
        const int stage = it % NS;
        mbar_wait(&S.ab_full[stage], (it / NS) & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const uint32_t d_addr = tmem + c * D_COLS;
          const uint32_t a_addr = tmem + A_COL0 + stage * A_STAGE_COLS + c * KB;
          const uint32_t b_addr = b_base + stage * B_STAGE_BYTES + c * B_CH_BYTES;
#pragma unroll
          for (int ks = 0; ks < KB / 8; ++ks) {
            const uint64_t desc_hi = smem_desc_kmajor_noswizzle(b_addr + ks * 2 * B_KQ_BYTES, B_KQ_BYTES, 128);
            const uint64_t desc_lo = smem_desc_kmajor_noswizzle(b_addr + ks * 2 * B_KQ_BYTES + 1024, B_KQ_BYTES, 128);
            mma_tf32_ts(d_addr, a_addr + ks * 8, desc_hi, IDESC, (first_kb && ks == 0) ? 0u : 1u);
            mma_tf32_ts(d_addr, a_addr + ks * 8, desc_lo, IDESC, 1u);
          }
        }
        mma_commit(&S.ab_empty[stage]);
        first_kb = false;
      }
      mma_commit(&S.d_full);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace fwdtc

// =============================================================================================
// host side
// =============================================================================================
bool tc_supported(int64_t npix, int bins, int method) {
  (void)method;
  return bins == 64 && npix >= 1;
}

// Pixel slices per image.  Two constraints: enough work items to fill the SMs, and at most
// MAX_CHAIN_PX pixels accumulated in one TMEM accumulator: the tensor core adds each K=8 product
// block into the fp32 accumulator with truncation, so the error grows with the number of chained
// MMAs (measured: 4096 px in one chain -> 8e-6 relative, 1024 px -> ~1e-6); slices are summed in
// fp32 by the finalise kernel.
constexpr int64_t MAX_CHAIN_PX = 1024;
static int tc_fwd_splits(int64_t batch, int64_t npix) {
  const int64_t target = (int64_t)cached_sm_count() * 2;
  int64_t s = ceil_div(target, batch);
  const int64_t max_s = ceil_div(npix, 8 * fwdtc::KB);
  if (s > max_s) s = max_s;
  const int64_t min_s = ceil_div(npix, MAX_CHAIN_PX);
  if (s < min_s) s = min_s;
  if (s < 1) s = 1;
  return (int)s;
}

size_t tc_workspace_bytes(int64_t batch, int64_t npix, int bins) {
  if (bins != 64) return 0;
  const size_t fwd = (size_t)batch * tc_fwd_splits(batch, npix) * 3 * bins * bins * sizeof(float);
  const size_t bwd = (size_t)batch * 3 * bins * bins * sizeof(float);
  return align_up(fwd > bwd ? fwd : bwd, 256) + 256;
}

// defined in hist_simt.cu
void launch_finalize(const float* partial, int splits, int nch, int bins, int normalise, float* hist,
                     float* denom, int64_t batch, cudaStream_t st);

int tc_hist_forward(const float* image, int64_t batch, int64_t npix, int channels, const float* dom, int bins,
                    int method, float sigma_sqr, float eps, float* hist, float* denom, void* workspace,
                    cudaStream_t st) {
  using namespace fwdtc;
  PH_CHECK_ARG(bins == BINS, "tensor-core forward is specialised for 64 bins");
  Params p{};
  p.image = image;
  p.dom = dom;
  p.partial = static_cast<float*>(workspace);
  p.npix = npix;
  p.channels = channels;
  p.splits = tc_fwd_splits(batch, npix);
  p.px_per_split = ceil_div(ceil_div(npix, p.splits), KB) * KB;
  p.items = batch * p.splits;
  p.inv_sigma_sqr = 1.0f / sigma_sqr;
  p.eps = eps;
  const size_t smem = sizeof(Smem);
  int grid = cached_sm_count();
  if (grid > p.items) grid = (int)p.items;
  if (method == PH_METHOD_INVERSE_QUADRATIC) {
    PH_CUDA_OK(cudaFuncSetAttribute(hist_fwd_tc_kernel<PH_METHOD_INVERSE_QUADRATIC>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hist_fwd_tc_kernel<PH_METHOD_INVERSE_QUADRATIC><<<grid, THREADS, smem, st>>>(p);
  } else {
    PH_CUDA_OK(cudaFuncSetAttribute(hist_fwd_tc_kernel<PH_METHOD_RBF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    hist_fwd_tc_kernel<PH_METHOD_RBF><<<grid, THREADS, smem, st>>>(p);
  }
  PH_LAUNCH_OK("hist_fwd_tc_kernel");
  launch_finalize(p.partial, p.splits, 3, bins, 1, hist, denom, batch, st);
  PH_LAUNCH_OK("hist_finalize_kernel");
  return PH_OK;
}

}  // namespace ph
