// RGB-uv histogram, tensor-core engine (tcgen05, sm_100a): the Ku^T.Kv contraction over pixels
// (histogram.py:29-30) as kind::tf32 MMAs with fp32 emulation by operand splitting
// (w = hi + lo, both tf32; hi.hi + hi.lo + lo.hi + lo.lo accumulated in fp32 in TMEM), fused with the
// per-image normalisation (histogram.py:75-79).
//
// Forward, 64 bins, one persistent CTA per SM, 20 warps:
//   warps 17-19 pixel pass: 128-bit RGBA loads, log-chroma u/v per channel and intensity Iy -> smem ring
//   warps 0-7   A operand (u side, Iy-weighted) written straight into TMEM, M = 128 rows =
//               64 bins x {hi, lo}: TMEM sub-partitions 0,1 hold the hi rows of bins 0-31 / 32-63,
//               sub-partitions 2,3 the lo rows, so a warp's role is uniform.  The hi warp and the lo
//               warp of the same bins evaluate 8 pixels each and swap halves through shared memory.
//               The same warps drain the accumulators (epilogue).
//   warps 8-15  B operand (v side) hi|lo into shared memory, K-major no-swizzle core matrices
//   warp 16     one thread issues tcgen05.mma (M=128, N=64, K=8) twice per k-step: B_hi and B_lo,
//               accumulating all four cross terms into the same 64 TMEM columns per channel
// The bin weights never exist in global memory: they are generated from 16 B per pixel with packed
// fp32x2 arithmetic (FADD2/FMUL2/FFMA2) and one MUFU.RCP per weight.
//
// Accuracy: the tensor core adds every K=8 block into the fp32 accumulator with truncation, so the
// error grows with the chain length (measured 8e-6 for 4096 pixels in one chain, ~1e-6 for 1024).
// Chains are cut at 1024 pixels and summed in fp32 in a shared-memory accumulator; when a CTA owns a
// whole image the normaliser D and H/D are produced in the same kernel.
#include <stdlib.h>

#include "common.cuh"
#include "hist_internal.cuh"
#include "tc_ptx.cuh"

namespace ph {

using namespace tc;

namespace fwdtc {

constexpr int BINS = 64;
constexpr int KB = 32;         // pixels per pipeline stage
constexpr int NS = 3;          // A/B operand stages
constexpr int CHAIN_KB = 32;   // stages per TMEM accumulation chain (1024 pixels)
constexpr int A_WARPS = 8, B_WARPS = 8, PXW = 3;  // 20 warps: 640 threads leave 96 registers per thread
constexpr int MMA_WARP = A_WARPS + B_WARPS;     // 16
constexpr int PX_WARP0 = MMA_WARP + 1;          // 17
constexpr int PR = 6;                           // pixel ring slots
constexpr int DEDUP_MAX = 512;                  // unique colours kept per image by the de-duplication pass
constexpr int DEDUP_SLOTS = 1024;
constexpr int THREADS = (PX_WARP0 + PXW) * 32;  // 672
constexpr int TMEM_COLS = 512;
constexpr int D_COLS = 64;                 // per channel
constexpr int A_COL0 = 3 * D_COLS;         // 192
constexpr int A_STAGE_COLS = 3 * KB;       // 96
constexpr int B_CH_BYTES = KB * 128 * 4;   // 16384: [kq 0..7][n-group 0..15][n%8][k%4]
constexpr int B_STAGE_BYTES = 3 * B_CH_BYTES;
constexpr int B_KQ_BYTES = 16 * 128;       // 2048: one 4-pixel quad for all 128 rows
static_assert(A_COL0 + NS * A_STAGE_COLS <= TMEM_COLS, "TMEM budget");
static_assert(KB == 32, "stage = 32 pixels (shifts below)");

struct PxSlot {
  float u[3][KB], v[3][KB], iy[KB];
};

struct Smem {
  alignas(128) unsigned char b[NS][B_STAGE_BYTES];  // 144 KB
  float acc[3][BINS][BINS + 1];                     // [c][j][i] running fp32 sum over chains; rows padded to 65
                                                    // floats so both the bin-major drain and the j-major
                                                    // write-out are free of bank conflicts
  float4 xbuf[4][2][3][2][32];                      // [pair][direction][channel][quad of 4 px][lane], 24 KB
  PxSlot px[PR];
  float dom[BINS];
  float red[8];
  alignas(8) uint64_t px_full[PR], px_empty[PR], ab_full[NS], ab_empty[NS], d_full, d_empty;
  uint32_t tmem_base;
};

struct Params {
  const float* image;
  const float* dom;
  float* partial;  // (B - n_whole, splits, 3, 64, 64) raw sums of the sliced ("tail") images
  float* hist;     // (B, 64, 64, 3) normalised, written directly for whole-image items
  float* denom;    // (B)
  const float4* ulist;  // optional (B, DEDUP_MAX): unique colours (r,g,b,count) of each image, or NULL
  const int* nunique;   // optional (B): number of unique colours, < 0 = image not de-duplicated
  int64_t npix;
  int channels;
  // Work items: images [0, n_whole) are contracted whole by one CTA (normalisation fused); the remaining
  // "tail" images are cut into `splits` pixel slices so that the last, partial wave of images does not leave
  // most SMs idle (their raw sums go to `partial` and are normalised by the finalise kernel).
  int64_t n_whole;
  int splits;
  int64_t px_per_split;
  int64_t items;  // n_whole + (B - n_whole) * splits
  float inv_sigma_sqr;
  float eps;
  int debug_skip_mma;  // tuning experiments only (env PH_DEBUG_SKIP_MMA): bit 0 = issue no MMA, bit 1 = B warps
                       // generate nothing.  Result (B200): skipping all MMAs does not shorten the kernel and
                       // skipping B saves 20 %: at 64 bins the A-operand chain (px wait -> weights -> hand-over
                       // barrier -> tcgen05.st -> wait::st -> arrive) and shared-memory bandwidth (145 KB per
                       // 32-pixel stage = 1 130 cycles) bound the forward, not the tensor pipe.
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Pixel range of a work item.  A de-duplicated image is a list of (colour, multiplicity) entries: the
// histogram is linear in the pixels, so identical pixels are contracted once with weight count*Iy.
struct ItemRange {  // 32-bit pixel range: keeps the role loops in the uniform datapath
  int64_t b;        // image
  int64_t pidx;     // slice index in `partial` (tail items)
  uint32_t px0, px1;
  bool dedup, whole;
};
__device__ __forceinline__ ItemRange item_range(const Params& p, int64_t w) {
  ItemRange r;
  r.dedup = false;
  if (w < p.n_whole) {
    r.b = w; r.pidx = 0; r.whole = true;
    r.px0 = 0; r.px1 = (uint32_t)p.npix;
    if (p.nunique != nullptr) {
      // broadcast from lane 0: lets ptxas prove the loop bounds derived from it warp-uniform
      const int nu = __shfl_sync(0xffffffffu, __ldg(p.nunique + w), 0);
      if (nu >= 0) { r.px1 = (uint32_t)nu; r.dedup = true; }
    }
    return r;
  }
  const int64_t t = w - p.n_whole;
  const int64_t split = t % p.splits;
  r.b = p.n_whole + t / p.splits; r.pidx = t; r.whole = false;
  r.px0 = (uint32_t)split * (uint32_t)p.px_per_split;
  r.px1 = min(r.px0 + (uint32_t)p.px_per_split, (uint32_t)p.npix);
  return r;
}

// two bin weights at once: d = x + (-c);  IQ: 1/(1 + d^2/s^2), RBF: exp(-d^2/s^2)
template <int METHOD>
__device__ __forceinline__ f32x2 weight2(f32x2 x, f32x2 negc, f32x2 inv2, f32x2 one2) {
  const f32x2 d = add2(x, negc);
  const f32x2 t = mul2(d, d);
  if (METHOD == PH_METHOD_INVERSE_QUADRATIC) {
    const f32x2 e = fma2(t, inv2, one2);
    return pack2(fast_rcp(lo_of(e)), fast_rcp(hi_of(e)));
  } else {
    const f32x2 e = mul2(t, inv2);
    return pack2(__expf(-lo_of(e)), __expf(-hi_of(e)));
  }
}

// A-operand producer + epilogue warp.  ROLE (0: hi rows, 1: lo rows) is a template parameter so the
// keep/hand-over choices compile to plain register assignments.
template <int METHOD, int ROLE>
__device__ __forceinline__ void a_warp_loop(Smem& S, const Params& p, uint32_t tmem, int tid, int warp, int lane,
                                            int64_t first, int64_t step, f32x2 inv2, f32x2 one2, f32x2 mone2) {
    // ===================== A operand (TMEM) + epilogue =====================
    const int quad = warp & 3;        // TMEM sub-partition of this warp
    const int sub = warp >> 2;        // which 16 of the 32 pixels of a stage
    constexpr int role = ROLE;        // 0: hi rows, 1: lo rows (warp-uniform, compile-time here)
    const int bin = (quad & 1) * 32 + lane;
    const int pair = (quad & 1) * 2 + sub;            // the two warps sharing (bins, pixel half)
    const float c_bin = S.dom[bin];
    const f32x2 negc = pack2(-c_bin, -c_bin);
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int px_own = sub * 16 + role * 8;  // the 8 pixels this warp evaluates
    uint32_t it = 0, chain = 0;
    for (int64_t w = first; w < p.items; w += step) {
      const ItemRange ir = item_range(p, w);
      const int64_t b = ir.b;
      const uint32_t nkb = (ir.px1 - ir.px0 + KB - 1) >> 5;
      for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
        const int slot = it % PR, stage = it % NS;
        mbar_wait(&S.px_full[slot], (it / PR) & 1);
        const PxSlot& in = S.px[slot];
        // all three channels at once: 24 weights per thread in flight, one hand-over with the partner warp
        // (barrier 1: partner has consumed the previous stage's hand-over; barrier 2: this stage's is written)
        f32x2 keep[3][4];
        named_bar_sync(1 + pair, 64);
        ulonglong2* xs = reinterpret_cast<ulonglong2*>(&S.xbuf[pair][role][0][0][lane]);
        const ulonglong2* xr = reinterpret_cast<const ulonglong2*>(&S.xbuf[pair][role ^ 1][0][0][lane]);
        const ulonglong2 ia = *reinterpret_cast<const ulonglong2*>(&in.iy[px_own]);
        const ulonglong2 ib = *reinterpret_cast<const ulonglong2*>(&in.iy[px_own + 4]);
        ulonglong2 uu[3][2];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uu[c][0] = *reinterpret_cast<const ulonglong2*>(&in.u[c][px_own]);
          uu[c][1] = *reinterpret_cast<const ulonglong2*>(&in.u[c][px_own + 4]);
        }
        mbar_arrive_warp(&S.px_empty[slot]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const ulonglong2 ua = uu[c][0], ub = uu[c][1];
          const f32x2 w0 = mul2(weight2<METHOD>(ua.x, negc, inv2, one2), ia.x);
          const f32x2 w1 = mul2(weight2<METHOD>(ua.y, negc, inv2, one2), ia.y);
          const f32x2 w2 = mul2(weight2<METHOD>(ub.x, negc, inv2, one2), ib.x);
          const f32x2 w3 = mul2(weight2<METHOD>(ub.y, negc, inv2, one2), ib.y);
          const f32x2 h0 = w0 & TF32_MASK2, h1 = w1 & TF32_MASK2, h2 = w2 & TF32_MASK2, h3 = w3 & TF32_MASK2;
          const f32x2 l0 = fma2(h0, mone2, w0), l1 = fma2(h1, mone2, w1), l2 = fma2(h2, mone2, w2),
                      l3 = fma2(h3, mone2, w3);
          // keep the part this row set needs, hand the other part to the partner warp
          if (role == 0) {
            keep[c][0] = h0; keep[c][1] = h1; keep[c][2] = h2; keep[c][3] = h3;
            xs[c * 64] = make_ulonglong2(l0, l1); xs[c * 64 + 32] = make_ulonglong2(l2, l3);
          } else {
            keep[c][0] = l0; keep[c][1] = l1; keep[c][2] = l2; keep[c][3] = l3;
            xs[c * 64] = make_ulonglong2(h0, h1); xs[c * 64 + 32] = make_ulonglong2(h2, h3);
          }
        }
        named_bar_sync(1 + pair, 64);
        mbar_wait(&S.ab_empty[stage], ((it / NS) & 1) ^ 1);  // the MMAs that read this TMEM stage are done
        tc_fence_after_sync();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const ulonglong2 ra = xr[c * 64], rb = xr[c * 64 + 32];
          const f32x2 got[4] = {ra.x, ra.y, rb.x, rb.y};
          uint32_t out[16];
          // columns sub*16 + 0..7 are the pixels of the hi warp, + 8..15 those of the lo warp
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const f32x2 first8 = role == 0 ? keep[c][e] : got[e];
            const f32x2 second8 = role == 0 ? got[e] : keep[c][e];
            out[2 * e] = (uint32_t)first8;
            out[2 * e + 1] = (uint32_t)(first8 >> 32);
            out[8 + 2 * e] = (uint32_t)second8;
            out[8 + 2 * e + 1] = (uint32_t)(second8 >> 32);
          }
          tmem_st16(tmem + lane_addr + A_COL0 + stage * A_STAGE_COLS + c * KB + sub * 16, out);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        mbar_arrive_warp(&S.ab_full[stage]);

        const bool chain_end = ((kb + 1) % CHAIN_KB == 0) || (kb + 1 == nkb);
        if (!chain_end) continue;
        // ---- chain epilogue: D (TMEM) += into the fp32 shared-memory accumulator ----
        mbar_wait(&S.d_full, chain & 1);
        ++chain;
        tc_fence_after_sync();
        // hi-row warps first (plain stores on the first chain of an item, else read-add-write), then the
        // lo-row warps add their share: no shared-memory float atomics (those are CAS loops)
        const bool first_chain = kb < CHAIN_KB;
        if (role == 0) {
#pragma unroll 1
          for (int c = 0; c < 3; ++c) {
            uint32_t v[32];
            tmem_ld32(tmem + lane_addr + c * D_COLS + sub * 32, v);
            tmem_ld_wait();
            if (first_chain) {
#pragma unroll
              for (int i = 0; i < 32; ++i) S.acc[c][sub * 32 + i][bin] = __uint_as_float(v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) S.acc[c][sub * 32 + i][bin] += __uint_as_float(v[i]);
            }
          }
        }
        named_bar_sync(5, A_WARPS * 32);
        if (role == 1) {
#pragma unroll 1
          for (int c = 0; c < 3; ++c) {
            uint32_t v[32];
            tmem_ld32(tmem + lane_addr + c * D_COLS + sub * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) S.acc[c][sub * 32 + i][bin] += __uint_as_float(v[i]);
          }
        }
        tc_fence_before_sync();
        mbar_arrive_warp(&S.d_empty);
        if (kb + 1 != nkb) continue;

        // ---- item epilogue: all 8 warps, normalise (whole image) or emit the raw partial ----
        named_bar_sync(5, A_WARPS * 32);
        const int t = tid;  // 0..255
        if (ir.whole) {
          float s = 0.f;
          for (int e = t; e < 3 * BINS * BINS; e += A_WARPS * 32) s += S.acc[e >> 12][(e >> 6) & 63][e & 63];
          s = warp_sum(s);
          if (lane == 0) S.red[warp] = s;
          named_bar_sync(5, A_WARPS * 32);
          float d = 0.f;
#pragma unroll
          for (int k = 0; k < A_WARPS; ++k) d += S.red[k];
          if (t == 0) p.denom[b] = d;
          const float inv_d = 1.0f / d;
          float* dst = p.hist + b * (int64_t)(3 * BINS * BINS);
          for (int e = t; e < 3 * BINS * BINS; e += A_WARPS * 32) {
            const int c = e % 3, ij = e / 3, i = ij >> 6, j = ij & 63;
            dst[e] = S.acc[c][j][i] * inv_d;
          }
        } else {
          float* dst = p.partial + ir.pidx * (int64_t)(3 * BINS * BINS);
          for (int e = t; e < 3 * BINS * BINS; e += A_WARPS * 32) {
            const int c = e >> 12, i = (e >> 6) & 63, j = e & 63;
            dst[e] = S.acc[c][j][i];
          }
        }
        named_bar_sync(5, A_WARPS * 32);
      }
    }
}

template <int METHOD>
__global__ void __launch_bounds__(THREADS, 1) hist_fwd_tc_kernel(Params p) {
  // no-swizzle operand tiles need only 16 B alignment; keeping the pointer derived from the
  // __shared__ symbol (no integer round trip) lets ptxas emit LDS/STS instead of generic LD/ST
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;

  if (tid == 0) {
    // producer/consumer barriers count WARPS (mbar_arrive_warp), the tcgen05.commit ones count 1
    for (int i = 0; i < PR; ++i) { mbar_init(&S.px_full[i], 1); mbar_init(&S.px_empty[i], A_WARPS + B_WARPS); }
    for (int i = 0; i < NS; ++i) { mbar_init(&S.ab_full[i], A_WARPS + B_WARPS); mbar_init(&S.ab_empty[i], 1); }
    mbar_init(&S.d_full, 1);
    mbar_init(&S.d_empty, A_WARPS);
    fence_mbar_init();
  }
  if (tid < BINS) S.dom[tid] = p.dom[tid];
  if (warp == MMA_WARP) tmem_alloc(&S.tmem_base, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = S.tmem_base;

  const int64_t first = blockIdx.x, step = gridDim.x;
  const f32x2 inv2 = pack2(p.inv_sigma_sqr, p.inv_sigma_sqr);
  const f32x2 one2 = pack2(1.0f, 1.0f);
  const f32x2 mone2 = pack2(-1.0f, -1.0f);

  if (warp >= PX_WARP0) {
    // ===================== pixel pass (PXW warps, round-robin over 32-pixel rounds) =====================
    const int me = warp - PX_WARP0;
    uint32_t it = 0;
    for (int64_t w = first; w < p.items; w += step) {
      const ItemRange ir = item_range(p, w);
      const int64_t b = ir.b;
      const uint32_t px0 = ir.px0, px1 = ir.px1;
      for (uint32_t base = px0; base < px1; base += KB, ++it) {
        if ((int)(it % PXW) != me) continue;
        const int slot = it % PR;
        const uint32_t px = base + lane;
        float r = 0.f, g = 0.f, bl = 0.f, mult = 1.f;
        const bool valid = px < px1;
        if (valid && ir.dedup) {
          const float4 q = __ldg(p.ulist + b * DEDUP_MAX + px);
          r = q.x; g = q.y; bl = q.z; mult = q.w;
        } else if (valid) {
          const float* src = p.image + (b * p.npix + px) * p.channels;
          if (p.channels == 4) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(src));
            r = q.x; g = q.y; bl = q.z;
          } else {
            r = __ldg(src); g = __ldg(src + 1); bl = __ldg(src + 2);
          }
        }
        // histogram.py:58-66, :13-17 — log of the ratio: one rounding instead of two at magnitude 13.8
        const float x0 = fmaf(r, 0.5f, 0.5f), x1 = fmaf(g, 0.5f, 0.5f), x2 = fmaf(bl, 0.5f, 0.5f);
        const float iy = sqrtf(x0 * x0 + x1 * x1 + x2 * x2 + p.eps);
        const float e0 = x0 + p.eps, e1 = x1 + p.eps, e2 = x2 + p.eps;
        const float d_rg = logf(e0 / e1), d_rb = logf(e0 / e2), d_gb = logf(e1 / e2);
        mbar_wait(&S.px_empty[slot], ((it / PR) & 1) ^ 1);
        PxSlot& o = S.px[slot];
        // (u,v): R:(rg, rb)  G:(-rg, gb)  B:(-rb, -gb)   (histogram.py:72-74)
        o.u[0][lane] = d_rg;  o.v[0][lane] = d_rb;
        o.u[1][lane] = -d_rg; o.v[1][lane] = d_gb;
        o.u[2][lane] = -d_rb; o.v[2][lane] = -d_gb;
        o.iy[lane] = valid ? iy * mult : 0.f;  // masked pixels contribute nothing (A operand = 0)
        mbar_arrive_warp(&S.px_full[slot]);
      }
    }
  } else if (warp < A_WARPS) {
    if (((warp & 3) >> 1) == 0) a_warp_loop<METHOD, 0>(S, p, tmem, tid, warp, lane, first, step, inv2, one2, mone2);
    else a_warp_loop<METHOD, 1>(S, p, tmem, tid, warp, lane, first, step, inv2, one2, mone2);
  } else if (warp < MMA_WARP) {
    // ===================== B operand (shared memory) =====================
    const int t = tid - A_WARPS * 32;
    const int j = t & 63, part = t >> 6;  // part: which 8 of the 32 pixels (two 4-pixel quads)
    const float c_bin = S.dom[j];
    const f32x2 negc = pack2(-c_bin, -c_bin);
    const uint32_t row_off = (uint32_t)((j >> 3) * 128 + (j & 7) * 16);  // hi row j; lo row j + 64 is +1024
    uint32_t it = 0;
    for (int64_t w = first; w < p.items; w += step) {
      const ItemRange ir = item_range(p, w);
      for (uint32_t base = ir.px0; base < ir.px1; base += KB, ++it) {
        const int slot = it % PR, stage = it % NS;
        mbar_wait(&S.px_full[slot], (it / PR) & 1);
        const PxSlot& in = S.px[slot];
        // all loads first (the compiler cannot hoist them across the shared-memory stores below)
        ulonglong2 vv[3][2];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int q4 = 0; q4 < 2; ++q4) vv[c][q4] = *reinterpret_cast<const ulonglong2*>(&in.v[c][(part * 2 + q4) * 4]);
        mbar_arrive_warp(&S.px_empty[slot]);
        mbar_wait(&S.ab_empty[stage], ((it / NS) & 1) ^ 1);
        if (!(p.debug_skip_mma & 2))
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          unsigned char* tile = &S.b[stage][c * B_CH_BYTES];
#pragma unroll
          for (int q4 = 0; q4 < 2; ++q4) {
            const int kq = part * 2 + q4;
            const f32x2 w0 = weight2<METHOD>(vv[c][q4].x, negc, inv2, one2);
            const f32x2 w1 = weight2<METHOD>(vv[c][q4].y, negc, inv2, one2);
            const f32x2 h0 = w0 & TF32_MASK2, h1 = w1 & TF32_MASK2;
            const f32x2 l0 = fma2(h0, mone2, w0), l1 = fma2(h1, mone2, w1);
            *reinterpret_cast<ulonglong2*>(tile + kq * B_KQ_BYTES + row_off) = make_ulonglong2(h0, h1);
            *reinterpret_cast<ulonglong2*>(tile + kq * B_KQ_BYTES + row_off + 1024) = make_ulonglong2(l0, l1);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive_warp(&S.ab_full[stage]);
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issue: the whole warp runs the (uniform) loop, one elected lane issues ====
    constexpr uint32_t IDESC = idesc_tf32(128, 64);
    const uint64_t desc0 = smem_desc_kmajor_noswizzle(smem_u32(&S.b[0][0]), B_KQ_BYTES, 128);
    const uint32_t dlo0 = (uint32_t)desc0, dhi = (uint32_t)(desc0 >> 32);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);  // provably uniform copy
    // Wrap-around counters only (no % or /) and no conditionally executed waits: ptxas then keeps the whole
    // loop in the uniform datapath and the MMA operands in uniform registers (a conditional mbarrier wait
    // inside the stage loop was enough to push every operand through R2UR moves).
    uint32_t stage = 0, phase = 0, chain_par = 0;
    for (int64_t w = first; w < p.items; w += step) {
      const ItemRange ir = item_range(p, w);
      const uint32_t nkb = (ir.px1 - ir.px0 + KB - 1) >> 5;
      for (uint32_t kb0 = 0; kb0 < nkb; kb0 += CHAIN_KB) {
        const uint32_t n_this = min((uint32_t)CHAIN_KB, nkb - kb0);
        for (uint32_t k = 0; k < n_this; ++k) {
          mbar_wait(&S.ab_full[stage], phase);
          tc_fence_after_sync();
          // the start-address field (bits 0-13, units of 16 B) never carries into the next field
          const uint32_t dstage = dlo0 + stage * (B_STAGE_BYTES >> 4);
          const uint32_t a_stage = tm + A_COL0 + stage * A_STAGE_COLS;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int ks = 0; ks < KB / 8; ++ks) {
              const uint32_t b_hi = dstage + ((c * B_CH_BYTES + ks * 2 * B_KQ_BYTES) >> 4);
              const uint32_t b_lo = b_hi + (1024 >> 4);
              const uint32_t acc0 = (k == 0 && ks == 0) ? 0u : 1u;
              if (!(p.debug_skip_mma & 1) && elect_one_sync()) {
                mma_tf32_ts2(tm + c * D_COLS, a_stage + c * KB + ks * 8, b_hi, dhi, IDESC, acc0);
                mma_tf32_ts2(tm + c * D_COLS, a_stage + c * KB + ks * 8, b_lo, dhi, IDESC, 1u);
              }
            }
          }
          if (elect_one_sync()) mma_commit(&S.ab_empty[stage]);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) mma_commit(&S.d_full);
        // the accumulators are free again once the epilogue warps have drained this chain
        mbar_wait(&S.d_empty, chain_par);
        tc_fence_after_sync();
        chain_par ^= 1;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace fwdtc


// =============================================================================================
// De-duplication pass: unique RGB triples of each image with their multiplicities.  The histogram is
// a sum over pixels of a function of the pixel's colour, so identical pixels contribute
// count x (one pixel): palette images (the reference's `real` sprites have 10-54 colours)
// contract 1-2 stages instead of 128.  One CTA per image; an open-addressing table in shared memory
// keyed by the bit pattern of (r,g,b): a slot is claimed by the index of the first pixel that hashes
// there (32-bit CAS) and later pixels compare their colour with that pixel's.  Lanes of a warp that
// hold the same colour are aggregated first (match + ballot), so the fully transparent background
// costs one table operation per warp.  More than DEDUP_MAX colours (a generator output): the image
// is flagged dense (nunique = -1) and the contraction reads its pixels directly.
// =============================================================================================
__global__ void __launch_bounds__(256) hist_dedup_kernel(const float* __restrict__ image, int64_t npix, int channels,
                                                         float4* __restrict__ ulist, int* __restrict__ nunique) {
  using namespace fwdtc;
  __shared__ int owner[DEDUP_SLOTS];
  __shared__ int count[DEDUP_SLOTS];
  __shared__ int s_n, s_out;
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const float* img = image + b * npix * channels;
  for (int i = tid; i < DEDUP_SLOTS; i += 256) { owner[i] = -1; count[i] = 0; }
  if (tid == 0) { s_n = 0; s_out = 0; }
  __syncthreads();
  const int64_t padded = (npix + 31) / 32 * 32;
  for (int64_t px = tid; px < padded; px += 256) {
    const bool active = px < npix;
    unsigned r = 0, g = 0, bl = 0;
    if (active) {
      const float* src = img + px * channels;
      if (channels == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(src));
        r = __float_as_uint(q.x); g = __float_as_uint(q.y); bl = __float_as_uint(q.z);
      } else {
        r = __float_as_uint(__ldg(src)); g = __float_as_uint(__ldg(src + 1)); bl = __float_as_uint(__ldg(src + 2));
      }
    }
    const unsigned h = (r * 2654435761u) ^ (g * 2246822519u) ^ (bl * 3266489917u);
    const unsigned amask = __ballot_sync(0xffffffffu, active);
    const bool dense = __any_sync(0xffffffffu, *(volatile int*)&s_n > DEDUP_MAX);  // warp-uniform exit
    if (!active || dense) continue;
    // warp aggregation: lanes with the same hash, then verified against the group leader's colour
    const unsigned peers = __match_any_sync(amask, h);
    const int leader = __ffs(peers) - 1;
    const unsigned lr = __shfl_sync(peers, r, leader), lg = __shfl_sync(peers, g, leader), lb = __shfl_sync(peers, bl, leader);
    const bool same = (r == lr) && (g == lg) && (bl == lb);
    const unsigned agree = __ballot_sync(peers, same) & peers;
    int add = 0;
    if (lane == leader) add = __popc(agree);
    else if (!same) add = 1;  // hash collision inside the warp: insert on its own
    if (add == 0) continue;
    unsigned slot = (h >> 16 ^ h) & (DEDUP_SLOTS - 1);
    for (int probe = 0; probe < DEDUP_SLOTS; ++probe) {
      int o = *(volatile int*)&owner[slot];
      if (o == -1) {
        o = atomicCAS(&owner[slot], -1, (int)px);
        if (o == -1) { o = (int)px; atomicAdd(&s_n, 1); }
      }
      const float* op = img + (int64_t)o * channels;
      if (__float_as_uint(__ldg(op)) == r && __float_as_uint(__ldg(op + 1)) == g && __float_as_uint(__ldg(op + 2)) == bl) {
        atomicAdd(&count[slot], add);
        break;
      }
      slot = (slot + 1) & (DEDUP_SLOTS - 1);
      if (*(volatile int*)&s_n > DEDUP_MAX) break;
    }
  }
  __syncthreads();
  const int n = s_n;
  if (n > DEDUP_MAX) {
    if (tid == 0) nunique[b] = -1;
    return;
  }
  float4* out = ulist + b * DEDUP_MAX;
  for (int i = tid; i < DEDUP_SLOTS; i += 256) {
    const int o = owner[i];
    if (o >= 0) {
      const float* op = img + (int64_t)o * channels;
      out[atomicAdd(&s_out, 1)] = make_float4(__ldg(op), __ldg(op + 1), __ldg(op + 2), (float)count[i]);
    }
  }
  if (tid == 0) nunique[b] = n;
}

// =============================================================================================
// host side
// =============================================================================================
bool tc_supported(int64_t npix, int bins, int method) {
  (void)method;
  return bins == 64 && npix >= 1;
}

// Pixel slices per image: 1 (whole image per CTA, normalisation fused) once the batch fills the SMs,
// otherwise enough slices to occupy them; the slices are summed by the finalise kernel.
static size_t dedup_bytes(int64_t batch) {
  return align_up((size_t)batch * fwdtc::DEDUP_MAX * sizeof(float4), 256) + align_up((size_t)batch * sizeof(int), 256);
}

struct FwdPlan { int64_t n_whole; int splits; };

// Which images a CTA contracts whole (normalisation fused) and how the rest are sliced.
static FwdPlan tc_fwd_plan(int64_t batch, int64_t npix, bool dedup) {
  const int64_t sms = cached_sm_count();
  const int64_t max_s = ceil_div(npix, 8 * fwdtc::KB);
  FwdPlan pl{batch, 1};
  if (batch < sms) {  // few images: slice all of them to occupy the SMs
    int64_t s = ceil_div(2 * sms, batch);
    if (s > max_s) s = max_s;
    if (s > 1) { pl.n_whole = 0; pl.splits = (int)s; }
    return pl;
  }
  if (dedup) return pl;  // de-duplicated images are a stage or two each
  static const bool tail_off = getenv("PH_FWD_TAIL") && atoi(getenv("PH_FWD_TAIL")) == 0;  // tuning knob
  if (tail_off) return pl;
  const int64_t n_tail = batch % sms;
  if (n_tail == 0) return pl;
  // the last partial wave: n_tail whole images keep sms - n_tail SMs idle for one image time; slices shorten it
  int best = 1;
  double best_cost = 1.0;
  for (int sidx = 2; sidx <= 4 && sidx <= max_s; ++sidx) {
    const double cost = (double)ceil_div(n_tail * sidx, sms) / sidx + 0.08;  // + finalise pass and per-item overhead
    if (cost < best_cost) { best_cost = cost; best = sidx; }
  }
  if (best > 1) { pl.n_whole = batch - n_tail; pl.splits = best; }
  return pl;
}

size_t tc_workspace_bytes(int64_t batch, int64_t npix, int bins) {
  if (bins != 64) return 0;
  const FwdPlan pl = tc_fwd_plan(batch, npix, false);
  size_t fwd = (size_t)(batch - pl.n_whole) * pl.splits * 3 * bins * bins * sizeof(float);
  if (tc_fwd_plan(batch, npix, true).n_whole == batch && dedup_bytes(batch) > fwd) fwd = dedup_bytes(batch);
  const size_t bwd = tc_bwd_workspace_bytes(batch);
  return align_up(fwd > bwd ? fwd : bwd, 256) + 256;
}

// defined in hist_simt.cu
void launch_finalize(const float* partial, int splits, int nch, int bins, int normalise, float* hist,
                     float* denom, int64_t batch, cudaStream_t st);

int tc_hist_forward(const float* image, int64_t batch, int64_t npix, int channels, const float* dom, int bins,
                    int method, float sigma_sqr, float eps, float* hist, float* denom, void* workspace, bool dedup,
                    cudaStream_t st) {
  using namespace fwdtc;
  PH_CHECK_ARG(bins == BINS, "tensor-core forward is specialised for 64 bins");
  Params p{};
  p.image = image;
  p.dom = dom;
  p.partial = static_cast<float*>(workspace);
  p.hist = hist;
  p.denom = denom;
  p.npix = npix;
  p.channels = channels;
  const FwdPlan pl = tc_fwd_plan(batch, npix, dedup);
  p.n_whole = pl.n_whole;
  p.splits = pl.splits;
  p.px_per_split = ceil_div(ceil_div(npix, p.splits), KB) * KB;
  p.items = p.n_whole + (batch - p.n_whole) * p.splits;
  p.inv_sigma_sqr = 1.0f / sigma_sqr;
  p.eps = eps;
  static const int skip_mma = getenv("PH_DEBUG_SKIP_MMA") ? atoi(getenv("PH_DEBUG_SKIP_MMA")) : 0;
  p.debug_skip_mma = skip_mma;
  if (dedup && p.n_whole == batch) {
    // unique colours + multiplicities per image (only worth it when a CTA owns whole images)
    float4* ulist = static_cast<float4*>(workspace);
    int* nunique = reinterpret_cast<int*>(static_cast<char*>(workspace) +
                                          align_up((size_t)batch * DEDUP_MAX * sizeof(float4), 256));
    hist_dedup_kernel<<<(unsigned)batch, 256, 0, st>>>(image, npix, channels, ulist, nunique);
    PH_LAUNCH_OK("hist_dedup_kernel");
    p.ulist = ulist;
    p.nunique = nunique;
  }
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
  if (p.n_whole < batch) {
    launch_finalize(p.partial, p.splits, 3, bins, 1, hist + p.n_whole * (int64_t)(3 * bins * bins), denom + p.n_whole,
                    batch - p.n_whole, st);
    PH_LAUNCH_OK("hist_finalize_kernel");
  }
  return PH_OK;
}

}  // namespace ph
