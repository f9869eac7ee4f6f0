// RGB-uv histogram, tensor-core engine (tcgen05, sm_100a): the Ku^T.Kv contraction over pixels
// (histogram.py:29-30) as kind::f16 MMAs with fp32 emulation by operand splitting, fused with the
// per-image normalisation (histogram.py:75-79).
//
// Emulation: every weight is generated directly in a power-of-two scaled form that fits fp16's range and
// split w = hi + lo (both fp16, 11 significant bits each).  The A tile stacks the hi rows and the lo rows of
// the 64 u-bins along M (128 rows), the B tile stacks hi and lo of the 64 v-bins along N (128 columns), so
// ONE M128 x N128 x K16 instruction accumulates all four cross terms hi.hi, hi.lo, lo.hi, lo.lo into four
// 64 x 64 quadrants of the accumulator, which the epilogue adds up.  (tools/emul_f16split.py: the split
// error is 1e-7 on the histogram, the same as the tf32 split; K = 16 per instruction and 2-byte operands
// halve the tensor-pipe time and the shared-memory traffic of the tf32 version.)
//
// Forward, 64 bins, one persistent CTA per SM, 20 warps:
//   warps 17-19 pixel pass: 128-bit RGBA loads, log-chroma u/v per channel and intensity Iy -> smem ring
//   warps 0-7   A operand (u side, Iy-weighted), warps 8-15 B operand (v side): thread = (bin, 8 pixels);
//               both operands go to shared memory as K-major no-swizzle core matrices, every thread writes
//               whole 16-byte core-matrix rows (hi row and lo row), no exchange between threads.
//               The same 16 warps drain the accumulators at the end of a chain.
//   warp 16     one thread issues tcgen05.mma (SS, M=128, N=128, K=16): 2 per channel and 32-pixel stage
// The bin weights never exist in global memory: they are generated from 16 B per pixel with packed
// fp32x2 arithmetic (FADD2/FMUL2/FFMA2) and one MUFU.RCP per weight pair.
//
// Accuracy: the tensor core adds every K block into the fp32 accumulator with truncation, so the
// error grows with the chain length (measured 8e-6 for 4096 pixels in one chain, ~1e-6 for 1024).
// Chains are cut at 1024 pixels and summed in fp32 in a shared-memory accumulator; when a CTA owns a
// whole image the normaliser D and H/D are produced in the same kernel.
#include <stdlib.h>

#include "common.cuh"
#include "hist_internal.cuh"
#include "tc_ptx.cuh"
#include "hist_tc_gen.cuh"

namespace ph {

using namespace tc;

namespace fwdtc {

using tcgen::named_bar_sync;
using tcgen::weight2;
using tcgen::weight4;

constexpr int BINS = 64;
constexpr int KB = 32;         // pixels per pipeline stage
constexpr int NS = 3;          // operand stages
constexpr int CHAIN_KB = 32;   // stages per TMEM accumulation chain (1024 pixels)
constexpr int A_WARPS = 8, B_WARPS = 8, PXW = 3;  // 20 warps: 640 threads leave 96 registers per thread
constexpr int PROD_WARPS = A_WARPS + B_WARPS;
constexpr int MMA_WARP = PROD_WARPS;            // 16
constexpr int PX_WARP0 = MMA_WARP + 1;          // 17
constexpr int PR = 6;                           // pixel ring slots
constexpr int DEDUP_MAX = 512;                  // unique colours kept per image by the de-duplication pass
constexpr int DEDUP_SLOTS = 1024;
constexpr int THREADS = (PX_WARP0 + PXW) * 32;  // 640
constexpr int TMEM_COLS = 512;
constexpr int D_COLS = 128;                // per channel: [B_hi columns | B_lo columns]
// operand tile of one channel and stage: 128 rows (64 hi + 64 lo) x 32 pixels of fp16, K-major, no swizzle:
//   [kb = pixel / 8 (4)][row / 8 (16)][row % 8][pixel % 8]     core matrix = 8 rows x 8 halfs = 128 B
// (the MN-major variant — thread = (pixel, 8 bins), one 4-byte shared load per coordinate instead of
// broadcast 16-byte loads — measured 6 % slower: tools/f16_layout_test.cu keeps the layout check for it)
constexpr int T_KB_BYTES = 16 * 128;       // 2048: one core-matrix column (8 pixels) for all 128 rows
constexpr int TILE_BYTES = 4 * T_KB_BYTES; // 8192
constexpr int STAGE_BYTES = 6 * TILE_BYTES;  // A and B tiles of the three channels: [c][A, B]
static_assert(3 * D_COLS <= TMEM_COLS, "TMEM budget");
static_assert(KB == 32, "stage = 32 pixels (shifts below)");

struct PxSlot {
  float u[3][KB], v[3][KB], iy[KB];
};

struct Smem {
  alignas(128) unsigned char ab[NS][STAGE_BYTES];   // 144 KB
  float acc[3][BINS][BINS + 1];                     // [c][j][i] running fp32 sum over chains; rows padded to 65
                                                    // floats so both the bin-major drain and the j-major
                                                    // write-out are free of bank conflicts
  PxSlot px[PR];
  float dom[2][BINS];  // [u side, v side]
  float red[PROD_WARPS];
  double red2[PROD_WARPS];
  int last_flag;
  alignas(8) uint64_t px_full[PR], px_empty[PR], ab_full[NS], ab_empty[NS], d_full, d_empty;
  uint32_t tmem_base;
};

struct Params {
  const float* image;
  const float* dom_u;  // 64 bin centres of the u side (rows i of the histogram block)
  const float* dom_v;  // 64 bin centres of the v side (columns j)
  float* raw_out;      // optional (B, 3, 64, 64): emit the un-normalised block sums instead of H/D (histograms with
                       // more than 64 bins are assembled from 64 x 64 blocks, one launch per block)
  float* partial;  // (B - n_whole, splits, 3, 64, 64) raw sums of the sliced ("tail") images
  float* hist;     // (B, 64, 64, 3) normalised, written directly for whole-image items
  float* denom;    // (B)
  const float* hist_true;  // optional (B, 64, 64, 3): accumulate the Hellinger sum of squares against it into *ssum
  double* ssum;            // (histogram.py:88) while the normalised histogram is still on chip, or NULL
  int* tail_counter;       // (B - n_whole): slices of a tail image that have delivered their partial sums
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
  float eps;
  // Coordinates and bin centres are pre-multiplied by the power of two `coord_scale` (exact), d = s (x - c), so that
  // the inverse-quadratic weight needs one FMA and a reciprocal:  IQ: K / w = 1 / (d d + wb),  wb = w = s^2 sigma^2
  // in [2^-14, 2^-13];  RBF: 2^14 K = 2^(wa d d + wb)
  float wa, wb, coord_scale;
  float iy_scale;    // power of two applied to the intensity x multiplicity of a DE-DUPLICATED image so that the A
                     // operand stays below fp16's 65504: 2^-ceil(log2 npix).  Dense images (also those a
                     // de-duplicated batch falls back on) keep scale 1: with the batch-wide scale their far-bin
                     // weights dropped into fp16 subnormals
  float inv_scale;        // 1 / (weight scale^2 * iy_scale): raw sums of a de-duplicated image -> true scale
  float inv_scale_dense;  // 1 / weight scale^2: raw sums of a dense image -> true scale
  int* status;            // mapped host word (ph_async_status): bit PH_ASYNC_RANGE when an A operand would overflow fp16
  float mirror_tol;       // mirrored-tile kernel: largest |c_j + c_{63-j}| the caller's PH_IMPL_MIRROR may stand for
};

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

// ---- item epilogue of the forward kernels (all 16 producer warps; named barrier 5).  A whole image is normalised
//      here; a slice of a tail image delivers its raw sums, and the CTA that delivers the last slice adds them up (in
//      slice order: deterministic) and normalises — no second kernel, no partials left for the host side to combine.
//      `S.acc[c][j][i]` holds the raw sums of the item (scaled by 1 / inv_scale). ----
template <bool FUSE_SSUM, typename SmemT>
__device__ __forceinline__ void finish_item(SmemT& S, const Params& p, const ItemRange& ir, int tid, int warp, int lane) {
  named_bar_sync(5, PROD_WARPS * 32);
  const int t = tid;  // 0..511
  const int64_t b = ir.b;
  bool finish = ir.whole;
  const float item_inv_scale = ir.dedup ? p.inv_scale : p.inv_scale_dense;
  float dscale = item_inv_scale;  // raw sum -> true scale of the normaliser D
  if (!ir.whole) {
    float* dst = p.partial + ir.pidx * (int64_t)(3 * BINS * BINS);
    for (int e = t; e < 3 * BINS * BINS; e += PROD_WARPS * 32) {
      const int c = e >> 12, i = (e >> 6) & 63, j = e & 63;
      dst[e] = S.acc[c][j][i] * item_inv_scale;
    }
    __threadfence();
    named_bar_sync(5, PROD_WARPS * 32);
    if (t == 0) S.last_flag = atomicAdd(p.tail_counter + (b - p.n_whole), 1) == p.splits - 1;
    named_bar_sync(5, PROD_WARPS * 32);
    finish = S.last_flag != 0;
    if (finish) {
      __threadfence();
      const float* src = p.partial + (b - p.n_whole) * (int64_t)p.splits * (3 * BINS * BINS);
      for (int e = t; e < 3 * BINS * BINS; e += PROD_WARPS * 32) {
        float v = 0.f;
        for (int sidx = 0; sidx < p.splits; ++sidx) v += __ldcg(src + (int64_t)sidx * (3 * BINS * BINS) + e);
        S.acc[e >> 12][e & 63][(e >> 6) & 63] = v;
      }
      dscale = 1.0f;
      named_bar_sync(5, PROD_WARPS * 32);
    }
  }
  if (finish && p.raw_out != nullptr) {
    float* dst = p.raw_out + b * (int64_t)(3 * BINS * BINS);
    for (int e = t; e < 3 * BINS * BINS; e += PROD_WARPS * 32) {
      const int c = e >> 12, i = (e >> 6) & 63, j = e & 63;
      dst[e] = S.acc[c][j][i] * dscale;
    }
  } else if (finish) {
    float s = 0.f;
    {
      // element k of this thread in [c][i][j] order: e = t + 512 k -> c = k / 8, i = t / 64 + 8 (k % 8), j = t % 64
      const float* a0 = &S.acc[0][t >> 6][t & 63];
      constexpr int PLANE = BINS * (BINS + 1), ROW8 = 8 * (BINS + 1);
#pragma unroll
      for (int k = 0; k < 3 * BINS * BINS / (PROD_WARPS * 32); ++k) s += a0[(k >> 3) * PLANE + (k & 7) * ROW8];
    }
    s = warp_sum(s);
    if (lane == 0) S.red[warp] = s;
    named_bar_sync(5, PROD_WARPS * 32);
    float d = 0.f;
#pragma unroll
    for (int k = 0; k < PROD_WARPS; ++k) d += S.red[k];
    if (t == 0) p.denom[b] = d * dscale;
    const float inv_d = 1.0f / d;
    float* dst = p.hist + b * (int64_t)(3 * BINS * BINS);
    // Element k of this thread is e = t + 512 k of the channel-last histogram, e = (i 64 + j) 3 + c.  512 x 3 is a
    // multiple of 3 and of 3 x 64: elements k and k + 3 share channel and v-bin and lie 8 u-bins apart, so three base
    // pointers serve all 24 (the divisions by 3 were a quarter of a de-duplicated item's instructions)
    constexpr int PER_THREAD = 3 * BINS * BINS / (PROD_WARPS * 32);  // 24
    const float* a3[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int e = t + r * (PROD_WARPS * 32), c = e % 3, ij = e / 3;
      a3[r] = &S.acc[c][ij & 63][ij >> 6];
    }
    if (!FUSE_SSUM) {
#pragma unroll
      for (int k = 0; k < PER_THREAD; ++k) dst[t + k * (PROD_WARPS * 32)] = a3[k % 3][8 * (k / 3)] * inv_d;
    } else {
      // fused Hellinger partial: sum (sqrt(Hp) - sqrt(Ht))^2 of this image (histogram.py:88) while Hp is
      // on chip; the real image's histogram arrives in two batches of 12 independent loads per thread
      constexpr int HALF = PER_THREAD / 2;  // 12
      static_assert(HALF % 3 == 0, "the three base pointers repeat every three elements");
      const float* ht = p.hist_true + b * (int64_t)(3 * BINS * BINS);
      float part = 0.f;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float htv[HALF];
#pragma unroll
        for (int k = 0; k < HALF; ++k) htv[k] = __ldg(ht + t + (half * HALF + k) * (PROD_WARPS * 32));
#pragma unroll
        for (int k = 0; k < HALF; ++k) {
          const int e = t + (half * HALF + k) * (PROD_WARPS * 32);
          const float h = a3[k % 3][8 * (half * (HALF / 3) + k / 3)] * inv_d;  // HALF is a multiple of 3
          dst[e] = h;
          const float df = fast_sqrt(h) - fast_sqrt(htv[k]);
          part = fmaf(df, df, part);
        }
      }
      double dpart = warp_sum((double)part);
      if (lane == 0) S.red2[warp] = dpart;
      named_bar_sync(5, PROD_WARPS * 32);
      if (t == 0) {
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < PROD_WARPS; ++k) tot += S.red2[k];
        atomicAdd(p.ssum, tot);
      }
    }
  }
  named_bar_sync(5, PROD_WARPS * 32);
}

template <int METHOD, bool FUSE_SSUM>
__global__ void __launch_bounds__(THREADS, 1) hist_fwd_tc_kernel(Params p) {
  // no-swizzle operand tiles need only 16 B alignment; keeping the pointer derived from the
  // __shared__ symbol (no integer round trip) lets ptxas emit LDS/STS instead of generic LD/ST
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;

  if (tid == 0) {
    // producer/consumer barriers count WARPS (mbar_arrive_warp), the tcgen05.commit ones count 1
    for (int i = 0; i < PR; ++i) { mbar_init(&S.px_full[i], 1); mbar_init(&S.px_empty[i], PROD_WARPS); }
    for (int i = 0; i < NS; ++i) { mbar_init(&S.ab_full[i], PROD_WARPS); mbar_init(&S.ab_empty[i], 1); }
    mbar_init(&S.d_full, 1);
    mbar_init(&S.d_empty, PROD_WARPS);
    fence_mbar_init();
  }
  if (tid < BINS) { S.dom[0][tid] = p.dom_u[tid] * p.coord_scale; S.dom[1][tid] = p.dom_v[tid] * p.coord_scale; }
  if (warp == MMA_WARP) tmem_alloc(&S.tmem_base, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = S.tmem_base;

  const int64_t first = blockIdx.x, step = gridDim.x;

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
        const float d_rg = logf(e0 / e1) * p.coord_scale, d_rb = logf(e0 / e2) * p.coord_scale,
                    d_gb = logf(e1 / e2) * p.coord_scale;
        mbar_wait_relaxed(&S.px_empty[slot], ((it / PR) & 1) ^ 1, 400);
        PxSlot& o = S.px[slot];
        // (u,v): R:(rg, rb)  G:(-rg, gb)  B:(-rb, -gb)   (histogram.py:72-74)
        o.u[0][lane] = d_rg;  o.v[0][lane] = d_rb;
        o.u[1][lane] = -d_rg; o.v[1][lane] = d_gb;
        o.u[2][lane] = -d_rb; o.v[2][lane] = -d_gb;
        const float a_iy = valid ? iy * mult * (ir.dedup ? p.iy_scale : 1.0f) : 0.f;  // masked pixels contribute nothing (A operand = 0)
        // the A operand is a_iy x (scaled weight <= 2^14): an image far outside [-1, 1] would overflow fp16 where the
        // reference computes a finite value (histogram.py:58) — flagged, never silent
        if (a_iy > tcgen::IY_OPERAND_LIMIT) *reinterpret_cast<volatile int*>(p.status) = PH_ASYNC_RANGE;
        o.iy[lane] = a_iy;
        mbar_arrive_warp(&S.px_full[slot]);
      }
    }
  } else if (warp < PROD_WARPS) {
    // ===================== operand producers (A: warps 0-7, B: warps 8-15) + epilogue =====================
    const int side = warp >> 3;                 // 0: A (u side, x intensity), 1: B (v side)   — warp-uniform
    const int bin = tid & 63;
    const int po = (tid >> 6) & 3;              // which 8 of the stage's 32 pixels
    const float c_bin = S.dom[side][bin];
    const f32x2 negc = pack2(-c_bin, -c_bin);
    const f32x2 wa2 = pack2(p.wa, p.wa), wb2 = pack2(p.wb, p.wb), mone2 = pack2(-1.0f, -1.0f);
    // hi row `bin`, lo row `bin + 64` (+ 8 row groups = 1024 B) of core-matrix column `po`
    const uint32_t row_off = (uint32_t)(side * TILE_BYTES + po * T_KB_BYTES + (bin >> 3) * 128 + (bin & 7) * 16);
    // epilogue role: TMEM sub-partition quad = warp % 4 -> rows 32 quad .. (hi rows of bins 0-63 in quads 0,1,
    // lo rows in quads 2,3); column block cs = warp / 4 -> v-bins j in [16 cs, 16 cs + 16)
    const int quad = warp & 3, cs = warp >> 2;
    const int ebin = (quad & 1) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    uint32_t it = 0, chain = 0;
    for (int64_t w = first; w < p.items; w += step) {
      const ItemRange ir = item_range(p, w);
      const int64_t b = ir.b;
      const uint32_t nkb = (ir.px1 - ir.px0 + KB - 1) >> 5;
      for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
        const int slot = it % PR, stage = it % NS;
        mbar_wait(&S.px_full[slot], (it / PR) & 1);
        const PxSlot& in = S.px[slot];
        // all loads first (the compiler cannot hoist them across the shared-memory stores below)
        ulonglong2 xx[3][2], iw[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float* src = side == 0 ? &in.u[c][po * 8] : &in.v[c][po * 8];
          xx[c][0] = *reinterpret_cast<const ulonglong2*>(src);
          xx[c][1] = *reinterpret_cast<const ulonglong2*>(src + 4);
        }
        if (side == 0) {
          iw[0] = *reinterpret_cast<const ulonglong2*>(&in.iy[po * 8]);
          iw[1] = *reinterpret_cast<const ulonglong2*>(&in.iy[po * 8 + 4]);
        }
        mbar_arrive_warp(&S.px_empty[slot]);
        uint4 hi[3], lo[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          f32x2 w0, w1, w2, w3;
          weight4<METHOD>(xx[c][0].x, xx[c][0].y, negc, wa2, wb2, w0, w1);
          weight4<METHOD>(xx[c][1].x, xx[c][1].y, negc, wa2, wb2, w2, w3);
          if (side == 0) { w0 = mul2(w0, iw[0].x); w1 = mul2(w1, iw[0].y); w2 = mul2(w2, iw[1].x); w3 = mul2(w3, iw[1].y); }
          split_f16x2(w0, mone2, hi[c].x, lo[c].x);
          split_f16x2(w1, mone2, hi[c].y, lo[c].y);
          split_f16x2(w2, mone2, hi[c].z, lo[c].z);
          split_f16x2(w3, mone2, hi[c].w, lo[c].w);
        }
        mbar_wait(&S.ab_empty[stage], ((it / NS) & 1) ^ 1);  // the MMAs that read this stage are done
        unsigned char* tile = &S.ab[stage][row_off];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          *reinterpret_cast<uint4*>(tile + c * 2 * TILE_BYTES) = hi[c];
          *reinterpret_cast<uint4*>(tile + c * 2 * TILE_BYTES + 1024) = lo[c];
        }
        fence_proxy_async_smem();
        mbar_arrive_warp(&S.ab_full[stage]);

        const bool chain_end = ((kb + 1) % CHAIN_KB == 0) || (kb + 1 == nkb);
        if (!chain_end) continue;
        // ---- chain epilogue: the four quadrants of D (TMEM) += into the fp32 shared-memory accumulator ----
        mbar_wait(&S.d_full, chain & 1);
        ++chain;
        tc_fence_after_sync();
        // hi-row warps first (plain stores on the first chain of an item, else read-add-write), then the
        // lo-row warps add their share: no shared-memory float atomics (those are CAS loops)
        const bool first_chain = kb < CHAIN_KB;
        float val[3][16];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uint32_t v1[16], v2[16];
          tmem_ld16(tmem + lane_addr + c * D_COLS + cs * 16, v1);
          tmem_ld16(tmem + lane_addr + c * D_COLS + 64 + cs * 16, v2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) val[c][i] = __uint_as_float(v1[i]) + __uint_as_float(v2[i]);
        }
        tc_fence_before_sync();
        mbar_arrive_warp(&S.d_empty);  // the accumulators are in registers: the next chain may start
        if (quad < 2) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float* a = &S.acc[c][cs * 16 + i][ebin];
              *a = first_chain ? val[c][i] : *a + val[c][i];
            }
        }
        named_bar_sync(5, PROD_WARPS * 32);
        if (quad >= 2) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int i = 0; i < 16; ++i) S.acc[c][cs * 16 + i][ebin] += val[c][i];
        }
        if (kb + 1 != nkb) { named_bar_sync(5, PROD_WARPS * 32); continue; }  // acc is read-modified by the next chain

        finish_item<FUSE_SSUM>(S, p, ir, tid, warp, lane);
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issue: the whole warp runs the (uniform) loop, one elected lane issues ====
    constexpr uint32_t IDESC = idesc_f16(128, 128);
    const uint64_t desc0 = smem_desc_kmajor_noswizzle(smem_u32(&S.ab[0][0]), T_KB_BYTES, 128);
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
          const uint32_t dstage = dlo0 + stage * (STAGE_BYTES >> 4);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int ks = 0; ks < KB / 16; ++ks) {
              const uint32_t a_d = dstage + ((c * 2 * TILE_BYTES + ks * 2 * T_KB_BYTES) >> 4);
              const uint32_t b_d = a_d + (TILE_BYTES >> 4);
              const uint32_t acc0 = (k == 0 && ks == 0) ? 0u : 1u;
              if (elect_one_sync()) mma_f16_ss2(tm + c * D_COLS, a_d, b_d, dhi, IDESC, acc0);
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


// =============================================================================================
// Variant with the A operand in TENSOR MEMORY (default; PH_FWD_A=smem selects the kernel above).
//
// ncu on the kernel above: the shared-memory pipe is the scarcest resource (operand tiles written once by the
// producers — 443 store wavefronts per 32-pixel stage — and read once by the MMAs, ~85 % busy); removing 14 % of the
// weight-generation instructions did not change its run time, leaving out the A tiles' stores (timing experiment)
// took 8 % off.  Here the u-side operand never touches shared memory:
//   * M = 64: a row per u-bin.  Row m of an M = 64 tile lives in TMEM lane 32 (m / 16) + m % 16 (+ 16 for a second,
//     interleaved tile of the same columns; an instruction's A and D must use the same lane offset —
//     tools/m64_layout_test.cu).  A producer thread = (bin, 16 pixels): lanes 0-15 of a warp hold the bins
//     16 q .. 16 q + 15 for pixels 0-15 of the stage, lanes 16-31 the same bins for pixels 16-31, and every thread
//     writes [hi (8 columns) | lo (8 columns)] of its own row with one tcgen05.st per channel: no exchange between
//     threads, hi and lo stacked along K instead of M.
//   * per channel, 16-pixel half h and stage:  D_h += A_hi[h] . [B_hi | B_lo]^T  (M64 N128 K16, 64 cycles) and
//     D_h[:, 0:64] += A_lo[h] . B_hi^T  (M64 N64 K16, 44 cycles): the three products of the emulation (the M = 128
//     kernel's single instruction also computes lo.lo); D_0 and D_1 (lane offsets 0 / 16) are partial sums over the two
//     pixel halves and are added in the epilogue.  Tensor pipe 648 cycles per stage instead of 384.
//   * the eight A warps alternate stages (set 0: even, set 1: odd), two 48-column A slots in TMEM; the v side is
//     produced into shared memory as before (24 KB per stage instead of 48).
// Shared-memory traffic per stage: 192 + ~60 store, ~290 load wavefronts, 288 cycles of MMA operand reads
// (B_hi | B_lo once, B_hi once more, per half and channel) instead of 443 / ~290 / 384.
// =============================================================================================
constexpr int NSB = 3;                      // B operand stages in shared memory
constexpr int NSA = 2;                      // A operand slots in tensor memory (one per A warp set)
constexpr int B_STAGE_BYTES = 3 * TILE_BYTES;  // B tiles of the three channels
constexpr int A_COL0 = 3 * D_COLS;          // 384: first A column
constexpr int A_SLOT_COLS = 3 * 16;         // per slot: three channels x [hi 8 | lo 8] columns
static_assert(A_COL0 + NSA * A_SLOT_COLS <= TMEM_COLS, "TMEM budget");

struct SmemA {
  alignas(128) unsigned char bt[NSB][B_STAGE_BYTES];  // 72 KB
  float acc[3][BINS][BINS + 1];
  PxSlot px[PR];
  float dom[2][BINS];
  float red[PROD_WARPS];
  double red2[PROD_WARPS];
  int last_flag;
  alignas(8) uint64_t px_full[PR], px_empty[PR], b_full[NSB], b_empty[NSB], a_full[NSA], a_empty[NSA], d_full, d_empty;
  uint32_t tmem_base;
};

template <int METHOD, bool FUSE_SSUM>
__global__ void __launch_bounds__(THREADS, 1) hist_fwd_tca_kernel(Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemA& S = *reinterpret_cast<SmemA*>(smem_raw);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;

  if (tid == 0) {
    // a pixel slot is read by the four A warps of one set and the eight B warps
    for (int i = 0; i < PR; ++i) { mbar_init(&S.px_full[i], 1); mbar_init(&S.px_empty[i], A_WARPS / 2 + B_WARPS); }
    for (int i = 0; i < NSB; ++i) { mbar_init(&S.b_full[i], B_WARPS); mbar_init(&S.b_empty[i], 1); }
    for (int i = 0; i < NSA; ++i) { mbar_init(&S.a_full[i], A_WARPS / 2); mbar_init(&S.a_empty[i], 1); }
    mbar_init(&S.d_full, 1);
    mbar_init(&S.d_empty, PROD_WARPS);
    fence_mbar_init();
  }
  if (tid < BINS) { S.dom[0][tid] = p.dom_u[tid] * p.coord_scale; S.dom[1][tid] = p.dom_v[tid] * p.coord_scale; }
  if (warp == MMA_WARP) tmem_alloc(&S.tmem_base, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = S.tmem_base;
  pdl_wait();               // set-up done; from here on the results of the stream's earlier kernels are read
  pdl_launch_dependents();  // the next kernel of the stream may move in as this one's CTAs retire

  const int64_t first = blockIdx.x, step = gridDim.x;

  if (warp >= PX_WARP0) {
    // ===================== pixel pass (as in the kernel above) =====================
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
        const float x0 = fmaf(r, 0.5f, 0.5f), x1 = fmaf(g, 0.5f, 0.5f), x2 = fmaf(bl, 0.5f, 0.5f);
        const float iy = sqrtf(x0 * x0 + x1 * x1 + x2 * x2 + p.eps);
        const float e0 = x0 + p.eps, e1 = x1 + p.eps, e2 = x2 + p.eps;
        const float d_rg = logf(e0 / e1) * p.coord_scale, d_rb = logf(e0 / e2) * p.coord_scale,
                    d_gb = logf(e1 / e2) * p.coord_scale;
        mbar_wait_relaxed(&S.px_empty[slot], ((it / PR) & 1) ^ 1, 400);
        PxSlot& o = S.px[slot];
        o.u[0][lane] = d_rg;  o.v[0][lane] = d_rb;
        o.u[1][lane] = -d_rg; o.v[1][lane] = d_gb;
        o.u[2][lane] = -d_rb; o.v[2][lane] = -d_gb;
        const float a_iy = valid ? iy * mult * (ir.dedup ? p.iy_scale : 1.0f) : 0.f;
        if (a_iy > tcgen::IY_OPERAND_LIMIT) *reinterpret_cast<volatile int*>(p.status) = PH_ASYNC_RANGE;
        o.iy[lane] = a_iy;
        mbar_arrive_warp(&S.px_full[slot]);
      }
    }
  } else if (warp < PROD_WARPS) {
    // ===================== operand producers (A -> TMEM: warps 0-7, B -> smem: warps 8-15) + epilogue ==========
    const int side = warp >> 3;                 // warp-uniform
    const f32x2 wa2 = pack2(p.wa, p.wa), wb2 = pack2(p.wb, p.wb), mone2 = pack2(-1.0f, -1.0f);
    // A side: quadrant q = warp % 4 (the TMEM lanes this warp may touch), set = stages it handles; thread = (bin, 16 px)
    const int quad = warp & 3, aset = (warp >> 2) & 1;
    const int a_bin = quad * 16 + (lane & 15), a_half = lane >> 4;
    // B side: thread = (bin, 8 px) as in the kernel above
    const int b_bin = tid & 63, po = (tid >> 6) & 3;
    const float c_bin = side == 0 ? S.dom[0][a_bin] : S.dom[1][b_bin];
    const f32x2 negc = pack2(-c_bin, -c_bin);
    const uint32_t row_off = (uint32_t)(po * T_KB_BYTES + (b_bin >> 3) * 128 + (b_bin & 7) * 16);
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    // epilogue role: lanes 32 quad .. (u-bins 16 quad .. + 15, tile = lane / 16), column block cs = warp / 4
    const int cs = warp >> 2;
    uint32_t it = 0, chain = 0;
    for (int64_t w = first; w < p.items; w += step) {
      const ItemRange ir = item_range(p, w);
      const uint32_t nkb = (ir.px1 - ir.px0 + KB - 1) >> 5;
      for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
        const int slot = it % PR;
        if (side == 1) {
          // ---------------- B operand of this stage -> shared memory ----------------
          const int stage = it % NSB;
          mbar_wait(&S.px_full[slot], (it / PR) & 1);
          const PxSlot& in = S.px[slot];
          ulonglong2 xx[3][2];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            xx[c][0] = *reinterpret_cast<const ulonglong2*>(&in.v[c][po * 8]);
            xx[c][1] = *reinterpret_cast<const ulonglong2*>(&in.v[c][po * 8 + 4]);
          }
          mbar_arrive_warp(&S.px_empty[slot]);
          uint4 hi[3], lo[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            f32x2 w0, w1, w2, w3;
            weight4<METHOD>(xx[c][0].x, xx[c][0].y, negc, wa2, wb2, w0, w1);
            weight4<METHOD>(xx[c][1].x, xx[c][1].y, negc, wa2, wb2, w2, w3);
            split_f16x2(w0, mone2, hi[c].x, lo[c].x);
            split_f16x2(w1, mone2, hi[c].y, lo[c].y);
            split_f16x2(w2, mone2, hi[c].z, lo[c].z);
            split_f16x2(w3, mone2, hi[c].w, lo[c].w);
          }
          mbar_wait(&S.b_empty[stage], ((it / NSB) & 1) ^ 1);  // the MMAs that read this stage are done
          unsigned char* tile = &S.bt[stage][row_off];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            *reinterpret_cast<uint4*>(tile + c * TILE_BYTES) = hi[c];
            *reinterpret_cast<uint4*>(tile + c * TILE_BYTES + 1024) = lo[c];
          }
          fence_proxy_async_smem();
          mbar_arrive_warp(&S.b_full[stage]);
        } else if ((int)(it & 1) == aset) {
          // ---------------- A operand of this stage -> tensor memory (this warp set's slot) ----------------
          mbar_wait(&S.px_full[slot], (it / PR) & 1);
          const PxSlot& in = S.px[slot];
          ulonglong2 iw[4];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) iw[q4] = *reinterpret_cast<const ulonglong2*>(&in.iy[a_half * 16 + q4 * 4]);
          // the MMAs that read this slot two stages ago are done
          mbar_wait(&S.a_empty[aset], ((it >> 1) & 1) ^ 1);
          tc_fence_after_sync();
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            ulonglong2 xx[4];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) xx[q4] = *reinterpret_cast<const ulonglong2*>(&in.u[c][a_half * 16 + q4 * 4]);
            uint32_t r[16];  // [hi: pixel pairs 0..7 | lo: pixel pairs 0..7] of this thread's row
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              f32x2 w0, w1;
              weight4<METHOD>(xx[q4].x, xx[q4].y, negc, wa2, wb2, w0, w1);
              w0 = mul2(w0, iw[q4].x);
              w1 = mul2(w1, iw[q4].y);
              split_f16x2(w0, mone2, r[2 * q4], r[8 + 2 * q4]);
              split_f16x2(w1, mone2, r[2 * q4 + 1], r[8 + 2 * q4 + 1]);
            }
            tmem_st16(tmem + lane_addr + A_COL0 + aset * A_SLOT_COLS + c * 16, r);
          }
          mbar_arrive_warp(&S.px_empty[slot]);
          tmem_st_wait();
          tc_fence_before_sync();
          mbar_arrive_warp(&S.a_full[aset]);
        }

        const bool chain_end = ((kb + 1) % CHAIN_KB == 0) || (kb + 1 == nkb);
        if (!chain_end) continue;
        // ---- chain epilogue: D (TMEM: two pixel-half tiles x [.B_hi | .B_lo] columns) += into the fp32 accumulator ----
        mbar_wait(&S.d_full, chain & 1);
        ++chain;
        tc_fence_after_sync();
        const bool first_chain = kb < CHAIN_KB;
        float val[3][16];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uint32_t v1[16], v2[16];
          tmem_ld16(tmem + lane_addr + c * D_COLS + cs * 16, v1);
          tmem_ld16(tmem + lane_addr + c * D_COLS + 64 + cs * 16, v2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) val[c][i] = __uint_as_float(v1[i]) + __uint_as_float(v2[i]);
        }
        tc_fence_before_sync();
        mbar_arrive_warp(&S.d_empty);  // the accumulators are in registers: the next chain may start
        // lanes l and l + 16 hold the same u-bin for the two pixel halves
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int i = 0; i < 16; ++i) val[c][i] += __shfl_xor_sync(0xffffffffu, val[c][i], 16);
        if (lane < 16) {
          const int ebin = quad * 16 + lane;
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float* a = &S.acc[c][cs * 16 + i][ebin];
              *a = first_chain ? val[c][i] : *a + val[c][i];
            }
        }
        if (kb + 1 != nkb) continue;  // every (c, j, i) is owned by one thread: no barrier between chains
        finish_item<FUSE_SSUM>(S, p, ir, tid, warp, lane);
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issue: uniform loop, one elected lane issues =====================
    constexpr uint32_t IDESC_FULL = idesc_f16(64, 128);  // . [B_hi | B_lo]
    constexpr uint32_t IDESC_HI = idesc_f16(64, 64);     // . B_hi only (the first 64 rows of the tile)
    const uint64_t desc0 = smem_desc_kmajor_noswizzle(smem_u32(&S.bt[0][0]), T_KB_BYTES, 128);
    const uint32_t dlo0 = (uint32_t)desc0, dhi = (uint32_t)(desc0 >> 32);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);  // provably uniform copy
    uint32_t stage = 0, bphase = 0, aslot = 0, aphase = 0, chain_par = 0;
    for (int64_t w = first; w < p.items; w += step) {
      const ItemRange ir = item_range(p, w);
      const uint32_t nkb = (ir.px1 - ir.px0 + KB - 1) >> 5;
      for (uint32_t kb0 = 0; kb0 < nkb; kb0 += CHAIN_KB) {
        const uint32_t n_this = min((uint32_t)CHAIN_KB, nkb - kb0);
        for (uint32_t k = 0; k < n_this; ++k) {
          mbar_wait(&S.b_full[stage], bphase);
          mbar_wait(&S.a_full[aslot], aphase);
          tc_fence_after_sync();
          const uint32_t dstage = dlo0 + stage * (B_STAGE_BYTES >> 4);
          const uint32_t acol = tm + A_COL0 + aslot * A_SLOT_COLS;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // pixel half h: K step h of the B tile, TMEM lane offset 16 h for A and D
              const uint32_t b_d = dstage + ((c * TILE_BYTES + h * 2 * T_KB_BYTES) >> 4);
              const uint32_t lanes = (uint32_t)(16 * h) << 16;
              const uint32_t acc0 = k == 0 ? 0u : 1u;
              if (elect_one_sync()) {
                mma_f16_ts2(tm + lanes + c * D_COLS, acol + lanes + c * 16, b_d, dhi, IDESC_FULL, acc0);
                mma_f16_ts2(tm + lanes + c * D_COLS, acol + lanes + c * 16 + 8, b_d, dhi, IDESC_HI, 1u);
              }
            }
          }
          if (elect_one_sync()) { mma_commit(&S.b_empty[stage]); mma_commit(&S.a_empty[aslot]); }
          if (++stage == NSB) { stage = 0; bphase ^= 1; }
          aslot ^= 1;
          if (aslot == 0) aphase ^= 1;
        }
        if (elect_one_sync()) mma_commit(&S.d_full);
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


// =============================================================================================
// Mirrored-tile forward for DENSE images (default for 64 bins when the bin centres are antisymmetric to within the
// rounding of `linspace`; PH_FWD_SYM=0 or impl without PH_IMPL_MIRROR selects the kernel above).
//
// The six coordinates of a pixel are +-a, +-b, +-c with a = log R/G, b = log R/B, c = log G/B (histogram.py:72-74):
//     R: (u, v) = (a, b)      G: (-a, c)      B: (-b, -c)
// and K(-x - c_i) = K(x + c_i) = K(x - c_{63-i}) when c_{63-i} = -c_i: the weight vector of -x is the bin-reversed
// vector of x.  TensorFlow's `linspace(-3, 3, 64)` is antisymmetric only to 3.6e-7 (start + i * delta in float32), so
// the vectors are generated once around the MIDPOINT centres t_j = (c_j - c_{63-j}) / 2 and used for both signs: every
// use is off by at most half the asymmetry, 1.8e-7 — the size of the float32 rounding of the coordinate itself.
// Measured cost (tools/tc_check_sym.py, tests): histogram 1.9-2.1e-6 from the float64 reference at 64 x 64 pixels (kernel
// above: 0.9-1.2e-6), 5.3e-6 for an 8 x 8 image (2.7e-6); gradients move by 0.4-1.9e-6 through G^.  The backward keeps the
// exact centres (dK/du is ten times as sensitive).
// With alpha = sqrt(Iy) k(a), beta = sqrt(Iy) k(b), gamma = sqrt(Iy) k(c)  (64 bins each, centres t):
//     H_R[i, j] = (alpha beta^T)[i, j]     H_G[i, j] = (alpha gamma^T)[63-i, j]     H_B[i, j] = (beta gamma^T)[63-i, 63-j]
// i.e. THREE weight vectors per pixel instead of six (192 weights instead of 384), and one accumulator
//     D (128 x 256) = [alpha; beta] . [gamma_hi | beta_hi | gamma_lo | beta_lo]^T
// A = [alpha; beta] lives in tensor memory (M = 128: lane = row, hi and lo stacked along K as in the kernel above),
// B in shared memory (256 rows).  Per 16-pixel K step: A_hi . B (M128 N256 K16, 128 cycles, the pipe's full rate) and
// A_lo . [gamma_hi | beta_hi] (M128 N128 K16, 64 cycles) — the three products of the emulation; the beta.beta
// quadrant is not used.  192 tensor-pipe cycles per 16 pixels instead of 324, 6 144 generated weights per 32 pixels
// instead of 12 288.
// 64-pixel stages; every one of the 16 producer warps does the same work per stage: one TMEM task (32 rows of
// alpha or beta x 16 pixels: one tcgen05.st.x16 per thread; the beta warps also store their rows to the B tile) and
// one gamma task (32 bins x 8 pixels -> B tile).
// =============================================================================================
// ---- float32 difference of two logarithms for the pixel pass below -------------------------------------------------
// The three pixel-pass warps are this kernel's critical path (3 warps x one 128-pixel stage each: a slower pixel pass
// slows the kernel one for one, profiles/r2d_fwd_timing_experiments.txt), and logf(e0 / e1) costs an IEEE division + a
// libdevice log per coordinate.  Instead: log x = k ln2 + L_i + log1p(r) with x = 2^k m, the mantissa's top 7 bits
// picking one of 128 intervals with centre c_i, r = fma(m, RN(1 / c_i), -1), |r| < 2^-8 (three series terms).
// L_i = -log(RN(1 / c_i)) is tabulated as L_hi (a multiple of 2^-23: differences of two L_hi are exact) + L_lo, ln 2 is
// split as LN2_HI (a multiple of 2^-18: k LN2_HI is exact) + LN2_LO, so the large part of a DIFFERENCE of two logs,
// dk LN2_HI + dL_hi, costs one FMA whose rounding error is recovered exactly, and everything else is < 0.01.  The result
// is the correctly rounded float32 difference in all but ~1e-9 of absolute error (mean 0.25 ulp against 1.3 ulp for
// logf of the rounded ratio; the generator script make_log_table.py beside the test oracles replays it on the CPU).
__device__ const float4 PH_LOGF_TABLE[128] = {
    {0x1.fe01fe0000000p-1f, 0x1.ff00000000000p-9f, 0x1.5856220000000p-26f, 0.0f},     {0x1.fa11ca0000000p-1f, 0x1.7dc5000000000p-7f, -0x1.861fbe0000000p-25f, 0.0f},
    {0x1.f6310a0000000p-1f, 0x1.3cea800000000p-6f, -0x1.105cae0000000p-25f, 0.0f},     {0x1.f25f640000000p-1f, 0x1.b9fc000000000p-6f, 0x1.5f5f240000000p-27f, 0.0f},
    {0x1.ee9c800000000p-1f, 0x1.1b0d800000000p-5f, 0x1.0923da0000000p-25f, 0.0f},     {0x1.eae8080000000p-1f, 0x1.58a5c00000000p-5f, -0x1.506e360000000p-26f, 0.0f},
    {0x1.e741aa0000000p-1f, 0x1.95c8400000000p-5f, -0x1.266e380000000p-26f, 0.0f},     {0x1.e3a9180000000p-1f, 0x1.d276c00000000p-5f, -0x1.ba49ea0000000p-26f, 0.0f},
    {0x1.e01e020000000p-1f, 0x1.0759800000000p-4f, 0x1.24c7240000000p-27f, 0.0f},     {0x1.dca01e0000000p-1f, 0x1.253f600000000p-4f, 0x1.20a1420000000p-28f, 0.0f},
    {0x1.d92f220000000p-1f, 0x1.42edc00000000p-4f, 0x1.b34c8e0000000p-25f, 0.0f},     {0x1.d5cac80000000p-1f, 0x1.6065800000000p-4f, 0x1.5a6ea20000000p-25f, 0.0f},
    {0x1.d272ca0000000p-1f, 0x1.7da7600000000p-4f, 0x1.20f6260000000p-25f, 0.0f},     {0x1.cf26e60000000p-1f, 0x1.9ab4200000000p-4f, 0x1.29019e0000000p-27f, 0.0f},
    {0x1.cbe6da0000000p-1f, 0x1.b78c800000000p-4f, -0x1.6a78920000000p-27f, 0.0f},     {0x1.c8b2660000000p-1f, 0x1.d431400000000p-4f, -0x1.5a4d320000000p-26f, 0.0f},
    {0x1.c5894e0000000p-1f, 0x1.f0a3000000000p-4f, 0x1.c88b160000000p-27f, 0.0f},     {0x1.c26b540000000p-1f, 0x1.0671500000000p-3f, -0x1.86b4d20000000p-28f, 0.0f},
    {0x1.bf583e0000000p-1f, 0x1.1478600000000p-3f, -0x1.c8c5ea0000000p-26f, 0.0f},     {0x1.bc4fd60000000p-1f, 0x1.2266f00000000p-3f, 0x1.9452d60000000p-26f, 0.0f},
    {0x1.b951e20000000p-1f, 0x1.303d700000000p-3f, 0x1.3192000000000p-25f, 0.0f},     {0x1.b65e2e0000000p-1f, 0x1.3dfc300000000p-3f, -0x1.ec99ce0000000p-26f, 0.0f},
    {0x1.b374840000000p-1f, 0x1.4ba3700000000p-3f, 0x1.34d2b00000000p-26f, 0.0f},     {0x1.b094b40000000p-1f, 0x1.5933900000000p-3f, -0x1.a59f7e0000000p-25f, 0.0f},
    {0x1.adbe880000000p-1f, 0x1.66acd00000000p-3f, 0x1.01cab60000000p-25f, 0.0f},     {0x1.aaf1d20000000p-1f, 0x1.740f900000000p-3f, 0x1.fe01be0000000p-26f, 0.0f},
    {0x1.a82e660000000p-1f, 0x1.815c000000000p-3f, 0x1.670d600000000p-25f, 0.0f},     {0x1.a574100000000p-1f, 0x1.8e92900000000p-3f, 0x1.4436a20000000p-30f, 0.0f},
    {0x1.a2c2a80000000p-1f, 0x1.9bb3600000000p-3f, 0x1.51f7ee0000000p-25f, 0.0f},     {0x1.a01a020000000p-1f, 0x1.a8bed00000000p-3f, -0x1.07be880000000p-26f, 0.0f},
    {0x1.9d79f20000000p-1f, 0x1.b5b5100000000p-3f, 0x1.d03ed60000000p-25f, 0.0f},     {0x1.9ae24e0000000p-1f, 0x1.c296900000000p-3f, -0x1.dbcf9c0000000p-25f, 0.0f},
    {0x1.9852f00000000p-1f, 0x1.cf63600000000p-3f, -0x1.b7d8e80000000p-25f, 0.0f},     {0x1.95cbb00000000p-1f, 0x1.dc1bd00000000p-3f, -0x1.1aa09c0000000p-26f, 0.0f},
    {0x1.934c680000000p-1f, 0x1.e8c0200000000p-3f, 0x1.42a96a0000000p-25f, 0.0f},     {0x1.90d4f20000000p-1f, 0x1.f550a00000000p-3f, 0x1.d96f6a0000000p-28f, 0.0f},
    {0x1.8e65280000000p-1f, 0x1.00e6c00000000p-2f, 0x1.c56a800000000p-25f, 0.0f},     {0x1.8bfce80000000p-1f, 0x1.071b880000000p-2f, -0x1.f32a700000000p-26f, 0.0f},
    {0x1.899c100000000p-1f, 0x1.0d46b00000000p-2f, 0x1.ecd5ba0000000p-25f, 0.0f},     {0x1.87427c0000000p-1f, 0x1.1368700000000p-2f, -0x1.7b15d40000000p-28f, 0.0f},
    {0x1.84f00c0000000p-1f, 0x1.1980d00000000p-2f, 0x1.a2a11c0000000p-25f, 0.0f},     {0x1.82a4a00000000p-1f, 0x1.1f8ff80000000p-2f, 0x1.1245180000000p-25f, 0.0f},
    {0x1.8060180000000p-1f, 0x1.2596000000000p-2f, 0x1.1df7640000000p-26f, 0.0f},     {0x1.7e22560000000p-1f, 0x1.2b93000000000p-2f, 0x1.3789d40000000p-26f, 0.0f},
    {0x1.7beb3a0000000p-1f, 0x1.3187180000000p-2f, 0x1.20a20c0000000p-25f, 0.0f},     {0x1.79baa60000000p-1f, 0x1.3772680000000p-2f, 0x1.3fec320000000p-29f, 0.0f},
    {0x1.7790820000000p-1f, 0x1.3d54f80000000p-2f, -0x1.7e08e40000000p-30f, 0.0f},     {0x1.756cac0000000p-1f, 0x1.432ef00000000p-2f, 0x1.7c27400000000p-25f, 0.0f},
    {0x1.734f0c0000000p-1f, 0x1.4900680000000p-2f, 0x1.d8013a0000000p-27f, 0.0f},     {0x1.7137860000000p-1f, 0x1.4ec9780000000p-2f, -0x1.3effec0000000p-25f, 0.0f},
    {0x1.6f26020000000p-1f, 0x1.548a280000000p-2f, 0x1.536e940000000p-25f, 0.0f},     {0x1.6d1a620000000p-1f, 0x1.5a42b00000000p-2f, -0x1.e659800000000p-25f, 0.0f},
    {0x1.6b14900000000p-1f, 0x1.5ff3080000000p-2f, 0x1.d4f27c0000000p-27f, 0.0f},     {0x1.6914740000000p-1f, 0x1.659b580000000p-2f, -0x1.c7c1e00000000p-26f, 0.0f},
    {0x1.6719f40000000p-1f, 0x1.6b3bb00000000p-2f, 0x1.6d65100000000p-28f, 0.0f},     {0x1.6524f80000000p-1f, 0x1.70d4300000000p-2f, -0x1.d0edba0000000p-27f, 0.0f},
    {0x1.63356c0000000p-1f, 0x1.7664e00000000p-2f, -0x1.a312160000000p-29f, 0.0f},     {0x1.614b360000000p-1f, 0x1.7bede00000000p-2f, 0x1.0fbd7e0000000p-25f, 0.0f},
    {0x1.5f66440000000p-1f, 0x1.816f400000000p-2f, -0x1.37cad80000000p-28f, 0.0f},     {0x1.5d867c0000000p-1f, 0x1.86e9180000000p-2f, 0x1.2d985e0000000p-25f, 0.0f},
    {0x1.5babcc0000000p-1f, 0x1.8c5b800000000p-2f, -0x1.293a5c0000000p-25f, 0.0f},     {0x1.59d6200000000p-1f, 0x1.91c6780000000p-2f, 0x1.fa2d420000000p-25f, 0.0f},
    {0x1.5805600000000p-1f, 0x1.972a380000000p-2f, -0x1.d765760000000p-25f, 0.0f},     {0x1.56397c0000000p-1f, 0x1.9c86b00000000p-2f, -0x1.b47ef40000000p-27f, 0.0f},
    {0x1.54725e0000000p-1f, 0x1.a1dc080000000p-2f, -0x1.ba919a0000000p-28f, 0.0f},     {0x1.52aff60000000p-1f, 0x1.a72a480000000p-2f, -0x1.7509840000000p-28f, 0.0f},
    {0x1.50f22e0000000p-1f, 0x1.ac71900000000p-2f, -0x1.d33a780000000p-25f, 0.0f},     {0x1.4f38f60000000p-1f, 0x1.b1b1e00000000p-2f, 0x1.77dfc60000000p-26f, 0.0f},
    {0x1.4d843c0000000p-1f, 0x1.b6eb580000000p-2f, 0x1.9bcf360000000p-26f, 0.0f},     {0x1.4bd3ee0000000p-1f, 0x1.bc1e080000000p-2f, 0x1.e6d6860000000p-29f, 0.0f},
    {0x1.4a27fa0000000p-1f, 0x1.c14a000000000p-2f, 0x1.ad5f040000000p-26f, 0.0f},     {0x1.4880520000000p-1f, 0x1.c66f500000000p-2f, -0x1.5c09000000000p-26f, 0.0f},
    {0x1.46dce40000000p-1f, 0x1.cb8e080000000p-2f, -0x1.81942a0000000p-25f, 0.0f},     {0x1.453d9e0000000p-1f, 0x1.d0a6380000000p-2f, 0x1.b990f40000000p-25f, 0.0f},
    {0x1.43a2740000000p-1f, 0x1.d5b7f80000000p-2f, -0x1.59d3960000000p-26f, 0.0f},     {0x1.420b520000000p-1f, 0x1.dac3580000000p-2f, -0x1.6c9d360000000p-25f, 0.0f},
    {0x1.40782e0000000p-1f, 0x1.dfc8580000000p-2f, -0x1.6b92a40000000p-26f, 0.0f},     {0x1.3ee8f40000000p-1f, 0x1.e4c7180000000p-2f, 0x1.8743b80000000p-25f, 0.0f},
    {0x1.3d5d9a0000000p-1f, 0x1.e9bfa00000000p-2f, 0x1.bac3100000000p-25f, 0.0f},     {0x1.3bd60e0000000p-1f, 0x1.eeb2080000000p-2f, 0x1.8006f00000000p-25f, 0.0f},
    {0x1.3a52440000000p-1f, 0x1.f39e580000000p-2f, 0x1.2008f40000000p-25f, 0.0f},     {0x1.38d22e0000000p-1f, 0x1.f884a00000000p-2f, 0x1.b7d3da0000000p-27f, 0.0f},
    {0x1.3755be0000000p-1f, 0x1.fd64f00000000p-2f, -0x1.b93d500000000p-27f, 0.0f},     {0x1.35dce60000000p-1f, 0x1.011fac0000000p-1f, -0x1.ef400e0000000p-26f, 0.0f},
    {0x1.34679a0000000p-1f, 0x1.0389f00000000p-1f, 0x1.4b98d00000000p-27f, 0.0f},     {0x1.32f5ce0000000p-1f, 0x1.05f14c0000000p-1f, 0x1.38645a0000000p-25f, 0.0f},
    {0x1.3187760000000p-1f, 0x1.0855c80000000p-1f, -0x1.ca5d780000000p-28f, 0.0f},     {0x1.301c820000000p-1f, 0x1.0ab76c0000000p-1f, 0x1.0ee14e0000000p-25f, 0.0f},
    {0x1.2eb4ea0000000p-1f, 0x1.0d163c0000000p-1f, 0x1.019d6c0000000p-25f, 0.0f},     {0x1.2d50a00000000p-1f, 0x1.0f72400000000p-1f, 0x1.e9b4980000000p-25f, 0.0f},
    {0x1.2bef980000000p-1f, 0x1.11cb840000000p-1f, -0x1.ff06600000000p-26f, 0.0f},     {0x1.2a91ca0000000p-1f, 0x1.1422000000000p-1f, 0x1.d887aa0000000p-26f, 0.0f},
    {0x1.2937260000000p-1f, 0x1.1675cc0000000p-1f, -0x1.bb45a00000000p-25f, 0.0f},     {0x1.27dfa40000000p-1f, 0x1.18c6e00000000p-1f, 0x1.9ae7840000000p-28f, 0.0f},
    {0x1.268b380000000p-1f, 0x1.1b154c0000000p-1f, -0x1.0025d60000000p-25f, 0.0f},     {0x1.2539d80000000p-1f, 0x1.1d61100000000p-1f, -0x1.0624000000000p-27f, 0.0f},
    {0x1.23eb7a0000000p-1f, 0x1.1faa340000000p-1f, -0x1.063dac0000000p-27f, 0.0f},     {0x1.22a0120000000p-1f, 0x1.21f0c00000000p-1f, 0x1.05beec0000000p-29f, 0.0f},
    {0x1.2157980000000p-1f, 0x1.2434b80000000p-1f, -0x1.037c6c0000000p-25f, 0.0f},     {0x1.2012020000000p-1f, 0x1.2676200000000p-1f, -0x1.7abcf20000000p-25f, 0.0f},
    {0x1.1ecf440000000p-1f, 0x1.28b5000000000p-1f, 0x1.ed81e00000000p-27f, 0.0f},     {0x1.1d8f560000000p-1f, 0x1.2af1600000000p-1f, -0x1.7cdfa80000000p-28f, 0.0f},
    {0x1.1c52300000000p-1f, 0x1.2d2b400000000p-1f, -0x1.7448d80000000p-27f, 0.0f},     {0x1.1b17c60000000p-1f, 0x1.2f62ac0000000p-1f, -0x1.84f6ac0000000p-25f, 0.0f},
    {0x1.19e0120000000p-1f, 0x1.3197a00000000p-1f, 0x1.21ff9c0000000p-27f, 0.0f},     {0x1.18ab080000000p-1f, 0x1.33ca2c0000000p-1f, 0x1.65132a0000000p-30f, 0.0f},
    {0x1.1778a20000000p-1f, 0x1.35fa500000000p-1f, -0x1.ecc9160000000p-25f, 0.0f},     {0x1.1648d60000000p-1f, 0x1.3828100000000p-1f, -0x1.d478680000000p-25f, 0.0f},
    {0x1.151b9a0000000p-1f, 0x1.3a53740000000p-1f, 0x1.77af7e0000000p-27f, 0.0f},     {0x1.13f0e80000000p-1f, 0x1.3c7c800000000p-1f, 0x1.8773200000000p-25f, 0.0f},
    {0x1.12c8b80000000p-1f, 0x1.3ea33c0000000p-1f, -0x1.a14d0a0000000p-25f, 0.0f},     {0x1.11a3020000000p-1f, 0x1.40c7a40000000p-1f, -0x1.af918a0000000p-28f, 0.0f},
    {0x1.107fbc0000000p-1f, 0x1.42e9c80000000p-1f, -0x1.5e07f40000000p-25f, 0.0f},     {0x1.0f5ee00000000p-1f, 0x1.4509a40000000p-1f, 0x1.cceec20000000p-27f, 0.0f},
    {0x1.0e40660000000p-1f, 0x1.4727440000000p-1f, -0x1.4ac5540000000p-25f, 0.0f},     {0x1.0d24460000000p-1f, 0x1.4942a80000000p-1f, -0x1.dfa07e0000000p-26f, 0.0f},
    {0x1.0c0a780000000p-1f, 0x1.4b5bd80000000p-1f, -0x1.4523b20000000p-26f, 0.0f},     {0x1.0af2f80000000p-1f, 0x1.4d72d00000000p-1f, 0x1.fb9fd00000000p-25f, 0.0f},
    {0x1.09ddba0000000p-1f, 0x1.4f87a40000000p-1f, 0x1.8604de0000000p-26f, 0.0f},     {0x1.08cabc0000000p-1f, 0x1.519a4c0000000p-1f, -0x1.785cbc0000000p-25f, 0.0f},
    {0x1.07b9f20000000p-1f, 0x1.53aad00000000p-1f, 0x1.8999b80000000p-25f, 0.0f},     {0x1.06ab5a0000000p-1f, 0x1.55b9340000000p-1f, 0x1.ba817a0000000p-26f, 0.0f},
    {0x1.059eea0000000p-1f, 0x1.57c5800000000p-1f, -0x1.7d21ce0000000p-26f, 0.0f},     {0x1.04949c0000000p-1f, 0x1.59cfb40000000p-1f, -0x1.228bbc0000000p-28f, 0.0f},
    {0x1.038c6c0000000p-1f, 0x1.5bd7d40000000p-1f, -0x1.fd8e380000000p-25f, 0.0f},     {0x1.0286500000000p-1f, 0x1.5ddde40000000p-1f, 0x1.0149920000000p-25f, 0.0f},
    {0x1.0182440000000p-1f, 0x1.5fe1ec0000000p-1f, 0x1.e462480000000p-27f, 0.0f},     {0x1.0080400000000p-1f, 0x1.61e3f00000000p-1f, 0x1.a464660000000p-29f, 0.0f},
};
constexpr float LN2_HI = 0x1.62e4p-1f, LN2_LO = 1.428606765330187e-06f;
struct LogParts { float k, hi, sm; };
__device__ __forceinline__ bool normal_positive(float x) {
  return (uint32_t)(__float_as_int(x) - 0x00800000) < (uint32_t)(0x7f800000 - 0x00800000);
}
__device__ __forceinline__ LogParts log_parts(float x, const float4* tab) {
  const int bits = __float_as_int(x);
  const float4 t = tab[(bits >> 16) & 127];
  const float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
  const float r = fmaf(m, t.x, -1.0f);
  float q = fmaf(r, -0.25f, 1.0f / 3.0f);
  q = fmaf(r, q, -0.5f);
  LogParts o;
  o.k = (float)((bits >> 23) - 127);
  o.hi = t.y;
  o.sm = (t.z + r) + (r * r) * q;
  return o;
}
static __device__ __noinline__ void log_diffs_rare(float e0, float e1, float e2, float& da, float& db, float& dc) {
  da = logf(e0 / e1); db = logf(e0 / e2); dc = logf(e1 / e2);  // out of line: keeps the pixel pass short
}
__device__ __forceinline__ float log_diff(const LogParts& a, const LogParts& b) {
  const float dk = a.k - b.k, dh = a.hi - b.hi;
  const float big = fmaf(dk, LN2_HI, dh);
  const float err = fmaf(dk, LN2_HI, -big) + dh;  // exact: dk LN2_HI and dh are multiples of 2^-23 below 16
  return big + (err + ((a.sm - b.sm) + dk * LN2_LO));
}

constexpr int NSUB = 2;                  // 64-pixel sub-stages per stage: one barrier round trip per NSUB x 64 pixels
constexpr int SKB = 64 * NSUB;           // pixels per stage
constexpr int SNS = 4 / NSUB;            // stages in flight: B tiles in shared memory, A slots in tensor memory
constexpr int S_CHAIN = 16 / NSUB;       // stages per accumulation chain (1024 pixels, as above)
constexpr int SPR = NSUB == 1 ? 6 : 3;   // pixel ring slots
constexpr int S_A_COL0 = 256;            // D = columns 0-255, A slots behind it
constexpr int S_A_SLOT_COLS = SKB;       // SKB / 16 K steps x [hi 8 | lo 8] columns
constexpr int SB_KB_BYTES = 32 * 128;    // 4096: one core-matrix column (8 pixels) of all 256 rows
constexpr int SB_STAGE_BYTES = (SKB / 8) * SB_KB_BYTES;  // 32768
constexpr int ROW_G_HI = 0, ROW_B_HI = 64, ROW_G_LO = 128, ROW_B_LO = 192;  // B tile rows / D columns
static_assert(S_A_COL0 + SNS * S_A_SLOT_COLS <= TMEM_COLS, "TMEM budget");

struct PxSlotS {
  float a[SKB], b[SKB], c[SKB], siy[SKB];  // scaled log-chroma differences rg, rb, gb; sqrt(intensity)
};

struct SmemS {
  alignas(128) unsigned char bt[SNS][SB_STAGE_BYTES];  // 128 KB
  float acc[3][BINS][BINS + 1];
  PxSlotS px[SPR];
  float ctr[BINS];  // scaled midpoint centres
  alignas(16) float4 ltab[128];  // PH_LOGF_TABLE
  float red[PROD_WARPS];
  double red2[PROD_WARPS];
  int last_flag;
  alignas(8) uint64_t px_full[SPR], px_empty[SPR], ab_full[SNS], ab_empty[SNS], d_full, d_empty;
  uint32_t tmem_base;
};

template <int METHOD, bool FUSE_SSUM>
__global__ void __launch_bounds__(THREADS, 1) hist_fwd_sym_kernel(Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemS& S = *reinterpret_cast<SmemS*>(smem_raw);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < SPR; ++i) { mbar_init(&S.px_full[i], 1); mbar_init(&S.px_empty[i], PROD_WARPS); }
    for (int i = 0; i < SNS; ++i) { mbar_init(&S.ab_full[i], PROD_WARPS); mbar_init(&S.ab_empty[i], 1); }
    mbar_init(&S.d_full, 1);
    mbar_init(&S.d_empty, PROD_WARPS);
    fence_mbar_init();
  }
  if (tid < BINS) {
    const float cj = p.dom_u[tid], cm = p.dom_u[BINS - 1 - tid];
    S.ctr[tid] = 0.5f * (cj - cm) * p.coord_scale;
    // the caller asserted antisymmetric centres (PH_IMPL_MIRROR): never silently wrong if they are not
    if (!(fabsf(cj + cm) <= p.mirror_tol)) *reinterpret_cast<volatile int*>(p.status) = PH_ASYNC_MIRROR;
  }
  if (tid >= 128 && tid < 256) S.ltab[tid - 128] = PH_LOGF_TABLE[tid - 128];
  if (warp == MMA_WARP) tmem_alloc(&S.tmem_base, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = S.tmem_base;
  pdl_wait();
  pdl_launch_dependents();

  const int64_t first = blockIdx.x, step = gridDim.x;

  if (warp >= PX_WARP0) {
    // ===================== pixel pass: one warp per 64-pixel stage, two pixels per lane =====================
    const int me = warp - PX_WARP0;
    uint32_t it = 0;
    for (int64_t w = first; w < p.items; w += step) {
      const ItemRange ir = item_range(p, w);
      const int64_t b = ir.b;
      const uint32_t px0 = ir.px0, px1 = ir.px1;
      for (uint32_t base = px0; base < px1; base += SKB, ++it) {
        if ((int)(it % PXW) != me) continue;
        const int slot = it % SPR;
        constexpr int PPL = SKB / 32;  // pixels per lane
        float va[PPL], vb[PPL], vc[PPL], vs[PPL];
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
          const uint32_t px = base + k * 32 + lane;
          const bool valid = px < px1;
          float r = 0.f, g = 0.f, bl = 0.f;
          if (valid) {
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
          float da, db, dc;
          if (normal_positive(e0) && normal_positive(e1) && normal_positive(e2)) {
            const LogParts l0 = log_parts(e0, S.ltab), l1 = log_parts(e1, S.ltab), l2 = log_parts(e2, S.ltab);
            da = log_diff(l0, l1); db = log_diff(l0, l2); dc = log_diff(l1, l2);
          } else {  // images outside [-1, 1]: whatever logf of the ratio gives (NaN, inf), as in the exact-centre kernel
            log_diffs_rare(e0, e1, e2, da, db, dc);
          }
          va[k] = da * p.coord_scale;
          vb[k] = db * p.coord_scale;
          vc[k] = dc * p.coord_scale;
          // every product of two operands carries the intensity once; masked pixels contribute nothing
          vs[k] = valid ? sqrtf(iy) : 0.f;
          // operand = sqrt(Iy) x (scaled weight <= 2^14) must stay below fp16's 65504 — flagged, never silent
          if (vs[k] > tcgen::IY_OPERAND_LIMIT) *reinterpret_cast<volatile int*>(p.status) = PH_ASYNC_RANGE;
        }
        mbar_wait_relaxed(&S.px_empty[slot], ((it / SPR) & 1) ^ 1, 400);
        PxSlotS& o = S.px[slot];
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
          o.a[k * 32 + lane] = va[k]; o.b[k * 32 + lane] = vb[k]; o.c[k * 32 + lane] = vc[k];
          o.siy[k * 32 + lane] = vs[k];
        }
        mbar_arrive_warp(&S.px_full[slot]);
      }
    }
  } else if (warp < PROD_WARPS) {
    // ===================== operand producers + epilogue (16 warps, identical work) =====================
    const f32x2 wa2 = pack2(p.wa, p.wa), wb2 = pack2(p.wb, p.wb), mone2 = pack2(-1.0f, -1.0f);
    const int quad = warp & 3;   // TMEM lanes 32 quad ..: rows of alpha (quad 0, 1) or beta (quad 2, 3)
    const int r16 = warp >> 2;   // TMEM task: pixels 16 r16 .. + 15 of the stage = K step r16
    const int g8 = warp >> 1;    // gamma task: pixels 8 g8 .. + 7 = core-matrix column g8
    const int bin = (warp & 1) * 32 + lane;  // the same bin in both tasks
    const float c_bin = S.ctr[bin];
    const f32x2 negc = pack2(-c_bin, -c_bin);
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t row16 = (uint32_t)((bin >> 3) * 128 + (bin & 7) * 16);  // byte offset of row `bin` in a core-matrix column
    // Wrap-around counters (no % or /).  The stores of a stage are PUBLISHED one stage late: fence, tcgen05.wait::st and
    // the arrival on ab_full follow the next stage's arithmetic, so their latency (all 16 warps reach them together)
    // is hidden behind it; a chain's last stage is published at once.
    uint32_t slot = 0, slot_par = 0, stage = 0, stage_par = 0, chain = 0;
    int pend = -1;  // stage stored but not yet published (warp-uniform)
    for (int64_t w = first; w < p.items; w += step) {
      const ItemRange ir = item_range(p, w);
      const uint32_t nkb = (ir.px1 - ir.px0 + SKB - 1) / SKB;
      for (uint32_t kb = 0; kb < nkb; ++kb) {
        mbar_wait(&S.px_full[slot], slot_par);
        const PxSlotS& in = S.px[slot];
        unsigned char* tile = &S.bt[stage][0];
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
          ulonglong2 xx[4], iw[4], xc[2], ic[2];
          {
            const float* src = (quad < 2 ? in.a : in.b) + sub * 64 + r16 * 16;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              xx[q4] = *reinterpret_cast<const ulonglong2*>(src + q4 * 4);
              iw[q4] = *reinterpret_cast<const ulonglong2*>(&in.siy[sub * 64 + r16 * 16 + q4 * 4]);
            }
#pragma unroll
            for (int q2 = 0; q2 < 2; ++q2) {
              xc[q2] = *reinterpret_cast<const ulonglong2*>(&in.c[sub * 64 + g8 * 8 + q2 * 4]);
              ic[q2] = *reinterpret_cast<const ulonglong2*>(&in.siy[sub * 64 + g8 * 8 + q2 * 4]);
            }
          }
          if (sub == NSUB - 1) {
            mbar_arrive_warp(&S.px_empty[slot]);
            if (++slot == SPR) { slot = 0; slot_par ^= 1; }
          }
          uint32_t rr[16];  // this thread's row: [hi: pixel pairs 0..7 | lo: pixel pairs 0..7]
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            f32x2 w0, w1;
            weight4<METHOD>(xx[q4].x, xx[q4].y, negc, wa2, wb2, w0, w1);
            w0 = mul2(w0, iw[q4].x);
            w1 = mul2(w1, iw[q4].y);
            split_f16x2(w0, mone2, rr[2 * q4], rr[8 + 2 * q4]);
            split_f16x2(w1, mone2, rr[2 * q4 + 1], rr[8 + 2 * q4 + 1]);
          }
          uint4 ghi, glo;
          {
            f32x2 w0, w1, w2, w3;
            weight4<METHOD>(xc[0].x, xc[0].y, negc, wa2, wb2, w0, w1);
            weight4<METHOD>(xc[1].x, xc[1].y, negc, wa2, wb2, w2, w3);
            w0 = mul2(w0, ic[0].x); w1 = mul2(w1, ic[0].y); w2 = mul2(w2, ic[1].x); w3 = mul2(w3, ic[1].y);
            split_f16x2(w0, mone2, ghi.x, glo.x);
            split_f16x2(w1, mone2, ghi.y, glo.y);
            split_f16x2(w2, mone2, ghi.z, glo.z);
            split_f16x2(w3, mone2, ghi.w, glo.w);
          }
          if (sub == 0) {
            if (pend >= 0) {  // publish the previous stage
              fence_proxy_async_smem();
              tmem_st_wait();
              tc_fence_before_sync();
              mbar_arrive_warp(&S.ab_full[pend]);
            }
            // the MMAs that read this stage's A slot and B tile (SNS stages ago) are done
            mbar_wait(&S.ab_empty[stage], stage_par ^ 1);
            tc_fence_after_sync();
          }
          // K step 4 sub + r16 of the stage; core-matrix columns 8 sub + 2 r16 (+ 1) for beta, 8 sub + g8 for gamma
          tmem_st16(tmem + lane_addr + S_A_COL0 + stage * S_A_SLOT_COLS + (sub * 4 + r16) * 16, rr);
          if (quad >= 2) {  // beta is also the B operand of alpha . beta^T: rows 64 + bin (hi), 192 + bin (lo)
            unsigned char* col = tile + (sub * 8 + 2 * r16) * SB_KB_BYTES + row16;
            *reinterpret_cast<uint4*>(col + ROW_B_HI * 16) = make_uint4(rr[0], rr[1], rr[2], rr[3]);
            *reinterpret_cast<uint4*>(col + ROW_B_HI * 16 + SB_KB_BYTES) = make_uint4(rr[4], rr[5], rr[6], rr[7]);
            *reinterpret_cast<uint4*>(col + ROW_B_LO * 16) = make_uint4(rr[8], rr[9], rr[10], rr[11]);
            *reinterpret_cast<uint4*>(col + ROW_B_LO * 16 + SB_KB_BYTES) = make_uint4(rr[12], rr[13], rr[14], rr[15]);
          }
          *reinterpret_cast<uint4*>(tile + (sub * 8 + g8) * SB_KB_BYTES + row16 + ROW_G_HI * 16) = ghi;
          *reinterpret_cast<uint4*>(tile + (sub * 8 + g8) * SB_KB_BYTES + row16 + ROW_G_LO * 16) = glo;
        }
        pend = (int)stage;
        if (++stage == SNS) { stage = 0; stage_par ^= 1; }

        const bool chain_end = ((kb + 1) % S_CHAIN == 0) || (kb + 1 == nkb);
        if (!chain_end) continue;
        fence_proxy_async_smem();
        tmem_st_wait();
        tc_fence_before_sync();
        mbar_arrive_warp(&S.ab_full[pend]);
        pend = -1;
        // ---- chain epilogue: D (128 rows x [. gamma_hi | . beta_hi | . gamma_lo | . beta_lo]) += into the fp32 accumulator.
        //      Warp (quad, r16): rows 32 quad .., columns 32 r16 .. + 31 of the hi half and of the lo half. ----
        mbar_wait(&S.d_full, chain & 1);
        ++chain;
        tc_fence_after_sync();
        const bool first_chain = kb < S_CHAIN;
        const bool used = quad < 2 || r16 < 2;  // the beta . beta^T quadrant is not a histogram
        float val[32];
        if (used) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t v1[16], v2[16];
            tmem_ld16(tmem + lane_addr + r16 * 32 + h * 16, v1);
            tmem_ld16(tmem + lane_addr + ROW_G_LO + r16 * 32 + h * 16, v2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) val[h * 16 + i] = __uint_as_float(v1[i]) + __uint_as_float(v2[i]);
          }
        }
        tc_fence_before_sync();
        mbar_arrive_warp(&S.d_empty);  // the accumulators are in registers: the next chain may start
        if (used) {
          // row of D = bin i' of alpha (quad < 2) or beta; column = bin j of gamma (r16 < 2) or beta
          float* dst;
          int jstride;
          if (quad >= 2)      { dst = &S.acc[2][BINS - 1 - r16 * 32][BINS - 1 - bin]; jstride = -(BINS + 1); }  // H_B[63-i'][63-j]
          else if (r16 < 2)   { dst = &S.acc[1][r16 * 32][BINS - 1 - bin];            jstride = BINS + 1; }     // H_G[63-i'][j]
          else                { dst = &S.acc[0][(r16 - 2) * 32][bin];                 jstride = BINS + 1; }     // H_R[i'][j]
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float* a = dst + i * jstride;
            *a = first_chain ? val[i] : *a + val[i];
          }
        }
        if (kb + 1 != nkb) continue;  // every (c, j, i) is owned by one thread: no barrier between chains
        finish_item<FUSE_SSUM>(S, p, ir, tid, warp, lane);
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issue: uniform loop, one elected lane issues =====================
    constexpr uint32_t IDESC_FULL = idesc_f16(128, 256);  // . [gamma_hi | beta_hi | gamma_lo | beta_lo]
    constexpr uint32_t IDESC_HI = idesc_f16(128, 128);    // . [gamma_hi | beta_hi] (the first 128 rows of the tile)
    const uint64_t desc0 = smem_desc_kmajor_noswizzle(smem_u32(&S.bt[0][0]), SB_KB_BYTES, 128);
    const uint32_t dlo0 = (uint32_t)desc0, dhi = (uint32_t)(desc0 >> 32);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);  // provably uniform copy
    uint32_t stage = 0, phase = 0, chain_par = 0;
    for (int64_t w = first; w < p.items; w += step) {
      const ItemRange ir = item_range(p, w);
      const uint32_t nkb = (ir.px1 - ir.px0 + SKB - 1) / SKB;
      for (uint32_t kb0 = 0; kb0 < nkb; kb0 += S_CHAIN) {
        const uint32_t n_this = min((uint32_t)S_CHAIN, nkb - kb0);
        for (uint32_t k = 0; k < n_this; ++k) {
          mbar_wait(&S.ab_full[stage], phase);
          tc_fence_after_sync();
          const uint32_t dstage = dlo0 + stage * (SB_STAGE_BYTES >> 4);
          const uint32_t acol = tm + S_A_COL0 + stage * S_A_SLOT_COLS;
#pragma unroll
          for (int ks = 0; ks < SKB / 16; ++ks) {
            const uint32_t b_d = dstage + ((ks * 2 * SB_KB_BYTES) >> 4);
            const uint32_t acc0 = (k == 0 && ks == 0) ? 0u : 1u;
            if (elect_one_sync()) {
              mma_f16_ts2(tm, acol + ks * 16, b_d, dhi, IDESC_FULL, acc0);
              mma_f16_ts2(tm, acol + ks * 16 + 8, b_d, dhi, IDESC_HI, 1u);
            }
          }
          if (elect_one_sync()) mma_commit(&S.ab_empty[stage]);
          if (++stage == SNS) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) mma_commit(&S.d_full);
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
  __shared__ int s_n;
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const float* img = image + b * npix * channels;
  for (int i = tid; i < DEDUP_SLOTS; i += 256) { owner[i] = -1; count[i] = 0; }
  if (tid == 0) s_n = 0;
  __syncthreads();
  pdl_wait();
  pdl_launch_dependents();
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
  // The list is written in the order of the colours' bit patterns (rank by lexicographic (r, g, b) comparison: the
  // colours are distinct, so the ranks are a permutation), not in the order in which warps happened to claim slots:
  // the order of the (colour, count) terms — and with it every bit of the contracted histogram — is then the same
  // from run to run.  n <= 512 entries, a few dozen for a sprite.
  __shared__ uint4 lst[DEDUP_MAX];
  __shared__ int s_out;
  if (tid == 0) s_out = 0;
  __syncthreads();
  for (int i = tid; i < DEDUP_SLOTS; i += 256) {
    const int o = owner[i];
    if (o >= 0) {
      const float* op = img + (int64_t)o * channels;
      lst[atomicAdd(&s_out, 1)] = make_uint4(__float_as_uint(__ldg(op)), __float_as_uint(__ldg(op + 1)),
                                             __float_as_uint(__ldg(op + 2)), (unsigned)count[i]);
    }
  }
  __syncthreads();
  float4* out = ulist + b * DEDUP_MAX;
  for (int k = tid; k < n; k += 256) {
    const uint4 me = lst[k];
    int rank = 0;
    for (int e = 0; e < n; ++e) {
      const uint4 q = lst[e];
      rank += (q.x < me.x || (q.x == me.x && (q.y < me.y || (q.y == me.y && q.z < me.z)))) ? 1 : 0;
    }
    out[rank] = make_float4(__uint_as_float(me.x), __uint_as_float(me.y), __uint_as_float(me.z), (float)me.w);
  }
  if (tid == 0) nunique[b] = n;
}

// =============================================================================================
// host side
// =============================================================================================
bool tc_supported(int64_t npix, int bins, int method) {
  (void)method;
  return bins >= 64 && bins <= 1024 && bins % 64 == 0 && npix >= 1;
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

static size_t tail_bytes(const FwdPlan& pl, int64_t batch) {
  const int64_t n_tail = batch - pl.n_whole;
  if (n_tail == 0) return 0;
  return align_up((size_t)n_tail * pl.splits * 3 * 64 * 64 * sizeof(float), 256) + align_up((size_t)n_tail * sizeof(int), 256);
}

size_t tc_workspace_bytes(int64_t batch, int64_t npix, int bins) {
  if (bins < 64 || bins % 64 != 0) return 0;
  const int64_t nb = bins / 64;
  // forward: [unique-colour lists] [slice partials + arrival counters] [raw 64 x 64 blocks when bins > 64]
  size_t fwd = tail_bytes(tc_fwd_plan(batch, npix, false), batch);
  if (tc_fwd_plan(batch, npix, true).n_whole == batch) fwd += dedup_bytes(batch);
  if (nb > 1) fwd += align_up((size_t)batch * nb * nb * 3 * 64 * 64 * sizeof(float), 256);
  if (bins == 256) {  // dedicated 256-bin kernel (hist_tc_fwd256.cu): [unique-colour lists][per-item partial sums]
    const size_t f256 = dedup_bytes(batch) + tc_fwd256_workspace_bytes(batch, npix);
    if (f256 > fwd) fwd = f256;
  }
  const size_t bwd = tc_bwd_workspace_bytes(batch, bins);
  return align_up(fwd > bwd ? fwd : bwd, 256) + 256;
}

// out[b, I, J, c] = raw[block(I/64, J/64)][b][c][I%64][J%64] / D_b,  D_b = sum of every block of image b
__global__ void __launch_bounds__(256) hist_block_finalize_kernel(const float* __restrict__ raw, int64_t batch, int nb,
                                                                  float* __restrict__ hist, float* __restrict__ denom) {
  __shared__ double scratch[32];
  const int64_t b = blockIdx.x;
  const int bins = nb * 64;
  const int64_t blk_stride = batch * (int64_t)(3 * 64 * 64);
  const float* mine = raw + b * (int64_t)(3 * 64 * 64);
  double acc = 0.0;
  for (int blk = 0; blk < nb * nb; ++blk)
    for (int e = threadIdx.x; e < 3 * 64 * 64; e += 256) acc += (double)__ldg(mine + blk * blk_stride + e);
  const float d = (float)block_sum(acc, scratch);
  if (threadIdx.x == 0) denom[b] = d;
  const float inv_d = 1.0f / d;
  float* out = hist + b * (int64_t)bins * bins * 3;
  for (int e = threadIdx.x; e < bins * bins * 3; e += 256) {
    const int c = e % 3, ij = e / 3, i_full = ij / bins, j_full = ij - i_full * bins;
    const int blk = (i_full >> 6) * nb + (j_full >> 6);
    out[e] = __ldg(mine + blk * blk_stride + c * 4096 + (i_full & 63) * 64 + (j_full & 63)) * inv_d;
  }
}

int tc_hist_forward(const float* image, int64_t batch, int64_t npix, int channels, const float* dom, int bins,
                    int method, float sigma_sqr, float eps, float* hist, float* denom, void* workspace, bool dedup,
                    bool mirror, const float* hist_true, double* ssum, cudaStream_t st) {
  using namespace fwdtc;
  PH_CHECK_ARG(bins >= BINS && bins % BINS == 0, "tensor-core forward needs a multiple of 64 bins");
  const int nb = bins / BINS;
  static const bool fwd256_off = getenv("PH_FWD256") && atoi(getenv("PH_FWD256")) == 0;  // tuning knob: block path
  if (bins == 256 && !fwd256_off) {
    // the whole 256 x 256 histogram per CTA (tensor-pipe bound) instead of 16 blocks of 64 x 64
    char* ws = static_cast<char*>(workspace);
    const float4* ulist = nullptr;
    const int* nunique = nullptr;
    float iy_scale = 1.0f;
    size_t off = 0;
    if (dedup) {
      int e = 0;
      while (((int64_t)1 << e) < npix) ++e;
      iy_scale = ldexpf(1.0f, -e);
      float4* ul = reinterpret_cast<float4*>(ws);
      int* nu = reinterpret_cast<int*>(ws + align_up((size_t)batch * DEDUP_MAX * sizeof(float4), 256));
      if (batch > 0) {
        hist_dedup_kernel<<<(unsigned)batch, 256, 0, st>>>(image, npix, channels, ul, nu);
        PH_LAUNCH_OK("hist_dedup_kernel");
      }
      ulist = ul;
      nunique = nu;
      off = dedup_bytes(batch);
    }
    const int rc = tc_fwd256_forward(image, batch, npix, channels, dom, method, sigma_sqr, eps, ulist, nunique,
                                     DEDUP_MAX, iy_scale, hist, denom, ws + off, st);
    if (rc != PH_OK) return rc;
    if (ssum != nullptr) return launch_hellinger_ssum_accumulate(hist_true, hist, batch * (int64_t)bins * bins * 3, ssum, st);
    return PH_OK;
  }
  Params p{};
  p.image = image;
  p.hist = hist;
  p.denom = denom;
  p.npix = npix;
  p.channels = channels;
  const FwdPlan pl = tc_fwd_plan(batch, npix, dedup);
  // mirrored-tile kernel: dense 64-bin batches whose centres the caller declared antisymmetric (PH_IMPL_MIRROR)
  static const bool sym_off = getenv("PH_FWD_SYM") && atoi(getenv("PH_FWD_SYM")) == 0;  // tuning knob
  // images of fewer than 1024 pixels keep the exact centres: the centre mismatch averages out over the pixels (3.5e-6 at
  // 32 x 32, 2.1e-6 at 64 x 64 against 5e-6 with a 1e-5 peak error for 8 x 8 or 20 x 12 images), and they cost nothing
  const bool sym = mirror && !dedup && nb == 1 && npix >= 1024 && !sym_off;
  p.n_whole = pl.n_whole;
  p.splits = pl.splits;
  p.px_per_split = ceil_div(ceil_div(npix, p.splits), sym ? SKB : KB) * (sym ? SKB : KB);
  // rounding the slice length up to whole stages can leave the last slices without pixels (4096 pixels in 10 slices of
  // 512): an item must own at least one stage, so the slice count follows the rounded length
  p.splits = (int)ceil_div(npix, p.px_per_split);
  p.items = p.n_whole + (batch - p.n_whole) * p.splits;
  p.eps = eps;
  const tcgen::WeightScales wsc = tcgen::weight_scales(method, sigma_sqr);
  p.coord_scale = wsc.coord_scale;
  p.wa = wsc.wa;
  p.wb = wsc.wb;
  const double weight_scale = wsc.weight_scale;  // the generated weights are weight_scale * K
  p.iy_scale = 1.0f;
  char* ws = static_cast<char*>(workspace);
  size_t off = 0;
  if (dedup && p.n_whole == batch) {
    // multiplicities reach npix: keep count * Iy * 2^14 K below fp16's maximum
    int e = 0;
    while (((int64_t)1 << e) < npix) ++e;
    p.iy_scale = ldexpf(1.0f, -e);
    // unique colours + multiplicities per image (only worth it when a CTA owns whole images)
    float4* ulist = reinterpret_cast<float4*>(ws);
    int* nunique = reinterpret_cast<int*>(ws + align_up((size_t)batch * DEDUP_MAX * sizeof(float4), 256));
    PH_CUDA_OK(launch_pdl(hist_dedup_kernel, dim3((unsigned)batch), dim3(256), 0, st, image, npix, channels, ulist, nunique));
    PH_LAUNCH_OK("hist_dedup_kernel");
    p.ulist = ulist;
    p.nunique = nunique;
    off += dedup_bytes(batch);
  }
  p.inv_scale = (float)(1.0 / (weight_scale * weight_scale * (double)p.iy_scale));
  p.inv_scale_dense = (float)(1.0 / (weight_scale * weight_scale));
  p.status = async_status_word();
  PH_CHECK_ARG(p.status != nullptr, "no mapped status word (cudaHostAlloc failed)");
  p.mirror_tol = 2.5e-5f * sqrtf(sigma_sqr);  // the flag's contract is 2e-5 sigma (palhist.h)
  const int64_t n_tail = batch - p.n_whole;
  if (n_tail > 0) {
    p.partial = reinterpret_cast<float*>(ws + off);
    p.tail_counter = reinterpret_cast<int*>(ws + off + align_up((size_t)n_tail * p.splits * 3 * BINS * BINS * sizeof(float), 256));
    off += tail_bytes(pl, batch);
  }
  float* raw = nb > 1 ? reinterpret_cast<float*>(ws + off) : nullptr;
  const bool fuse = ssum != nullptr && nb == 1;
  p.hist_true = fuse ? hist_true : nullptr;
  p.ssum = fuse ? ssum : nullptr;
  static const bool a_in_smem = getenv("PH_FWD_A") && getenv("PH_FWD_A")[0] == 's';  // tuning knob: PH_FWD_A=smem
  const size_t smem = sym ? sizeof(SmemS) : (a_in_smem ? sizeof(Smem) : sizeof(SmemA));
  int grid = cached_sm_count();
  if (grid > p.items) grid = (int)p.items;
  void (*kern)(Params) = nullptr;
  if (sym) {
    if (method == PH_METHOD_INVERSE_QUADRATIC)
      kern = fuse ? hist_fwd_sym_kernel<PH_METHOD_INVERSE_QUADRATIC, true> : hist_fwd_sym_kernel<PH_METHOD_INVERSE_QUADRATIC, false>;
    else
      kern = fuse ? hist_fwd_sym_kernel<PH_METHOD_RBF, true> : hist_fwd_sym_kernel<PH_METHOD_RBF, false>;
  } else if (a_in_smem) {
    if (method == PH_METHOD_INVERSE_QUADRATIC)
      kern = fuse ? hist_fwd_tc_kernel<PH_METHOD_INVERSE_QUADRATIC, true> : hist_fwd_tc_kernel<PH_METHOD_INVERSE_QUADRATIC, false>;
    else
      kern = fuse ? hist_fwd_tc_kernel<PH_METHOD_RBF, true> : hist_fwd_tc_kernel<PH_METHOD_RBF, false>;
  } else {
    if (method == PH_METHOD_INVERSE_QUADRATIC)
      kern = fuse ? hist_fwd_tca_kernel<PH_METHOD_INVERSE_QUADRATIC, true> : hist_fwd_tca_kernel<PH_METHOD_INVERSE_QUADRATIC, false>;
    else
      kern = fuse ? hist_fwd_tca_kernel<PH_METHOD_RBF, true> : hist_fwd_tca_kernel<PH_METHOD_RBF, false>;
  }
  PH_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // one launch per 64 x 64 block of the histogram (a single one at 64 bins)
  for (int blk = 0; blk < nb * nb; ++blk) {
    p.dom_u = dom + (blk / nb) * BINS;
    p.dom_v = dom + (blk % nb) * BINS;
    p.raw_out = raw ? raw + (int64_t)blk * batch * (3 * BINS * BINS) : nullptr;
    if (n_tail > 0) PH_CUDA_OK(cudaMemsetAsync(p.tail_counter, 0, (size_t)n_tail * sizeof(int), st));
    if (a_in_smem && !sym) {
      kern<<<grid, THREADS, smem, st>>>(p);
    } else {
      PH_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(THREADS), smem, st, p));
    }
    PH_LAUNCH_OK("hist_fwd_tc_kernel");
  }
  if (nb > 1) {
    hist_block_finalize_kernel<<<(unsigned)batch, 256, 0, st>>>(raw, batch, nb, hist, denom);
    PH_LAUNCH_OK("hist_block_finalize_kernel");
    if (ssum != nullptr) return launch_hellinger_ssum_accumulate(hist_true, hist, batch * (int64_t)bins * bins * 3, ssum, st);
  }
  return PH_OK;
}

}  // namespace ph
