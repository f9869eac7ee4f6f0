// RGB-uv histogram forward with 256 bins (cfgE: 256 x 256 images, SURVEY.md §8 "scale sweep"), tensor-core
// engine: the regime where the Ku^T.Kv contraction (histogram.py:29-30) is bound by the tensor pipe instead of
// by the generation of the operands.
//
// The 64-bin kernel of hist_tc.cu covers 256 bins as 16 blocks of 64 x 64, regenerating the 64 + 64 weights of
// every pixel for every block (4-fold redundant generation).  Here one CTA contracts the whole 256 x 256
// histogram of ONE channel at a time: the accumulator is 256 rows (u-bins) x 256 columns (v-bins) of fp32 =
// two M=128 halves x 256 TMEM columns = all 512 columns of the SM's tensor memory, and every generated weight
// is used against 256 others (512 algorithmic FLOP per weight instead of 128).
//
//   per 32-pixel stage and channel:  A_hi, A_lo (256 u-bins x 32 pixels, Iy-weighted), B_hi, B_lo (256 v-bins x 32)
//   as fp16 K-major no-swizzle tiles in shared memory (64 KB per stage, 2 stages); the MMA warp issues, per K step
//   of 16 pixels and half h of the u-bins, D[h] += A_hi[h].B_hi + A_hi[h].B_lo + A_lo[h].B_hi  (tcgen05.mma
//   kind::f16, SS, M128 N256 K16: 12 instructions x 128 cycles per stage = the tensor pipe's full rate; the fp16 hi+lo split with three
//   products carries ~22 significant bits, hist_tc.cu / DESIGN.md §5).
//
//   warps 0-7 generate the A side, 8-15 the B side (thread = four bins x 8 pixels per stage), warp 16 issues the
//   MMAs, warps 17-23 are the pixel pass (128-bit loads, float64 log-chroma u/v of the current channel as hi + lo
//   pairs and Iy into a shared-memory ring; seven warps because the float64 logs are a long dependent chain).  The three channels of an item run one after the other (the pixel pass re-reads the
//   16 B pixel per channel: 3 MB per 256 x 256 image against 77 GFLOP).
//
//   Accuracy: as in the 64-bin kernel the accumulation chains are cut, here every 512 pixels; the 16 producer warps
//   then drain the 64 K accumulators (tcgen05.ld) through a shared-memory staging buffer into the item's fp32
//   partial sums in global memory ([channel][v-bin][u-bin]): a bulk copy for the first chain, bulk fp32 reductions
//   (cp.reduce.async.bulk: the additions happen at the L2, asynchronously) afterwards — every address is owned by
//   one CTA and the operations of successive chains are ordered, so the sums are deterministic.  `hist256_finalize_kernel` adds the slices of an image
//   in order, computes the normaliser D (histogram.py:77-79) and writes H / D channel-last (B,256,256,3).
#include <stdlib.h>

#include "common.cuh"
#include "hist_internal.cuh"
#include "hist_tc_gen.cuh"
#include "tc_ptx.cuh"

namespace ph {

using namespace tc;

namespace fwd256 {

using tcgen::named_bar_sync;
using tcgen::weight2;
using tcgen::weight4;

constexpr int BINS = 256;
constexpr int KB = 32;         // pixels per stage = two K steps of the instruction (= one pixel-ring slot): the
                               // per-stage overhead of a producer warp (barrier waits / arrivals, proxy fence, address
                               // arithmetic, ~110 of its instructions) is paid once per 32 weights of a thread
constexpr int SLOT_PX = KB;
constexpr int NS = 2;          // operand stages (64 KB each)
constexpr int CHAIN_KB = 16;   // stages per TMEM accumulation chain (512 pixels = 96 accumulating MMAs per half)
constexpr int A_WARPS = 8, B_WARPS = 8, PXW = 7;  // 24 warps = 6 per scheduler at <= 80 registers
constexpr int PROD_WARPS = A_WARPS + B_WARPS;
constexpr int MMA_WARP = PROD_WARPS;            // 16
constexpr int PX_WARP0 = MMA_WARP + 1;          // 17
constexpr int PR = 8;                           // pixel ring slots
constexpr int THREADS = (PX_WARP0 + PXW) * 32;  // 768
constexpr int TMEM_COLS = 512;
// one operand part (hi or lo) of one side for one stage: 256 rows x 32 pixels of fp16, K-major, no swizzle:
//   [kcol = pixel / 8 (4)][row / 8 (32)][row % 8][pixel % 8]     core matrix = 8 rows x 8 halfs = 128 B
constexpr int KCOL_BYTES = 32 * 128;          // 4096: LBO, the distance of the two core-matrix columns of a K step
constexpr int TILE_BYTES = 4 * KCOL_BYTES;    // 16384
constexpr int STAGE_BYTES = 4 * TILE_BYTES;   // A_hi | A_lo | B_hi | B_lo = 65536
constexpr int HIST_ELEMS = 3 * BINS * BINS;   // per item: [c][j][i]

struct PxSlot {
  float u[SLOT_PX], v[SLOT_PX], ul[SLOT_PX], vl[SLOT_PX], iy[SLOT_PX];  // coordinates as hi + lo pairs
};

constexpr int DRAIN_COLS = 32;                       // v-bins j per drain chunk: 32 rows of 256 u-bins = 32 KB
constexpr int DRAIN_CHUNKS = BINS / DRAIN_COLS;      // 8
constexpr int DRAIN_BYTES = DRAIN_COLS * BINS * 4;

struct Smem {
  alignas(128) unsigned char ab[NS][STAGE_BYTES];  // 128 KB
  alignas(128) float stage_out[2][DRAIN_COLS * BINS];  // 64 KB: chain drain staging [j][i], double-buffered
  PxSlot px[PR];
  float dom[BINS];
  alignas(8) uint64_t px_full[PR], px_empty[PR], ab_full[NS], ab_empty[NS], d_full, d_empty;
  uint32_t tmem_base;
};

struct Params {
  const float* image;
  const float* dom;      // 256 bin centres (both sides)
  float* partial;        // (items, 3, 256, 256) raw scaled sums, [c][j][i]
  const float4* ulist;   // optional (B, dedup_max): unique colours (r,g,b,count) of each image, or NULL
  const int* nunique;    // optional (B): number of unique colours, < 0 = image not de-duplicated
  int dedup_max;
  int64_t npix;
  int channels;
  int splits;            // pixel slices per image (a de-duplicated image uses slice 0 only)
  int64_t px_per_split;  // multiple of SLOT_PX
  int64_t items;         // B * splits
  float eps;
  float wa, wb, coord_scale, iy_scale;  // as in hist_tc.cu
  int* status;                          // as in hist_tc.cu
};

struct ItemRange {
  int64_t b;
  uint32_t px0, px1;
  bool dedup;
};
__device__ __forceinline__ ItemRange item_range(const Params& p, int64_t w) {
  ItemRange r;
  r.b = w / p.splits;
  const uint32_t split = (uint32_t)(w - r.b * p.splits);
  r.px0 = split * (uint32_t)p.px_per_split;
  r.px1 = min(r.px0 + (uint32_t)p.px_per_split, (uint32_t)p.npix);
  r.dedup = false;
  if (p.nunique != nullptr) {
    const int nu = __shfl_sync(0xffffffffu, __ldg(p.nunique + r.b), 0);
    // a de-duplicated image is one short list: slice 0 contracts it, the other slices of the image stay empty
    if (nu >= 0) { r.px0 = 0; r.px1 = split == 0 ? (uint32_t)nu : 0u; r.dedup = true; }
  }
  return r;
}

template <int METHOD>
__global__ void __launch_bounds__(THREADS, 1) hist_fwd256_tc_kernel(Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < PR; ++i) { mbar_init(&S.px_full[i], 1); mbar_init(&S.px_empty[i], PROD_WARPS); }
    for (int i = 0; i < NS; ++i) { mbar_init(&S.ab_full[i], PROD_WARPS); mbar_init(&S.ab_empty[i], 1); }
    mbar_init(&S.d_full, 1);
    mbar_init(&S.d_empty, PROD_WARPS);
    fence_mbar_init();
  }
  if (tid < BINS) S.dom[tid] = p.dom[tid] * p.coord_scale;
  if (warp == MMA_WARP) tmem_alloc(&S.tmem_base, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = S.tmem_base;

  const int64_t first = blockIdx.x, step = gridDim.x;

  if (warp >= PX_WARP0) {
    // ===================== pixel pass: one 32-pixel slot of the current channel per round =====================
    const int me = warp - PX_WARP0;
    uint32_t sit = 0;
    for (int64_t w = first; w < p.items; w += step) {
      const ItemRange ir = item_range(p, w);
      for (int c = 0; c < 3; ++c) {
        for (uint32_t base = ir.px0; base < ir.px1; base += SLOT_PX, ++sit) {
          if ((int)(sit % PXW) != me) continue;
          const int slot = sit % PR;
          const uint32_t px = base + lane;
          float r = 0.f, g = 0.f, bl = 0.f, mult = 1.f;
          const bool valid = px < ir.px1;
          if (valid && ir.dedup) {
            const float4 q = __ldg(p.ulist + ir.b * p.dedup_max + px);
            r = q.x; g = q.y; bl = q.z; mult = q.w;
          } else if (valid) {
            const float* src = p.image + (ir.b * p.npix + px) * p.channels;
            if (p.channels == 4) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(src));
              r = q.x; g = q.y; bl = q.z;
            } else {
              r = __ldg(src); g = __ldg(src + 1); bl = __ldg(src + 2);
            }
          }
          // histogram.py:58-66, :13-17, :72-74.  The log-chroma differences are taken in float64 and carried as
          // float hi + lo pairs into every (u - c), as the backward kernels do (common.cuh): seven pixel warps have
          // the time for it, one more packed add per weight pair in the producers, and the 256-bin gradients land
          // 2e-6 closer to the float64 oracle (the forward's histogram error feeds sqrt(Ht / Hp) of the backward)
          const PixelTerms pt = pixel_terms(r, g, bl, p.eps);
          float u, ul, v, vl;
          channel_uv(pt, c, u, ul, v, vl);
          const float iy = pt.iy, cs = p.coord_scale;  // power of two: exact for both parts
          mbar_wait_relaxed(&S.px_empty[slot], ((sit / PR) & 1) ^ 1, 400);
          PxSlot& o = S.px[slot];
          o.u[lane] = u * cs;
          o.v[lane] = v * cs;
          o.ul[lane] = ul * cs;
          o.vl[lane] = vl * cs;
          // masked pixels contribute nothing (A operand = 0).  The intensity scale that keeps count * Iy * K below
          // fp16's maximum applies to de-duplicated images only: a dense image of the same batch (more than 512 colours)
          // would lose its far-bin weights to fp16 subnormals under 2^-16 — any per-image power of two cancels in H / D
          const float a_iy = valid ? iy * mult * (ir.dedup ? p.iy_scale : 1.0f) : 0.f;
          if (a_iy > tcgen::IY_OPERAND_LIMIT) *reinterpret_cast<volatile int*>(p.status) = PH_ASYNC_RANGE;  // see hist_tc.cu
          o.iy[lane] = a_iy;
          mbar_arrive_warp(&S.px_full[slot]);
        }
      }
    }
  } else if (warp < PROD_WARPS) {
    // ===================== operand producers (A: warps 0-7, B: warps 8-15) + chain drain =====================
    // thread = (side, 8-pixel core-matrix column kc of the stage, bins b, b + 64, b + 128, b + 192): the 8 coordinates
    // it loads serve four bins — ncu shows this kernel at 94 % of the shared-memory bandwidth (the MMAs alone read 96 B
    // per clock), so every broadcast load that is not issued counts; rows b of a warp stay contiguous (conflict-free
    // 16-byte stores)
    const int side = warp >> 3;        // warp-uniform
    const int kc = (warp >> 1) & 3;    // warp-uniform: pixels [8 kc, 8 kc + 8) of the stage
    const int bin0 = tid & 63;
    f32x2 negc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float cb = S.dom[bin0 + 64 * q]; negc[q] = pack2(-cb, -cb); }
    const f32x2 wa2 = pack2(p.wa, p.wa), wb2 = pack2(p.wb, p.wb), mone2 = pack2(-1.0f, -1.0f);
    // row of bin0 in core-matrix column kc (bin0 + 64 q: 8 q row groups = 1024 q bytes further)
    const uint32_t row_off = (uint32_t)(side * 2 * TILE_BYTES + kc * KCOL_BYTES + (bin0 >> 3) * 128 + (bin0 & 7) * 16);
    // drain role: TMEM sub-partition quad = warp % 4 (lanes 32 quad ..), cg = warp / 4: u-bin half h = cg % 2
    // (accumulator D[h] = columns 256 h ..), and within every 32-column chunk the columns 16 (cg / 2) .. + 15
    const int quad = warp & 3, cg = warp >> 2;
    const int dh = cg & 1, jhalf = cg >> 1;
    const int i_row = dh * 128 + quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    uint32_t it = 0, chain = 0;
    // ---- chain drain: D (TMEM) -> registers -> shared-memory staging [j][i] -> bulk copy (first chain) or bulk fp32
    //      reduction (later chains) into the item's partial sums [c][j][i] in global memory.  The TMA engine performs
    //      the 64 K additions of a chain at the L2 while the SM is back at the MMAs (with per-lane REDs the 16 warps
    //      sat in the memory-pipeline throttle: ncu lg_throttle 1.1 per issue, now 0).
    auto drain = [&](float* dst, bool first_chain) {
      mbar_wait(&S.d_full, chain & 1);
      ++chain;
      tc_fence_after_sync();
      // every earlier bulk operation of this CTA has been performed (they were issued a whole chain ago): the
      // additions to an address happen in chain order, so the sums are deterministic
      if (tid == 0) bulk_wait_group<0>();
#pragma unroll 1
      for (int ch = 0; ch < DRAIN_CHUNKS; ++ch) {
        uint32_t vals[16];
        tmem_ld16(tmem + lane_addr + dh * 256 + ch * DRAIN_COLS + jhalf * 16, vals);
        tmem_ld_wait();
        if (ch == DRAIN_CHUNKS - 1) {
          tc_fence_before_sync();
          mbar_arrive_warp(&S.d_empty);  // everything of this chain has left TMEM: the next chain may start
        }
        named_bar_sync(5, PROD_WARPS * 32);  // staging buffer ch % 2 is free (thread 0 waited for its last reader)
        float* so = &S.stage_out[ch & 1][(jhalf * 16) * BINS + i_row];
#pragma unroll
        for (int k = 0; k < 16; ++k) so[k * BINS] = __uint_as_float(vals[k]);
        fence_proxy_async_smem();
        named_bar_sync(5, PROD_WARPS * 32);  // the chunk is complete in shared memory
        if (tid == 0) {
          float* g = dst + (int64_t)ch * (DRAIN_COLS * BINS);
          if (first_chain) bulk_store_s2g(g, &S.stage_out[ch & 1][0], DRAIN_BYTES);
          else bulk_reduce_add_f32_s2g(g, &S.stage_out[ch & 1][0], DRAIN_BYTES);
          bulk_commit_group();
          bulk_wait_group_read<1>();  // the operation issued one chunk ago has read its buffer = the next one's
        }
      }
    };
    float* pend_dst = nullptr;  // accumulators of a finished chain still to be drained
    bool pend_first = false;
    for (int64_t w = first; w < p.items; w += step) {
      const ItemRange ir = item_range(p, w);
      const uint32_t nkb = (ir.px1 - ir.px0 + KB - 1) / KB;
      float* item_out = p.partial + w * (int64_t)HIST_ELEMS;
      for (int c = 0; c < 3; ++c) {
        for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
          const int slot = it % PR, stage = it % NS;
          mbar_wait(&S.px_full[slot], (it / PR) & 1);
          const PxSlot& in = S.px[slot];
          const float* src = (side == 0 ? in.u : in.v) + 8 * kc;
          const float* srl = (side == 0 ? in.ul : in.vl) + 8 * kc;
          ulonglong2 xx[2], xl[2], iw[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            xx[q] = *reinterpret_cast<const ulonglong2*>(src + 4 * q);
            xl[q] = *reinterpret_cast<const ulonglong2*>(srl + 4 * q);
          }
          if (side == 0) {
#pragma unroll
            for (int q = 0; q < 2; ++q) iw[q] = *reinterpret_cast<const ulonglong2*>(&in.iy[8 * kc + 4 * q]);
          }
          mbar_arrive_warp(&S.px_empty[slot]);
          mbar_wait(&S.ab_empty[stage], ((it / NS) & 1) ^ 1);  // the MMAs that read this stage are done
          unsigned char* tile = &S.ab[stage][row_off];
#pragma unroll
          for (int q = 0; q < 4; ++q) {    // the four bins
            // d = (x_hi - c) + x_lo: the first sum is exact near the bin centre, where the weight is steep
            f32x2 w0, w1, w2, w3;
            weight4<METHOD>(add2(xx[0].x, negc[q]), add2(xx[0].y, negc[q]), xl[0].x, xl[0].y, wa2, wb2, w0, w1);
            weight4<METHOD>(add2(xx[1].x, negc[q]), add2(xx[1].y, negc[q]), xl[1].x, xl[1].y, wa2, wb2, w2, w3);
            if (side == 0) { w0 = mul2(w0, iw[0].x); w1 = mul2(w1, iw[0].y); w2 = mul2(w2, iw[1].x); w3 = mul2(w3, iw[1].y); }
            uint4 hi, lo;
            split_f16x2(w0, mone2, hi.x, lo.x);
            split_f16x2(w1, mone2, hi.y, lo.y);
            split_f16x2(w2, mone2, hi.z, lo.z);
            split_f16x2(w3, mone2, hi.w, lo.w);
            *reinterpret_cast<uint4*>(tile + q * (8 * 128)) = hi;
            *reinterpret_cast<uint4*>(tile + TILE_BYTES + q * (8 * 128)) = lo;
          }
          fence_proxy_async_smem();
          mbar_arrive_warp(&S.ab_full[stage]);

          // A chain that ended with the PREVIOUS stage is drained now, after this stage has been generated: when the
          // MMA warp gets the accumulators back it finds the first stage of the next chain ready instead of waiting
          // for the producers to refill the pipeline.
          if (pend_dst != nullptr) { drain(pend_dst, pend_first); pend_dst = nullptr; }
          if (((kb + 1) % CHAIN_KB == 0) || (kb + 1 == nkb)) {
            pend_dst = item_out + (int64_t)c * (BINS * BINS);
            pend_first = kb < CHAIN_KB;
          }
        }
      }
    }
    if (pend_dst != nullptr) drain(pend_dst, pend_first);
    if (tid == 0) bulk_wait_group<0>();
  } else if (warp == MMA_WARP) {
    // ===================== MMA issue: the whole warp runs the (uniform) loop, one elected lane issues ====
    constexpr uint32_t IDESC = idesc_f16(128, 256);
    const uint64_t desc0 = smem_desc_kmajor_noswizzle(smem_u32(&S.ab[0][0]), KCOL_BYTES, 128);
    const uint32_t dlo0 = (uint32_t)desc0, dhi = (uint32_t)(desc0 >> 32);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    uint32_t stage = 0, phase = 0, chain_par = 0;
    for (int64_t w = first; w < p.items; w += step) {
      const ItemRange ir = item_range(p, w);
      const uint32_t nkb = (ir.px1 - ir.px0 + KB - 1) / KB;
      for (int c = 0; c < 3; ++c) {
        for (uint32_t kb0 = 0; kb0 < nkb; kb0 += CHAIN_KB) {
          const uint32_t n_this = min((uint32_t)CHAIN_KB, nkb - kb0);
          for (uint32_t k = 0; k < n_this; ++k) {
            mbar_wait(&S.ab_full[stage], phase);
            tc_fence_after_sync();
            const uint32_t dstage = dlo0 + stage * (STAGE_BYTES >> 4);
#pragma unroll
            for (int ks = 0; ks < KB / 16; ++ks) {
              const uint32_t b_hi = dstage + ((2 * TILE_BYTES + ks * 2 * KCOL_BYTES) >> 4), b_lo = b_hi + (TILE_BYTES >> 4);
              const uint32_t acc0 = (k == 0 && ks == 0) ? 0u : 1u;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                // rows 128 h .. of the A tiles: 16 row groups x 128 B further
                const uint32_t a_hi = dstage + ((ks * 2 * KCOL_BYTES + h * 16 * 128) >> 4), a_lo = a_hi + (TILE_BYTES >> 4);
                if (elect_one_sync()) {
                  mma_f16_ss2(tm + h * 256, a_hi, b_hi, dhi, IDESC, acc0);
                  mma_f16_ss2(tm + h * 256, a_hi, b_lo, dhi, IDESC, 1u);
                  mma_f16_ss2(tm + h * 256, a_lo, b_hi, dhi, IDESC, 1u);
                }
              }
            }
            if (elect_one_sync()) mma_commit(&S.ab_empty[stage]);
            if (++stage == NS) { stage = 0; phase ^= 1; }
          }
          if (elect_one_sync()) mma_commit(&S.d_full);
          mbar_wait(&S.d_empty, chain_par);
          tc_fence_after_sync();
          chain_par ^= 1;
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem, TMEM_COLS);
}

// out[b, i, j, c] = (sum over the slices of partial[b][s][c][j][i]) / D_b,  D_b = sum of all of it
// (histogram.py:75-79).  One CTA per image: pass 1 adds the slices in order into slice 0 and reduces D, pass 2
// transposes 32 x 32 (i, j) tiles through shared memory so that both sides are coalesced.
__global__ void __launch_bounds__(256) hist256_finalize_kernel(float* __restrict__ partial, int splits,
                                                               const int* __restrict__ nunique, float inv_scale,
                                                               float iy_scale, float* __restrict__ hist,
                                                               float* __restrict__ denom) {
  constexpr int CH_STRIDE = 32 * 33 + 11;  // channel planes land on different banks
  __shared__ double scratch[32];
  __shared__ float tile[3 * CH_STRIDE];
  const int64_t b = blockIdx.x;
  float* mine = partial + b * splits * (int64_t)HIST_ELEMS;
  if (nunique != nullptr && nunique[b] >= 0) splits = 1;  // de-duplicated image: only slice 0 was written
  // a dense image inside a de-duplicated batch was contracted without the intensity scale (inv_scale contains 1 / iy_scale)
  if (nunique != nullptr && nunique[b] < 0) inv_scale *= iy_scale;
  double acc = 0.0;
  for (int e = threadIdx.x * 4; e < HIST_ELEMS; e += 256 * 4) {
    float4 v = *reinterpret_cast<const float4*>(mine + e);
    for (int s = 1; s < splits; ++s) {
      const float4 q = *reinterpret_cast<const float4*>(mine + (int64_t)s * HIST_ELEMS + e);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    if (splits > 1) *reinterpret_cast<float4*>(mine + e) = v;
    acc += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
  }
  const float d = (float)block_sum(acc, scratch);
  if (threadIdx.x == 0) denom[b] = d * inv_scale;
  const float inv_d = 1.0f / d;
  __syncthreads();  // pass 1's writes to slice 0 are visible to the whole block
  float* out = hist + b * (int64_t)HIST_ELEMS;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int t = 0; t < (BINS / 32) * (BINS / 32); ++t) {
    const int i0 = (t / (BINS / 32)) * 32, j0 = (t % (BINS / 32)) * 32;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int jj = ty; jj < 32; jj += 8)
        tile[c * CH_STRIDE + jj * 33 + tx] = mine[c * (BINS * BINS) + (j0 + jj) * BINS + i0 + tx];
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * 96; idx += 256) {
      const int ii = idx / 96, r = idx - ii * 96, jj = r / 3, c = r - jj * 3;
      out[(int64_t)(i0 + ii) * (BINS * 3) + j0 * 3 + r] = tile[c * CH_STRIDE + jj * 33 + ii] * inv_d;
    }
    __syncthreads();
  }
}

struct Plan { int splits; int64_t px_per_split; };

// Pixel slices per image: whole images while they fill the SMs (>= 95 % of the last wave), else the smallest
// number of slices (<= 16, each a multiple of the 1024-pixel chain) that does.
static Plan plan(int64_t batch, int64_t npix) {
  Plan pl{1, ceil_div(npix, SLOT_PX) * SLOT_PX};
  if (batch == 0) return pl;
  const int64_t sms = cached_sm_count();
  const int64_t chain_px = (int64_t)CHAIN_KB * KB;
  const int64_t max_s = npix / chain_px < 1 ? 1 : (npix / chain_px > 16 ? 16 : npix / chain_px);
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= max_s; ++s) {
    const int64_t items = batch * s;
    const double eff = (double)items / (double)(ceil_div(items, sms) * sms);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    if (eff >= 0.95) break;
  }
  int64_t pps = ceil_div(ceil_div(npix, best), chain_px) * chain_px;
  pl.splits = (int)ceil_div(npix, pps);  // every slice is non-empty
  pl.px_per_split = pps;
  return pl;
}

}  // namespace fwd256

void tc_fwd256_plan(int64_t batch, int64_t npix, int* slices_per_image, int64_t* px_per_slice) {
  const fwd256::Plan pl = fwd256::plan(batch, npix);
  *slices_per_image = pl.splits;
  *px_per_slice = pl.px_per_split;
}

size_t tc_fwd256_workspace_bytes(int64_t batch, int64_t npix) {
  const fwd256::Plan pl = fwd256::plan(batch, npix);
  return align_up((size_t)batch * pl.splits * fwd256::HIST_ELEMS * sizeof(float), 256);
}

int tc_fwd256_forward(const float* image, int64_t batch, int64_t npix, int channels, const float* dom, int method,
                      float sigma_sqr, float eps, const float4* ulist, const int* nunique, int dedup_max,
                      float iy_scale, float* hist, float* denom, void* workspace, cudaStream_t st) {
  using namespace fwd256;
  PH_CHECK_ARG(npix < (1ll << 31), "too many pixels per image");
  if (batch == 0) return PH_OK;
  // one plan for dense and de-duplicated batches: whether an image de-duplicates is known on the device only
  // (more than 512 colours -> dense), so a de-duplicated image simply leaves its slices > 0 empty
  const Plan pl = plan(batch, npix);
  Params p{};
  p.image = image;
  p.dom = dom;
  p.partial = static_cast<float*>(workspace);
  p.ulist = ulist;
  p.nunique = nunique;
  p.dedup_max = dedup_max;
  p.npix = npix;
  p.channels = channels;
  p.splits = pl.splits;
  p.px_per_split = pl.px_per_split;
  p.items = batch * pl.splits;
  p.eps = eps;
  const tcgen::WeightScales wsc = tcgen::weight_scales(method, sigma_sqr);
  p.wa = wsc.wa;
  p.wb = wsc.wb;
  p.coord_scale = wsc.coord_scale;
  p.iy_scale = iy_scale;
  p.status = async_status_word();
  PH_CHECK_ARG(p.status != nullptr, "no mapped status word (cudaHostAlloc failed)");
  const float inv_scale = (float)(1.0 / (wsc.weight_scale * wsc.weight_scale * (double)iy_scale));
  void (*kern)(Params) = method == PH_METHOD_INVERSE_QUADRATIC ? hist_fwd256_tc_kernel<PH_METHOD_INVERSE_QUADRATIC>
                                                               : hist_fwd256_tc_kernel<PH_METHOD_RBF>;
  const size_t smem = sizeof(Smem);
  PH_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = cached_sm_count();
  if (grid > p.items) grid = (int)p.items;
  kern<<<grid, THREADS, smem, st>>>(p);
  PH_LAUNCH_OK("hist_fwd256_tc_kernel");
  hist256_finalize_kernel<<<(unsigned)batch, 256, 0, st>>>(p.partial, pl.splits, nunique, inv_scale, iy_scale, hist, denom);
  PH_LAUNCH_OK("hist256_finalize_kernel");
  return PH_OK;
}

}  // namespace ph
