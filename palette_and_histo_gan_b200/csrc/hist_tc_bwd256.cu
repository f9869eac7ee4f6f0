// RGB-uv histogram backward with 256 bins (cfgE), tensor-core engine: the regime where the two products
//     P[n,i]  = sum_j Kv[n,j]  G^[i,j]        P'[n,i] = sum_j dKv[n,j] G^[i,j]          
// are bound by the tensor pipe.  hist_tc_bwd.cu covers 256 bins as 16 blocks of 64 x 64, regenerating the 64
// A-operand weights and the 64 epilogue weights of every pixel for every block; here one CTA multiplies a
// 128-pixel tile by the WHOLE 256 x 256 G^ of a channel: M = 128 pixels (one TMEM lane per pixel), N = 256 bins
// i, K = 256 bins j in 16 steps.  The accumulators P and P' are 256 columns each = all 512 columns of tensor
// memory, so the A operand [Kv_hi | Kv_lo | dKv_hi | dKv_lo] goes through shared memory (16 KB per K-step, 4
// stages) and G^ — 768 KB per image as fp16 hi + lo tiles, far more than shared memory — is streamed from
// L2 one K-step at a time (16 KB bulk copies into a 6-stage ring; the same 256 KB per channel serve every tile of
// the image).  Per K-step: P += Kv_hi.G_hi + Kv_hi.G_lo + Kv_lo.G_hi, the same three products for P'
// (tcgen05.mma kind::f16, SS, M128 N256 K16: 6 x 128 cycles).
//
//   warps 0-15   pixel p = tid % 128, cq = tid / 128: generate the A rows of pixel p for the K-steps of parity
//                cq / 2, bins j = 16 k + 8 (cq % 2) .. + 7 (whole 16-byte core-matrix rows: conflict-free stores);
//                epilogue of the round: bins i in [64 cq, 64 cq + 64) of the pixel's own TMEM lane, dot products
//                with Ku, dKu generated on the fly; the four partial sums of a pixel meet in shared memory
//   warp 16      one thread issues the MMAs
//   warps 17-18  pixel terms of the next tiles (float64 log-chroma hi + lo, intensity) -> smem ring
//   warp 19      one lane streams the G^ tiles (cp.async.bulk -> mbarrier)
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "hist_internal.cuh"
#include "hist_tc_gen.cuh"
#include "tc_ptx.cuh"

namespace ph {

using namespace tc;

namespace bwd256 {

using tcgen::named_bar_sync;
using tcgen::weight_pair2;

constexpr int BINS = 256;
constexpr int TILE = 128;
constexpr int KSTEPS = BINS / 16;               // 16 K-steps of 16 bins j
constexpr int GEN_WARPS = 16;
constexpr int MMA_WARP = 16;
constexpr int PX_WARP0 = 17;
constexpr int PXW = 2;                          // a tile lasts ~50 000 cycles here: two warps of pixel terms are plenty
constexpr int COPY_WARP = PX_WARP0 + PXW;       // 19
constexpr int THREADS = (COPY_WARP + 1) * 32;   // 640 = 5 warps per scheduler: 96 registers per thread
constexpr int PR = 4;                           // pixel-term ring slots (tiles)
constexpr int NSA = 4, NSB = 6;                 // A / B operand stages
constexpr int TMEM_COLS = 512;
constexpr int COL_P = 0, COL_PP = 256;
// A stage: [part: Kv_hi, Kv_lo, dKv_hi, dKv_lo][kcol = j / 8 (2)][pixel / 8 (16)][pixel % 8][j % 8]  fp16, K-major
constexpr int A_KCOL_BYTES = 16 * 128;          // 2048 (LBO)
constexpr int A_PART_BYTES = 2 * A_KCOL_BYTES;  // 4096
constexpr int A_STAGE_BYTES = 4 * A_PART_BYTES; // 16384
// B stage = one K-step of G^ of one channel: [hi, lo][kcol (2)][i / 8 (32)][i % 8][j % 8]
constexpr int B_KCOL_BYTES = 32 * 128;          // 4096 (LBO)
constexpr int B_PART_BYTES = 2 * B_KCOL_BYTES;  // 8192
constexpr int B_STAGE_BYTES = 2 * B_PART_BYTES; // 16384
constexpr int64_t G_CH_BYTES = (int64_t)KSTEPS * B_STAGE_BYTES;  // 262144
constexpr int64_t G_IMG_BYTES = 3 * G_CH_BYTES;                  // 786432

struct PxTile {
  float d_hi[3][TILE], d_lo[3][TILE];  // log-chroma differences rg, rb, gb as hi + lo
  float iy[TILE];
  float x[3][TILE];
};

struct Smem {
  alignas(128) unsigned char a[NSA][A_STAGE_BYTES];  // 64 KB
  alignas(128) unsigned char b[NSB][B_STAGE_BYTES];  // 96 KB
  PxTile px[PR];                                     // 20 KB
  float dom[BINS];
  float4 part[2][3][TILE];                           // partial sums of cq = 1..3, double-buffered by tile parity
  alignas(8) uint64_t px_full[PR], px_empty[PR], a_full[NSA], a_empty[NSA], b_full[NSB], b_empty[NSB], d_full, d_empty;
  uint32_t tmem_base;
};

struct Params {
  const float* image;
  const float* dom;          // 256 bin centres
  const unsigned char* g16;  // (B, G_IMG_BYTES): scaled G^ as fp16 hi | lo operand tiles, [c][kstep][hi, lo]...
  const float* gscale;       // (B): 1 / (power-of-two scale of the image's G^)
  float* grad;               // (B, npix, channels)
  int64_t npix;
  int64_t batch;
  int channels;
  int bsplit;          // work items per image (tile ranges)
  int tiles_per_item;
  float eps;
  float wa, wb, wc, coord_scale, inv_sk2, inv_sk_sdk;  // hist_tc_gen.cuh: bwd_scales
};

template <int METHOD>
__global__ void __launch_bounds__(THREADS, 1) hist_bwd256_tc_kernel(Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < PR; ++i) { mbar_init(&S.px_full[i], 1); mbar_init(&S.px_empty[i], GEN_WARPS); }
    for (int i = 0; i < NSA; ++i) { mbar_init(&S.a_full[i], GEN_WARPS / 2); mbar_init(&S.a_empty[i], 1); }
    for (int i = 0; i < NSB; ++i) { mbar_init(&S.b_full[i], 1); mbar_init(&S.b_empty[i], 1); }
    mbar_init(&S.d_full, 1);
    mbar_init(&S.d_empty, GEN_WARPS);
    fence_mbar_init();
  }
  if (tid < BINS) S.dom[tid] = p.dom[tid] * p.coord_scale;
  if (warp == MMA_WARP) tmem_alloc(&S.tmem_base, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = S.tmem_base;

  const int64_t tiles = (p.npix + TILE - 1) / TILE;
  const int64_t first = blockIdx.x, step = gridDim.x;
  const int64_t items = p.batch * p.bsplit;  // item = (image, contiguous range of 128-pixel tiles)

  if (warp == COPY_WARP) {
    // ===================== G^ stream: one 16 KB K-step tile per bulk copy =====================
    if (lane == 0) {
      uint32_t sb = 0, ph = 0;
      for (int64_t w = first; w < items; w += step) {
        const int64_t b = w / p.bsplit;
        const int64_t t_begin = (w % p.bsplit) * p.tiles_per_item;
        const int64_t t_end = min(t_begin + p.tiles_per_item, tiles);
        const unsigned char* gimg = p.g16 + b * G_IMG_BYTES;
        for (int64_t t = t_begin; t < t_end; ++t) {
          for (int ck = 0; ck < 3 * KSTEPS; ++ck) {  // (channel, K-step) in the order of the layout
            mbar_wait(&S.b_empty[sb], ph ^ 1);
            mbar_arrive_expect_tx(&S.b_full[sb], B_STAGE_BYTES);
            bulk_copy_g2s(&S.b[sb][0], gimg + (int64_t)ck * B_STAGE_BYTES, B_STAGE_BYTES, &S.b_full[sb]);
            if (++sb == NSB) { sb = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp >= PX_WARP0) {
    // ===================== pixel terms: whole tiles round-robin over the pixel warps =====================
    const int me = warp - PX_WARP0;
    uint32_t tt = 0;
    for (int64_t w = first; w < items; w += step) {
      const int64_t b = w / p.bsplit;
      const int64_t t_begin = (w % p.bsplit) * p.tiles_per_item;
      const int64_t t_end = min(t_begin + p.tiles_per_item, tiles);
      for (int64_t t = t_begin; t < t_end; ++t, ++tt) {
        if ((int)(tt % PXW) != me) continue;
        const int slot = tt % PR;
        PixelTerms pt[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int64_t px = t * TILE + k * 32 + lane;
          float r = 0.f, g = 0.f, bl = 0.f;
          if (px < p.npix) {
            const float* src = p.image + (b * p.npix + px) * p.channels;
            if (p.channels == 4) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(src));
              r = q.x; g = q.y; bl = q.z;
            } else {
              r = __ldg(src); g = __ldg(src + 1); bl = __ldg(src + 2);
            }
          }
          pt[k] = pixel_terms(r, g, bl, p.eps);
        }
        mbar_wait_relaxed(&S.px_empty[slot], ((tt / PR) & 1) ^ 1, 1000);
        PxTile& o = S.px[slot];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int pl = k * 32 + lane;
          const float cs = p.coord_scale;  // power of two: exact for both parts
          o.d_hi[0][pl] = pt[k].d_rg * cs; o.d_lo[0][pl] = pt[k].l_rg * cs;
          o.d_hi[1][pl] = pt[k].d_rb * cs; o.d_lo[1][pl] = pt[k].l_rb * cs;
          o.d_hi[2][pl] = pt[k].d_gb * cs; o.d_lo[2][pl] = pt[k].l_gb * cs;
          o.iy[pl] = pt[k].iy;
          o.x[0][pl] = pt[k].x0; o.x[1][pl] = pt[k].x1; o.x[2][pl] = pt[k].x2;
        }
        mbar_arrive_warp(&S.px_full[slot]);
      }
    }
  } else if (warp < GEN_WARPS) {
    // ===================== A-operand generation + epilogue =====================
    const int quad = warp & 3;   // TMEM sub-partition (== warp % 4)
    const int cq = warp >> 2;    // epilogue bins i in [64 cq, 64 cq + 64)
    const int gpar = cq >> 1;    // generates the K-steps of this parity ...
    const int kcol = cq & 1;     // ... bins j = 16 k + 8 kcol .. + 7
    const int pl = quad * 32 + lane;
    const uint32_t tbase = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t a_row = (uint32_t)(kcol * A_KCOL_BYTES + pl * 16);  // [pixel / 8][pixel % 8] rows are 16 B apart
    const f32x2 wa2 = pack2(p.wa, p.wa), wb2 = pack2(p.wb, p.wb), wc2 = pack2(p.wc, p.wc), mone2 = pack2(-1.0f, -1.0f);
    uint32_t nround = 0, tt = 0;  // (tile, channel) rounds = d_full phases; tiles = pixel-ring phases
    // A rows [Kv_hi | Kv_lo | dKv_hi | dKv_lo] of this thread's pixel for K-step k (bins j = 16 k + 8 kcol .. + 7) of
    // round `round`, from the pixel's v coordinate (hi + lo)
    auto gen_kstep = [&](int k, uint32_t round, f32x2 vh2, f32x2 vl2) {
      const uint32_t ks = round * KSTEPS + k;  // global K-step: stage ks % NSA, phase ks / NSA
      const uint32_t sa = ks % NSA;
      uint4 kh, kl, dh4, dl4;
      {
        const float* cj = &S.dom[k * 16 + kcol * 8];
        const ulonglong2 c0 = *reinterpret_cast<const ulonglong2*>(cj);
        const ulonglong2 c1 = *reinterpret_cast<const ulonglong2*>(cj + 4);
        f32x2 kk, dk;
        weight_pair2<METHOD>(vh2, vl2, c0.x, wa2, wb2, wc2, mone2, kk, dk);
        split_f16x2(kk, mone2, kh.x, kl.x); split_f16x2(dk, mone2, dh4.x, dl4.x);
        weight_pair2<METHOD>(vh2, vl2, c0.y, wa2, wb2, wc2, mone2, kk, dk);
        split_f16x2(kk, mone2, kh.y, kl.y); split_f16x2(dk, mone2, dh4.y, dl4.y);
        weight_pair2<METHOD>(vh2, vl2, c1.x, wa2, wb2, wc2, mone2, kk, dk);
        split_f16x2(kk, mone2, kh.z, kl.z); split_f16x2(dk, mone2, dh4.z, dl4.z);
        weight_pair2<METHOD>(vh2, vl2, c1.y, wa2, wb2, wc2, mone2, kk, dk);
        split_f16x2(kk, mone2, kh.w, kl.w); split_f16x2(dk, mone2, dh4.w, dl4.w);
      }
      mbar_wait(&S.a_empty[sa], ((ks / NSA) & 1) ^ 1);  // the MMAs that read this stage are done
      unsigned char* row = &S.a[sa][a_row];
      *reinterpret_cast<uint4*>(row) = kh;
      *reinterpret_cast<uint4*>(row + A_PART_BYTES) = kl;
      *reinterpret_cast<uint4*>(row + 2 * A_PART_BYTES) = dh4;
      *reinterpret_cast<uint4*>(row + 3 * A_PART_BYTES) = dl4;
      fence_proxy_async_smem();
      mbar_arrive_warp(&S.a_full[sa]);
    };
    for (int64_t w = first; w < items; w += step) {
      const int64_t b = w / p.bsplit;
      const int64_t t_begin = (w % p.bsplit) * p.tiles_per_item;
      const int64_t t_end = min(t_begin + p.tiles_per_item, tiles);
      const float inv_g = __ldg(p.gscale + b);
      const float c_iy = inv_g * p.inv_sk2, c_uv = inv_g * p.inv_sk_sdk;
      for (int64_t t = t_begin; t < t_end; ++t, ++tt) {
        const int slot = tt % PR;
        mbar_wait(&S.px_full[slot], (tt / PR) & 1);
        const PxTile& in = S.px[slot];
        const float dh[3] = {in.d_hi[0][pl], in.d_hi[1][pl], in.d_hi[2][pl]};
        const float dl[3] = {in.d_lo[0][pl], in.d_lo[1][pl], in.d_lo[2][pl]};
        const float iy = in.iy[pl];
        const float x0 = in.x[0][pl], x1 = in.x[1][pl], x2 = in.x[2][pl];
        mbar_arrive_warp(&S.px_empty[slot]);
        float g_rg = 0.f, g_rb = 0.f, g_gb = 0.f, g_iy = 0.f;
#pragma unroll 1
        for (int c = 0; c < 3; ++c, ++nround) {
          // (u, v) of channel c: R:(rg, rb)  G:(-rg, gb)  B:(-rb, -gb)   (histogram.py:72-74)
          const float u_hi = c == 0 ? dh[0] : (c == 1 ? -dh[0] : -dh[1]);
          const float u_lo = c == 0 ? dl[0] : (c == 1 ? -dl[0] : -dl[1]);
          const float v_hi = c == 0 ? dh[1] : (c == 1 ? dh[2] : -dh[2]);
          const float v_lo = c == 0 ? dl[1] : (c == 1 ? dl[2] : -dl[2]);
          const f32x2 uh2 = pack2(u_hi, u_hi), ul2 = pack2(u_lo, u_lo), vh2 = pack2(v_hi, v_hi), vl2 = pack2(v_lo, v_lo);
          // ---- A rows of this pixel for the K-steps of this thread's parity ----
          // (the first one was generated before the previous round's epilogue when that round was of the same tile)
#pragma unroll 1
          for (int k = gpar + (c > 0 ? 2 : 0); k < KSTEPS; k += 2) gen_kstep(k, nround, vh2, vl2);
          if (c < 2) {
            // first K-step of the NEXT channel's round, before this round's epilogue: when the MMA warp gets the
            // accumulators back it finds K-steps 0 and 1 ready instead of waiting for the producers to restart
            const float nv_hi = c == 0 ? dh[2] : -dh[2], nv_lo = c == 0 ? dl[2] : -dl[2];
            gen_kstep(gpar, nround + 1, pack2(nv_hi, nv_hi), pack2(nv_lo, nv_lo));
          }
          // ---- epilogue of the round: dot products of P, P' (own lane) with the u-side weights of 64 bins ----
          // the weights of the first 16 bins are evaluated before waiting for the accumulators
          f32x2 ku0[8], dku0[8];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const ulonglong2 ci = *reinterpret_cast<const ulonglong2*>(&S.dom[cq * 64 + q4 * 4]);
            weight_pair2<METHOD>(uh2, ul2, ci.x, wa2, wb2, wc2, mone2, ku0[q4 * 2], dku0[q4 * 2]);
            weight_pair2<METHOD>(uh2, ul2, ci.y, wa2, wb2, wc2, mone2, ku0[q4 * 2 + 1], dku0[q4 * 2 + 1]);
          }
          mbar_wait(&S.d_full, nround & 1);
          tc_fence_after_sync();
          f32x2 s_iy2 = 0ull, s_u2 = 0ull, s_v2 = 0ull;  // two-lane partial sums (bit pattern 0 = +0.0f, +0.0f)
          {
            uint32_t pv[16], qv[16];
            tmem_ld16(tbase + COL_P + cq * 64, pv);
            tmem_ld16(tbase + COL_PP + cq * 64, qv);
            tmem_ld_wait();
#pragma unroll
            for (int o2 = 0; o2 < 8; ++o2) {
              const f32x2 pp = (unsigned long long)pv[2 * o2] | ((unsigned long long)pv[2 * o2 + 1] << 32);
              const f32x2 qq = (unsigned long long)qv[2 * o2] | ((unsigned long long)qv[2 * o2 + 1] << 32);
              s_iy2 = fma2(ku0[o2], pp, s_iy2);
              s_u2 = fma2(dku0[o2], pp, s_u2);
              s_v2 = fma2(ku0[o2], qq, s_v2);
            }
          }
#pragma unroll 1
          for (int ch = 1; ch < 4; ++ch) {
            uint32_t pv[16], qv[16];
            tmem_ld16(tbase + COL_P + cq * 64 + ch * 16, pv);
            tmem_ld16(tbase + COL_PP + cq * 64 + ch * 16, qv);
            tmem_ld_wait();
            if (ch == 3) {
              // both accumulators are in registers: the next round's MMAs may overwrite them
              tc_fence_before_sync();
              mbar_arrive_warp(&S.d_empty);
            }
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const ulonglong2 ci = *reinterpret_cast<const ulonglong2*>(&S.dom[cq * 64 + ch * 16 + q4 * 4]);
              const f32x2 cc[2] = {ci.x, ci.y};
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                f32x2 k, dk;
                weight_pair2<METHOD>(uh2, ul2, cc[e], wa2, wb2, wc2, mone2, k, dk);
                const int o = q4 * 4 + e * 2;
                const f32x2 pp = (unsigned long long)pv[o] | ((unsigned long long)pv[o + 1] << 32);
                const f32x2 qq = (unsigned long long)qv[o] | ((unsigned long long)qv[o + 1] << 32);
                s_iy2 = fma2(k, pp, s_iy2);
                s_u2 = fma2(dk, pp, s_u2);
                s_v2 = fma2(k, qq, s_v2);
              }
            }
          }
          const float s_iy = (lo_of(s_iy2) + hi_of(s_iy2)) * c_iy;
          const float a = (lo_of(s_u2) + hi_of(s_u2)) * (c_uv * iy), bq = (lo_of(s_v2) + hi_of(s_v2)) * (c_uv * iy);
          g_iy += s_iy;
          if (c == 0) { g_rg += a; g_rb += bq; }
          else if (c == 1) { g_rg -= a; g_gb += bq; }
          else { g_rb -= a; g_gb -= bq; }
        }
        // ---- combine the four bin-quarters of the pixel and write its gradient ----
        const int par = (int)(tt & 1);
        if (cq > 0) S.part[par][cq - 1][pl] = make_float4(g_rg, g_rb, g_gb, g_iy);
        named_bar_sync(1, GEN_WARPS * 32);
        if (cq == 0) {
#pragma unroll
          for (int o3 = 0; o3 < 3; ++o3) {
            const float4 o = S.part[par][o3][pl];
            g_rg += o.x; g_rb += o.y; g_gb += o.z; g_iy += o.w;
          }
          const int64_t px = t * TILE + pl;
          if (px < p.npix) {
            const float dl_r = g_rg + g_rb, dl_g = -g_rg + g_gb, dl_b = -g_rb - g_gb;
            const float wy = g_iy / iy;
            const float gx0 = dl_r / (x0 + p.eps) + wy * x0;
            const float gx1 = dl_g / (x1 + p.eps) + wy * x1;
            const float gx2 = dl_b / (x2 + p.eps) + wy * x2;
            float* dst = p.grad + (b * p.npix + px) * p.channels;
            if (p.channels == 4) {
              *reinterpret_cast<float4*>(dst) = make_float4(0.5f * gx0, 0.5f * gx1, 0.5f * gx2, 0.f);
            } else {
              dst[0] = 0.5f * gx0; dst[1] = 0.5f * gx1; dst[2] = 0.5f * gx2;
            }
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issue: uniform loop, one elected lane issues =====================
    constexpr uint32_t IDESC = idesc_f16(128, 256);
    const uint64_t adesc0 = smem_desc_kmajor_noswizzle(smem_u32(&S.a[0][0]), A_KCOL_BYTES, 128);
    const uint64_t bdesc0 = smem_desc_kmajor_noswizzle(smem_u32(&S.b[0][0]), B_KCOL_BYTES, 128);
    const uint32_t alo0 = (uint32_t)adesc0, blo0 = (uint32_t)bdesc0, dhi = (uint32_t)(adesc0 >> 32);  // same SBO
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);  // provably uniform copy
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0, dpar = 0;
    for (int64_t w = first; w < items; w += step) {
      const int64_t t_begin = (w % p.bsplit) * p.tiles_per_item;
      const int64_t t_end = min(t_begin + p.tiles_per_item, tiles);
      for (int64_t t = t_begin; t < t_end; ++t) {
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
          mbar_wait(&S.d_empty, dpar ^ 1);  // the epilogue of the previous round has read P, P'
          tc_fence_after_sync();
#pragma unroll 1
          for (int k = 0; k < KSTEPS; ++k) {
            mbar_wait(&S.a_full[sa], pa);
            mbar_wait(&S.b_full[sb], pb);
            tc_fence_after_sync();
            const uint32_t a0 = alo0 + sa * (A_STAGE_BYTES >> 4);
            const uint32_t b_hi = blo0 + sb * (B_STAGE_BYTES >> 4), b_lo = b_hi + (B_PART_BYTES >> 4);
            const uint32_t acc0 = k == 0 ? 0u : 1u;
            if (elect_one_sync()) {
              mma_f16_ss2(tm + COL_P, a0, b_hi, dhi, IDESC, acc0);
              mma_f16_ss2(tm + COL_P, a0, b_lo, dhi, IDESC, 1u);
              mma_f16_ss2(tm + COL_P, a0 + (A_PART_BYTES >> 4), b_hi, dhi, IDESC, 1u);
              mma_f16_ss2(tm + COL_PP, a0 + (2 * A_PART_BYTES >> 4), b_hi, dhi, IDESC, acc0);
              mma_f16_ss2(tm + COL_PP, a0 + (2 * A_PART_BYTES >> 4), b_lo, dhi, IDESC, 1u);
              mma_f16_ss2(tm + COL_PP, a0 + (3 * A_PART_BYTES >> 4), b_hi, dhi, IDESC, 1u);
            }
            if (elect_one_sync()) { mma_commit(&S.a_empty[sa]); mma_commit(&S.b_empty[sb]); }
            if (++sa == NSA) { sa = 0; pa ^= 1; }
            if (++sb == NSB) { sb = 0; pb ^= 1; }
          }
          if (elect_one_sync()) mma_commit(&S.d_full);
          dpar ^= 1;
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem, TMEM_COLS);
}

// G^ (float, [c][i][j] over the 256 x 256 histogram, written by the generic prologue) -> per image the B-operand
// stream of the contraction kernel: [c][kstep = j / 16][hi, lo][kcol = (j / 8) % 2][i][j % 8] fp16, scaled by one
// image-wide power of two (max |G^| -> [2^13, 2^14)).  One thread per 16-byte operand row (8 bins j of one i).
__global__ void __launch_bounds__(256) ghat_to_f16_256_kernel(const float* __restrict__ ghat,
                                                              const float* __restrict__ gmax,
                                                              unsigned char* __restrict__ g16,
                                                              float* __restrict__ gscale) {
  const int64_t b = blockIdx.x;
  const float vmax = gmax[b];
  int ex = 0;
  if (vmax > 0.f && vmax < 3.0e38f) ex = 13 - ilogbf(vmax);
  ex = max(-100, min(100, ex));
  const float scale = ldexpf(1.0f, ex);
  const float* src = ghat + b * (int64_t)3 * BINS * BINS;
  uint4* dst = reinterpret_cast<uint4*>(g16 + b * G_IMG_BYTES);
  const f32x2 mone2 = pack2(-1.0f, -1.0f);
  // r = ((c * 16 + kstep) * 2 + kcol) * 256 + i
  for (int r = blockIdx.y * 256 + threadIdx.x; r < 3 * KSTEPS * 2 * BINS; r += gridDim.y * 256) {
    const int i = r & 255, kcol = (r >> 8) & 1, ck = r >> 9, c = ck / KSTEPS, kstep = ck - c * KSTEPS;
    const float* row = src + ((int64_t)c * BINS + i) * BINS + kstep * 16 + kcol * 8;
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(row)), v1 = __ldg(reinterpret_cast<const float4*>(row) + 1);
    uint4 hi, lo;
    split_f16x2(pack2(v0.x * scale, v0.y * scale), mone2, hi.x, lo.x);
    split_f16x2(pack2(v0.z * scale, v0.w * scale), mone2, hi.y, lo.y);
    split_f16x2(pack2(v1.x * scale, v1.y * scale), mone2, hi.z, lo.z);
    split_f16x2(pack2(v1.z * scale, v1.w * scale), mone2, hi.w, lo.w);
    const int o16 = ck * (B_STAGE_BYTES / 16) + kcol * (B_KCOL_BYTES / 16) + i;
    dst[o16] = hi;
    dst[o16 + B_PART_BYTES / 16] = lo;
  }
  if (threadIdx.x == 0 && blockIdx.y == 0) gscale[b] = ldexpf(1.0f, -ex);
}

__global__ void __launch_bounds__(256) ghat_absmax256_kernel(const float* __restrict__ ghat, float* __restrict__ gmax) {
  __shared__ float wmax[8];
  const float* g = ghat + blockIdx.x * (int64_t)(3 * BINS * BINS);
  float m = 0.f;
  for (int e = threadIdx.x * 4; e < 3 * BINS * BINS; e += 256 * 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g + e));
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) m = fmaxf(m, wmax[k]);
    gmax[blockIdx.x] = m;
  }
}

}  // namespace bwd256

// Items per image (tile ranges): the split that minimises (waves of items) x (item size) — with at least sms / 32
// items per image, so that the CTAs of a wave stream the G^ of at most ~32 images (24 MB) and find it in L2: at one
// image per CTA the 148 streams (114 MB) evict each other and every K-step comes from HBM (measured 1.6x slower).
void tc_bwd256_plan(int64_t batch, int64_t npix, int* items_per_image, int* tiles_per_item) {
  const int64_t sms = cached_sm_count();
  const int64_t tiles = ceil_div(npix, bwd256::TILE);
  int64_t smin = ceil_div(sms, 32);
  if (smin > tiles) smin = tiles;
  int best = (int)smin;
  double best_cost = 1e30;
  for (int64_t sidx = smin; sidx <= 32 && sidx <= tiles; ++sidx) {
    const int64_t tpi = ceil_div(tiles, sidx);
    if (sidx > 1 && tpi * (sidx - 1) >= tiles) continue;
    const double cost = (double)ceil_div(batch * sidx, sms) * ((double)tpi + 0.25);  // + pipeline fill / drain
    if (cost < best_cost * 0.97) { best_cost = cost; best = (int)sidx; }
  }
  int bsplit = best;
  const int tpi = (int)ceil_div(tiles, best);
  if ((int64_t)tpi * (bsplit - 1) >= tiles) bsplit = (int)ceil_div(tiles, tpi);  // no empty item
  *items_per_image = bsplit;
  *tiles_per_item = tpi;
}

// workspace: [fp16 operand stream of G^ (B x 768 KB)][gscale (B)][float G^ (B x 768 KB)][gmax (B)]
size_t tc_bwd256_workspace_bytes(int64_t batch) {
  return align_up((size_t)batch * bwd256::G_IMG_BYTES, 256) + 2 * align_up((size_t)batch * sizeof(float), 256) +
         align_up((size_t)batch * 3 * 256 * 256 * sizeof(float), 256);
}

int tc_bwd256_backward(const float* image, int64_t batch, int64_t npix, int channels, const float* dom, int method,
                       float sigma_sqr, float eps, const float* hist_pred, const float* denom, const float* grad_hist,
                       const float* hist_true, const double* ssum, int64_t global_batch, const float* loss_scale,
                       float* grad_image, void* workspace, cudaStream_t st) {
  using namespace bwd256;
  if (batch == 0) return PH_OK;
  char* ws = static_cast<char*>(workspace);
  unsigned char* g16 = reinterpret_cast<unsigned char*>(ws);
  ws += align_up((size_t)batch * G_IMG_BYTES, 256);
  float* gscale = reinterpret_cast<float*>(ws);
  ws += align_up((size_t)batch * sizeof(float), 256);
  float* ghat = reinterpret_cast<float*>(ws);
  ws += align_up((size_t)batch * 3 * BINS * BINS * sizeof(float), 256);
  float* gmax = reinterpret_cast<float*>(ws);
  int rc = launch_bwd_prep(hist_pred, denom, grad_hist, hist_true, ssum, global_batch, loss_scale, batch, BINS, 0, ghat, st);
  if (rc != PH_OK) return rc;
  ghat_absmax256_kernel<<<(unsigned)batch, 256, 0, st>>>(ghat, gmax);
  PH_LAUNCH_OK("ghat_absmax256_kernel");
  {
    // enough CTAs per image to fill the machine at small batches (24 576 rows of 16 bytes per image)
    int gy = (int)ceil_div(4 * (int64_t)cached_sm_count(), batch);
    gy = gy < 1 ? 1 : (gy > 96 ? 96 : gy);
    ghat_to_f16_256_kernel<<<dim3((unsigned)batch, (unsigned)gy), 256, 0, st>>>(ghat, gmax, g16, gscale);
    PH_LAUNCH_OK("ghat_to_f16_256_kernel");
  }
  Params p{};
  p.image = image;
  p.dom = dom;
  p.g16 = g16;
  p.gscale = gscale;
  p.grad = grad_image;
  p.npix = npix;
  p.batch = batch;
  p.channels = channels;
  p.eps = eps;
  const tcgen::BwdScales sc = tcgen::bwd_scales(method, sigma_sqr);
  p.coord_scale = sc.coord_scale;
  p.wa = sc.wa; p.wb = sc.wb; p.wc = sc.wc;
  p.inv_sk2 = sc.inv_sk2;
  p.inv_sk_sdk = sc.inv_sk_sdk;
  tc_bwd256_plan(batch, npix, &p.bsplit, &p.tiles_per_item);
  const size_t smem = sizeof(Smem);
  int grid = cached_sm_count();
  if (grid > batch * p.bsplit) grid = (int)(batch * p.bsplit);
  void (*kern)(Params) = method == PH_METHOD_INVERSE_QUADRATIC ? hist_bwd256_tc_kernel<PH_METHOD_INVERSE_QUADRATIC>
                                                                : hist_bwd256_tc_kernel<PH_METHOD_RBF>;
  PH_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, THREADS, smem, st>>>(p);
  PH_LAUNCH_OK("hist_bwd256_tc_kernel");
  return PH_OK;
}

}  // namespace ph
