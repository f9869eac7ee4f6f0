// Inline-PTX wrappers for the sm_100a pieces the tensor-core engine uses: mbarrier, tcgen05
// (alloc / mma / commit / ld / st / fences) and shared-memory matrix descriptors.
// Bit layouts follow the PTX ISA tables for tcgen05 (cross-checked against the descriptor structs in
// the vendored CUTLASS headers, cute/arch/mma_sm100_desc.hpp).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace ph {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Warp index the compiler can prove warp-uniform (shuffle from lane 0): code that branches on it keeps
// its operands in uniform registers, which is what tcgen05.mma / commit take.  With `threadIdx.x >> 5`
// ptxas wraps every MMA in an ELECT / R2UR.BROADCAST waterfall loop (~100 cycles per instruction).
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar))
               : "memory");
}
// One arrival per warp: the warp synchronises (ordering every lane's prior writes before lane 0's
// releasing arrive) and lane 0 arrives.  32 lanes arriving on the same mbarrier word serialise in the
// shared-memory atomic unit (~1 cycle per lane): with 16 producer warps and two barriers per stage that was
// ~1000 cycles of atomics per 32-pixel stage.
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Same with an explicit suspend-time hint: the thread sleeps in hardware until the phase completes or `ns`
// nanoseconds have passed.  Without the hint try_wait returns after a few tens of cycles, and the polling loops
// of the waiting warps (pixel and B-operand warps run far ahead of their consumers) took ~30 % of all issued
// instructions of the forward kernel (ncu source page), competing with the producer warps for issue slots.
__device__ __forceinline__ bool mbar_try_wait_for(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must abort the kernel (trap -> launch error), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  // every poll sleeps in hardware (>= tens of ns), so 2^27 polls are several seconds: counting polls instead
  // of reading the clock keeps the loop at a handful of instructions (the waiting warps share issue slots
  // with the producers)
  uint32_t polls = 0;
  while (!mbar_try_wait_for(bar, parity, 200000u)) {
    if (++polls > (1u << 27)) {
      printf("palhist: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}
// Wait of a warp that runs far ahead of its consumer (pixel warps, several ring slots of slack): poll with
// real sleeps in between, so that the warp leaves the issue slots to the others.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t sleep_ns) {
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(sleep_ns);
    if (++polls > (1u << 24)) {
      printf("palhist: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// ---- bulk asynchronous copy global -> shared (TMA, 1-D): completion counted in bytes on an mbarrier ----
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- bulk asynchronous copy / reduction shared -> global (TMA, 1-D), tracked by bulk async-groups ----
// Issued by ONE thread after the writers' fence.proxy.async + a barrier.  `add`: dst[i] += src[i] (fp32) performed at
// the L2, fire-and-forget for the SM: the threads never stall on the memory pipeline as they do with per-lane REDs.
__device__ __forceinline__ void bulk_store_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_f32_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's most recent groups may still be READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// at most N of this thread's most recent groups may still be incomplete (writes to global not yet performed)
template <int N>
__device__ __forceinline__ void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- proxies / fences ---------------------------------------------------------------------------
// generic-proxy smem writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- TMEM allocation (one full warp) ------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors --------------------------------------------------------------------------------
// K-major, no swizzle ("interleave"): core matrix = 8 rows x 16 B stored as 128 contiguous bytes;
// LBO = byte distance between the two core matrices along K, SBO = between 8-row groups along M/N.
__device__ __forceinline__ uint64_t smem_desc_kmajor_noswizzle(uint32_t smem_addr, uint32_t lbo_bytes,
                                                               uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  return d;         // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// kind::tf32 instruction descriptor: fp32 accumulate, tf32 A and B, both K-major, dense.  (The tf32 wrappers below
// are what the first version of the kernels used; tools/mma_bench.cu still times them against kind::f16.)
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n) {
  return (1u << 4)                    // c_format = F32
         | (2u << 7)                  // a_format = TF32
         | (2u << 10)                 // b_format = TF32
         | ((uint32_t)(n >> 3) << 17)  // n_dim
         | ((uint32_t)(m >> 4) << 24); // m_dim
}

// kind::f16 instruction descriptor: fp32 accumulate, fp16 A and B (format 0), both K-major, dense.
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- MMA: D[tmem] (+)= A[tmem] * B[smem], issued by ONE thread ---------------------------------
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the descriptor given as (low word, high word): the issuing warp keeps the constant high
// word and one base low word in uniform registers and adds compile-time offsets to the low word.
__device__ __forceinline__ void mma_tf32_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_desc_lo, uint32_t b_desc_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "r"(b_desc_lo), "r"(b_desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16, K = 16 per instruction.  A in TMEM: lane = row, each 32-bit column holds two consecutive K
// elements (low half = even k); B in shared memory: core matrix = 8 rows x 8 halfs (tools/f16_layout_test.cu).
__device__ __forceinline__ void mma_f16_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_desc_lo, uint32_t b_desc_hi,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "r"(b_desc_lo), "r"(b_desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 with both operands in shared memory; the two descriptors share their high word
__device__ __forceinline__ void mma_f16_ss2(uint32_t d_tmem, uint32_t a_desc_lo, uint32_t b_desc_lo, uint32_t desc_hi,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 ad, {%1, %3};\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %4, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_desc_lo), "r"(b_desc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM <-> registers: 32 lanes x 32 bit, N consecutive columns per thread ----------------------
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// tf32 split of an fp32 value: hi keeps the top 19 bits (exactly representable in tf32), lo = w - hi
// is exact in fp32; hi*hi + hi*lo + lo*hi (+ lo*lo) recovers the fp32 product to ~2^-21 relative.
__device__ __forceinline__ uint32_t tf32_hi(float w) { return __float_as_uint(w) & 0xFFFFE000u; }
__device__ __forceinline__ uint32_t tf32_lo(float w) {
  return __float_as_uint(w - __uint_as_float(__float_as_uint(w) & 0xFFFFE000u));
}
// ---- packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2 on sm_100): two fp32 lanes per instruction ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
  return (unsigned long long)__float_as_uint(a) | ((unsigned long long)__float_as_uint(b) << 32);
}
__device__ __forceinline__ float lo_of(f32x2 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi_of(f32x2 v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
constexpr unsigned long long TF32_MASK2 = 0xFFFFE000FFFFE000ull;

// ---- fp16 operand split: w (already scaled into fp16's range) = hi + lo, both fp16 ----------------------
// hi = fp16(w), lo = fp16(w - hi): hi.hi + hi.lo + lo.hi recovers the fp32 product to ~2^-21 like the tf32 split does, at
// twice the tensor-core rate (K = 16 per instruction) and half the operand bytes.
__device__ __forceinline__ uint32_t cvt_f16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
__device__ __forceinline__ void split_f16x2(f32x2 w, f32x2 /*mone2*/, uint32_t& hi, uint32_t& lo) {
  // hi = fp16(w) (round to nearest, packed conversion); lo = fp16(w - hi) with the mixed-precision FMA
  // (FHFMA: fp16 x fp16 + fp32, exact here): 4 instructions per pair of weights
  const float w0 = lo_of(w), w1 = hi_of(w);
  float l0, l1;
  asm("{\n\t.reg .f16 h0, h1, m1;\n\t"
      "cvt.rn.f16x2.f32 %0, %4, %3;\n\t"
      "mov.b32 {h0, h1}, %0;\n\t"
      "mov.b16 m1, 0xBC00;\n\t"  // -1.0
      "fma.rn.f32.f16 %1, h0, m1, %3;\n\t"
      "fma.rn.f32.f16 %2, h1, m1, %4;\n\t}"
      : "=&r"(hi), "=f"(l0), "=f"(l1)
      : "f"(w0), "f"(w1));
  lo = cvt_f16x2(l0, l1);
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

}  // namespace tc
}  // namespace ph
