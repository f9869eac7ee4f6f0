// Layout check for tcgen05.mma kind::f16 with the A operand in TMEM (packed half2 per 32-bit column)
// and B in shared memory as K-major no-swizzle core matrices (8 rows x 8 halfs).  Prints max |D - ref|.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../palette_and_histo_gan_b200/csrc/tc_ptx.cuh"
using namespace ph::tc;

__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ inline float a_val(int m, int k) { return (float)((m * 7 + k * 3) % 17 - 8) * 0.25f; }
__host__ __device__ inline float b_val(int n, int k) { return (float)((n * 5 + k * 11) % 13 - 6) * 0.5f; }

// K = 32 (two K16 steps), N = 64 or 128
__global__ void test(int N, float* out, long long* cyc) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tmem = tmem_base;
  // B: [kb = k/8 (4)][ng = n/8][n%8][k%8] halfs; kb stride = N/8*128 bytes
  const int kb_stride = N / 8 * 128;
  for (int e = tid; e < N * 32; e += blockDim.x) {
    const int n = e / 32, k = e % 32;
    __half* dst = reinterpret_cast<__half*>(smem + (k / 8) * kb_stride + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2);
    *dst = __float2half(b_val(n, k));
  }
  // A: lane m = tid, column c holds k = 2c (low half), 2c+1 (high half); 16 columns for K = 32
  {
    uint32_t w[16];
    for (int c = 0; c < 16; ++c) {
      const __half2 h = __floats2half2_rn(a_val(tid, 2 * c), a_val(tid, 2 * c + 1));
      w[c] = *reinterpret_cast<const uint32_t*>(&h);
    }
    tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 256, w);
    tmem_st_wait();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // f16 x f16 -> f32
    const uint32_t sb = smem_u32(smem);
    const long long t0 = clock64();
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t bdesc = smem_desc_kmajor_noswizzle(sb + ks * 2 * kb_stride, kb_stride, 128);
      mma_f16_ts(tmem, tmem + 256 + ks * 8, bdesc, idesc, ks);
    }
    mma_commit(&bar); mbar_wait(&bar, 0);
    // timing: 2048 more MMAs into the other columns
    const uint64_t bdesc = smem_desc_kmajor_noswizzle(sb, kb_stride, 128);
    const long long t1 = clock64();
    for (int i = 0; i < 2048; ++i) mma_f16_ts(tmem + 128, tmem + 256 + (i & 1) * 8, bdesc, idesc, 1);
    mma_commit(&bar); mbar_wait(&bar, 1);
    cyc[0] = clock64() - t1; cyc[1] = t1 - t0;
  }
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[tid * N + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}


// SS mode, both operands MN-major (row index contiguous), no swizzle:
//   element (r, k) at (r/8)*rb_stride + (k/8)*kb_stride + (k%8)*16 + (r%8)*2   (core matrix = 8 k x 8 rows = 128 B)
__global__ void test_ss_mn(int swap, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tmem = tmem_base;
  const int rb_stride = 512, kb_stride = 128;  // [row block (16)][k block (4)][k%8][row%8]
  unsigned char* A = smem; unsigned char* B = smem + 8192;
  for (int e = tid; e < 128 * 32; e += blockDim.x) {
    const int r = e / 32, k = e % 32;
    const int off = (r / 8) * rb_stride + (k / 8) * kb_stride + (k % 8) * 16 + (r % 8) * 2;
    *reinterpret_cast<__half*>(A + off) = __float2half(a_val(r, k));
    *reinterpret_cast<__half*>(B + off) = __float2half(b_val(r, k));
  }
  fence_proxy_async_smem();
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  if (tid == 0) {
    // a_major (bit 15) = b_major (bit 16) = 1: MN-major
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sa = smem_u32(A), sb = smem_u32(B);
    for (int ks = 0; ks < 2; ++ks) {
      const uint32_t lbo = swap ? rb_stride : kb_stride, sbo = swap ? kb_stride : rb_stride;
      const uint64_t adesc = smem_desc_kmajor_noswizzle(sa + ks * 2 * kb_stride, lbo, sbo);
      const uint64_t bdesc = smem_desc_kmajor_noswizzle(sb + ks * 2 * kb_stride, lbo, sbo);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" :: "r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(ks) : "memory");
    }
    mma_commit(&bar); mbar_wait(&bar, 0);
  }
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  for (int c0 = 0; c0 < 128; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[tid * 128 + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  float* d; long long* c;
  cudaMalloc(&d, 128 * 128 * sizeof(float)); cudaMalloc(&c, 16);
  cudaFuncSetAttribute(test, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int N : {64, 128}) {
    test<<<1, 128, 65536>>>(N, d, c);
    cudaError_t e = cudaDeviceSynchronize();
    static float h[128 * 128]; long long hc[2];
    cudaMemcpy(h, d, 128 * N * sizeof(float), cudaMemcpyDeviceToHost); cudaMemcpy(hc, c, 16, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
      double ref = 0; for (int k = 0; k < 32; ++k) ref += (double)a_val(m, k) * b_val(n, k);
      const double er = fabs(ref - h[m * N + n]); if (er > maxerr) maxerr = er; if (er > 1e-3) ++bad;
    }
    printf("f16 TS M128 N%d K16 x2: %s  max err %.3g  mismatches %d  | %.1f cycles/MMA\n", N, cudaGetErrorString(e), maxerr, bad, hc[0] / 2048.0);
  }
  cudaFuncSetAttribute(test_ss_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int swap : {0, 1}) {
    test_ss_mn<<<1, 128, 65536>>>(swap, d);
    cudaError_t e = cudaDeviceSynchronize();
    static float h[128 * 128];
    cudaMemcpy(h, d, 128 * 128 * sizeof(float), cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) {
      double ref = 0; for (int k = 0; k < 32; ++k) ref += (double)a_val(m, k) * b_val(n, k);
      const double er = fabs(ref - h[m * 128 + n]); if (er > maxerr) maxerr = er; if (er > 1e-3) ++bad;
    }
    printf("f16 SS MN-major M128 N128 K16 x2 (lbo/sbo swap=%d): %s  max err %.3g  mismatches %d\n", swap, cudaGetErrorString(e), maxerr, bad);
  }
  return 0;
}
