// Micro-benchmark: cycles per tcgen05.mma for the shapes the histogram kernels use.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../palette_and_histo_gan_b200/csrc/tc_ptx.cuh"
using namespace ph::tc;

__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int MODE>  // 0: tf32 TS, 1: tf32 SS, 2: bf16 TS
__global__ void bench(int n_mma, int M, int N, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tmem = tmem_base;
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t idesc = idesc_tf32(M, N);
    if (MODE == 2) idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t sb = smem_u32(smem);
    const uint64_t bdesc = smem_desc_kmajor_noswizzle(sb, 2048, 128);
    const uint64_t adesc = smem_desc_kmajor_noswizzle(sb + 32768, 2048, 128);
    // warm-up
    for (int i = 0; i < 8; ++i) {
      if (MODE == 1) mma_tf32_ss(tmem, adesc, bdesc, idesc, 1); else if (MODE == 0) mma_tf32_ts(tmem, tmem + 256, bdesc, idesc, 1);
      else mma_f16_ts(tmem, tmem + 256, bdesc, idesc, 1);
    }
    mma_commit(&bar); mbar_wait(&bar, 0);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      if (MODE == 1) mma_tf32_ss(tmem, adesc, bdesc, idesc, 1); else if (MODE == 0) mma_tf32_ts(tmem, tmem + 256 + (i & 3) * 8, bdesc, idesc, 1);
      else mma_f16_ts(tmem, tmem + 256 + (i & 3) * 8, bdesc, idesc, 1);
    }
    mma_commit(&bar); mbar_wait(&bar, 1);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  const int n = 4096;
  struct Cfg { int mode, M, N; const char* name; } cfgs[] = {
    {0, 128, 64, "tf32 TS M128 N64 K8"}, {0, 128, 128, "tf32 TS M128 N128 K8"}, {0, 128, 256, "tf32 TS M128 N256 K8"},
    {0, 64, 64, "tf32 TS M64 N64 K8"}, {1, 128, 64, "tf32 SS M128 N64 K8"}, {1, 128, 256, "tf32 SS M128 N256 K8"},
    {2, 128, 64, "bf16 TS M128 N64 K16"}, {2, 128, 256, "bf16 TS M128 N256 K16"}};
  for (auto& c : cfgs) {
    for (int grid : {1, 148}) {
      cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      if (c.mode == 0) bench<0><<<grid, 128, 65536>>>(n, c.M, c.N, d);
      else if (c.mode == 1) bench<1><<<grid, 128, 65536>>>(n, c.M, c.N, d);
      else bench<2><<<grid, 128, 65536>>>(n, c.M, c.N, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%-28s grid %3d: %s  %.1f cycles/MMA (max over CTAs)\n", c.name, grid, cudaGetErrorString(e), (double)mx / n);
    }
  }
  return 0;
}
