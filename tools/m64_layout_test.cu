// tcgen05.mma kind::f16 with M = 64 and the A operand in TMEM: which lanes hold the rows, may the A tile and the
// accumulator sit at lane offset 16 ("interleaved" second tile of a 32-lane quadrant), and what does an M = 64
// instruction cost?  Row m of a 64-row tile lives in quadrant m / 16, lane (m % 16) + 16 t for tile t in {0, 1}.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
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

// K = 32 (two K16 steps); a_tile / d_tile: lane offset 16 * tile of the A operand / the accumulator
__global__ void test(int N, int a_tile, int d_tile, float* out, long long* cyc) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tmem = tmem_base;
  const int kb_stride = N / 8 * 128;
  for (int e = tid; e < N * 32; e += blockDim.x) {
    const int n = e / 32, k = e % 32;
    __half* dst = reinterpret_cast<__half*>(smem + (k / 8) * kb_stride + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2);
    *dst = __float2half(b_val(n, k));
  }
  {  // every thread writes the A row of ITS lane: tile t = lane / 16, row m = 16 warp + lane % 16 (value tagged by tile)
    const int t = lane >> 4, m = 16 * warp + (lane & 15);
    uint32_t w[16];
    for (int c = 0; c < 16; ++c) {
      const __half2 h = __floats2half2_rn(a_val(m + 64 * t, 2 * c), a_val(m + 64 * t, 2 * c + 1));
      w[c] = *reinterpret_cast<const uint32_t*>(&h);
    }
    tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 256, w);
    uint32_t z[16];
    for (int c = 0; c < 16; ++c) z[c] = 0x7fc00000u;  // NaN: rows the MMA must overwrite
    for (int c0 = 0; c0 < 128; c0 += 16) tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + c0, z);
    tmem_st_wait();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);  // M = 64
    const uint32_t sb = smem_u32(smem);
    const uint32_t d_addr = tmem + ((uint32_t)(16 * d_tile) << 16), a_addr = tmem + ((uint32_t)(16 * a_tile) << 16) + 256;
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t bdesc = smem_desc_kmajor_noswizzle(sb + ks * 2 * kb_stride, kb_stride, 128);
      mma_f16_ts(d_addr, a_addr + ks * 8, bdesc, idesc, ks);
    }
    mma_commit(&bar); mbar_wait(&bar, 0);
    const uint64_t bdesc = smem_desc_kmajor_noswizzle(sb, kb_stride, 128);
    const long long t1 = clock64();
    for (int i = 0; i < 2048; ++i) mma_f16_ts(d_addr + 128, a_addr + (i & 1) * 8, bdesc, idesc, 1);
    mma_commit(&bar); mbar_wait(&bar, 1);
    cyc[0] = clock64() - t1;
  }
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {  // out[lane-tile][row][col]: what every TMEM lane holds
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    const int t = lane >> 4, m = 16 * warp + (lane & 15);
    for (int i = 0; i < 16; ++i) out[(t * 64 + m) * N + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;  // run one combination per process: an illegal address poisons the context
  int combo = 0;
  float* d; long long* c;
  cudaMalloc(&d, 2 * 64 * 128 * sizeof(float)); cudaMalloc(&c, 16);
  cudaFuncSetAttribute(test, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int N : {64, 128})
    for (int at : {0, 1})
      for (int dt : {0, 1}) {
        if (only >= 0 && combo++ != only) continue;
        test<<<1, 128, 65536>>>(N, at, dt, d, c);
        cudaError_t e = cudaDeviceSynchronize();
        static float h[2 * 64 * 128]; long long hc[2] = {0, 0};
        cudaMemcpy(h, d, 2 * 64 * N * sizeof(float), cudaMemcpyDeviceToHost); cudaMemcpy(hc, c, 16, cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0, untouched_ok = 0;
        for (int m = 0; m < 64; ++m) for (int n = 0; n < N; ++n) {
          double ref = 0; for (int k = 0; k < 32; ++k) ref += (double)a_val(m + 64 * at, k) * b_val(n, k);
          const float got = h[(dt * 64 + m) * N + n];
          const double er = got == got ? fabs(ref - got) : 1e9; if (er > maxerr) maxerr = er; if (er > 1e-3) ++bad;
          const float other = h[((1 - dt) * 64 + m) * N + n];
          if (other != other) ++untouched_ok;  // the other lane tile still holds its NaN fill
        }
        printf("f16 TS M64 N%d K16 x2, A at lane offset %2d, D at lane offset %2d: %s  max err %.3g  mismatches %d  other tile untouched %d/%d | %.1f cycles/MMA\n",
               N, 16 * at, 16 * dt, cudaGetErrorString(e), maxerr, bad, untouched_ok, 64 * N, hc[0] / 2048.0);
        if (e != cudaSuccess) { cudaDeviceReset(); cudaMalloc(&d, 2 * 64 * 128 * sizeof(float)); cudaMalloc(&c, 16); cudaFuncSetAttribute(test, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); }
      }
  return 0;
}
