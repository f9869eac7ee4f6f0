// Host-buffer entry points: the calls a CPU-side caller (a tf.data worker, a cgo/JNI-style binding)
// makes with plain host arrays.  Device staging is owned by an opaque per-thread context; batches
// are cut into chunks so the H2D copy of chunk k+1 overlaps the kernels of chunk k.
//
// The Hellinger loss couples all images through one scalar (histogram.py:88-89): loss = sqrt(S) / (sqrt2 B)
// with S the sum over the whole (global) batch, hence the two phases.  But the gradient depends on S and B
// only through the common factor 1 / (B sqrt(S)): d loss / d fake_b = h_b / (B sqrt(S)) with h_b a function
// of image b alone.  So phase 1 already runs the backward kernels of every chunk with S = B = 1, right behind
// that chunk's forward kernels and under the upload of the next chunk, and phase 2 only multiplies the stored
// gradient by 1 / (B sqrt(S)) — one pass at HBM speed — once the global S is known.
#include <stdlib.h>

#include <new>

#include "common.cuh"
#include "hist_internal.cuh"

struct ph_host_ctx {
  int device = 0;
  cudaStream_t s_in = nullptr, s_compute = nullptr, s_compute2 = nullptr, s_out = nullptr;
  // grow-only device arena
  void* arena = nullptr;
  size_t arena_bytes = 0;
  static constexpr int kMaxChunks = 64;
  cudaEvent_t ev_in[kMaxChunks];
  cudaEvent_t ev_done[kMaxChunks];
  cudaEvent_t ev_free[2];
  cudaEvent_t ev_join;
  bool events_ready = false;
  // state carried from ph_host_hist_begin to ph_host_hist_finish
  struct Job {
    unsigned char* d_u8[2];
    float *d_fake, *d_real[2], *d_hreal, *d_hfake, *d_denom_r, *d_denom_f, *d_gradfull, *d_dom, *d_loss;
    double *d_ssum, *d_ssum2, *d_one;
    bool with_grad;
    char* d_ws;
    char* d_ws2;
    size_t ws_bytes;
    int64_t batch, npix, chunk;  // chunk = the largest chunk (staging buffers)
    int64_t start[kMaxChunks + 1];  // chunk k covers images [start[k], start[k+1])
    int channels, bins, method, impl, nchunks;
    float sigma_sqr, epsilon;
  } job;
  bool job_valid = false;
};

namespace ph {

// out[i] = in[i] / (B sqrt(S)): turns the unit-scale gradient of phase 1 into the gradient of the loss
// (S == 0 gives inf / NaN exactly like the reference's 0 * inf)
__global__ void __launch_bounds__(256) grad_rescale_kernel(const float4* __restrict__ in, float4* __restrict__ out,
                                                           int64_t n4, const double* __restrict__ ssum,
                                                           double global_batch) {
  const float c = (float)(1.0 / (global_batch * sqrt(*ssum)));
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = in[i];
    v.x *= c; v.y *= c; v.z *= c; v.w *= c;
    out[i] = v;
  }
}
__global__ void __launch_bounds__(256) grad_rescale_scalar_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                  int64_t n, const double* __restrict__ ssum,
                                                                  double global_batch) {
  const float c = (float)(1.0 / (global_batch * sqrt(*ssum)));
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[i] * c;
}
static int launch_grad_rescale(const float* in, float* out, int64_t n, const double* ssum, int64_t global_batch,
                               cudaStream_t st) {
  const bool vec = n % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  int64_t grid = ceil_div(vec ? n / 4 : n, 256 * 4);
  const int64_t cap = (int64_t)cached_sm_count() * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  if (vec)
    grad_rescale_kernel<<<(unsigned)grid, 256, 0, st>>>(reinterpret_cast<const float4*>(in),
                                                        reinterpret_cast<float4*>(out), n / 4, ssum, (double)global_batch);
  else
    grad_rescale_scalar_kernel<<<(unsigned)grid, 256, 0, st>>>(in, out, n, ssum, (double)global_batch);
  PH_LAUNCH_OK("grad_rescale_kernel");
  return PH_OK;
}

// *out = parts[0] (+ parts[1]): the two compute streams' shares of the shard's sum of squares
__global__ void sum_parts_kernel(const double* __restrict__ parts, int n, double* __restrict__ out) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += parts[i];
  *out = s;
}

static int ensure_arena(ph_host_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->arena_bytes) return PH_OK;
  if (ctx->arena) {
    PH_CUDA_OK(cudaDeviceSynchronize());
    PH_CUDA_OK(cudaFree(ctx->arena));
    ctx->arena = nullptr;
    ctx->arena_bytes = 0;
  }
  PH_CUDA_OK(cudaMalloc(&ctx->arena, bytes));
  ctx->arena_bytes = bytes;
  return PH_OK;
}

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

}  // namespace ph

using namespace ph;

extern "C" {

int ph_host_ctx_create(int device, ph_host_ctx** out) {
  PH_CHECK_ARG(out != nullptr, "ctx output pointer is NULL");
  PH_CUDA_OK(cudaSetDevice(device));
  ph_host_ctx* ctx = new (std::nothrow) ph_host_ctx();
  PH_CHECK_ARG(ctx != nullptr, "out of host memory");
  ctx->device = device;
  PH_CUDA_OK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
  PH_CUDA_OK(cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking));
  PH_CUDA_OK(cudaStreamCreateWithFlags(&ctx->s_compute2, cudaStreamNonBlocking));
  PH_CUDA_OK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < ph_host_ctx::kMaxChunks; ++i) {
    PH_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
    PH_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < 2; ++i) PH_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming));
  PH_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  ctx->events_ready = true;
  *out = ctx;
  return PH_OK;
}

void ph_host_ctx_destroy(ph_host_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->events_ready) {
    for (int i = 0; i < ph_host_ctx::kMaxChunks; ++i) {
      cudaEventDestroy(ctx->ev_in[i]);
      cudaEventDestroy(ctx->ev_done[i]);
    }
    for (int i = 0; i < 2; ++i) cudaEventDestroy(ctx->ev_free[i]);
    cudaEventDestroy(ctx->ev_join);
  }
  if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
  if (ctx->s_compute) cudaStreamDestroy(ctx->s_compute);
  if (ctx->s_compute2) cudaStreamDestroy(ctx->s_compute2);
  if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
  if (ctx->arena) cudaFree(ctx->arena);
  delete ctx;
}

// Two-phase form (ph_host_hist_begin / ph_host_hist_finish): the caller may all-reduce the local sum
// of squares over ranks between the phases; ph_host_hist_loss is the single-process composition.
static int host_hist_begin_impl(ph_host_ctx* ctx, const void* real_host, bool real_is_u8, const float* fake_host,
                                int64_t batch, int64_t npix, int channels, const float* bin_centers_host, int bins,
                                int method, float sigma_sqr, float epsilon, int impl, bool with_grad,
                                double* ssum_local_host) {
  // ssum_local_host == NULL: the shard's sum is not read back (no stream synchronisation here): the caller goes on
  // to the peer all-reduce on the device (ph_host_hist_loss_sharded)
  PH_CHECK_ARG(ctx && real_host && fake_host && bin_centers_host, "NULL pointer argument");
  PH_CHECK_ARG(batch > 0 && npix > 0 && (channels == 3 || channels == 4), "bad shape");
  PH_CHECK_ARG(bins >= 1 && bins <= 1024, "bins must be in [1,1024]");
  PH_CUDA_OK(cudaSetDevice(ctx->device));
  ctx->job_valid = false;
  // the bin centres are host data here, so the library checks the contract of PH_IMPL_MIRROR itself: antisymmetric
  // centres (tf.linspace(-3, 3, 64), histogram.py:55) let the fake images' forward share the weight vectors of +x and -x
  {
    float asym = 0.f;
    for (int j = 0; j < bins; ++j) asym = fmaxf(asym, fabsf(bin_centers_host[j] + bin_centers_host[bins - 1 - j]));
    if (asym <= 2e-5f * sqrtf(sigma_sqr)) impl |= PH_IMPL_MIRROR; else impl &= ~PH_IMPL_MIRROR;
  }

  // chunking: at least two images per SM per chunk (each chunk then takes the whole-image kernel path) and at
  // least 8 MiB per copy.  (Measured at cfgC: uniform 296-image chunks 520 k pairs/s; 148: 441 k; 592: 478 k; a
  // small first chunk followed by larger ones: 486 k.)
  const size_t img_bytes = (size_t)npix * channels * sizeof(float);
  int64_t base = (int64_t)((8u << 20) / img_bytes);
  if (base < 296) base = 296;
  static const int64_t chunk_env = getenv("PH_HOST_CHUNK") ? atoll(getenv("PH_HOST_CHUNK")) : 0;  // tuning knob
  ph_host_ctx::Job& J = ctx->job;
  int nchunks = 0;
  int64_t chunk = 0;
  // tuning knob: explicit chunk sizes, e.g. PH_HOST_CHUNK_LIST=148,296,592,1184 (the last size repeats)
  static const char* list_env = getenv("PH_HOST_CHUNK_LIST");
  if (list_env != nullptr && list_env[0] != 0) {
    int64_t pos = 0, sz = base;
    const char* q = list_env;
    J.start[0] = 0;
    while (pos < batch && nchunks < ph_host_ctx::kMaxChunks) {
      if (*q) {
        char* end = nullptr;
        const long long v = strtoll(q, &end, 10);
        if (v > 0) sz = v;
        q = (*end == ',') ? end + 1 : end;
      }
      int64_t take = sz < batch - pos ? sz : batch - pos;
      if (nchunks == ph_host_ctx::kMaxChunks - 1) take = batch - pos;
      pos += take;
      J.start[++nchunks] = pos;
      if (take > chunk) chunk = take;
    }
  } else {
    int64_t pos = 0;
    J.start[0] = 0;
    while (pos < batch) {
      int64_t sz = chunk_env > 0 ? chunk_env : base;
      const int64_t left = batch - pos;
      const int slots_left = ph_host_ctx::kMaxChunks - nchunks;
      if (slots_left * sz < left) sz = ceil_div(left, slots_left);  // huge batches: stay within kMaxChunks
      if (sz > left || left - sz < base / 2) sz = left;               // no tiny last chunk
      pos += sz;
      J.start[++nchunks] = pos;
      if (sz > chunk) chunk = sz;
    }
  }
  const size_t hist_elems = (size_t)bins * bins * 3;
  size_t ws_bytes = 0;  // the workspace depends on how a chunk splits into whole and sliced images: take the largest
  for (int k = 0; k < nchunks; ++k) {
    const size_t w = ph_hist_workspace_bytes(J.start[k + 1] - J.start[k], npix, bins, impl);
    if (w > ws_bytes) ws_bytes = w;
  }

  for (int pass = 0; pass < 2; ++pass) {
    Carver cv(pass == 0 ? nullptr : ctx->arena);
    J.d_fake = cv.take<float>((size_t)batch * npix * channels);
    J.d_real[0] = cv.take<float>((size_t)chunk * npix * channels);
    J.d_real[1] = cv.take<float>((size_t)chunk * npix * channels);
    J.d_hreal = cv.take<float>((size_t)batch * hist_elems);
    J.d_hfake = cv.take<float>((size_t)batch * hist_elems);
    J.d_denom_r = cv.take<float>((size_t)batch);
    J.d_denom_f = cv.take<float>((size_t)batch);
    J.d_gradfull = with_grad ? cv.take<float>((size_t)batch * npix * channels) : nullptr;
    J.d_dom = cv.take<float>((size_t)bins);
    J.d_ssum = cv.take<double>(1);
    J.d_one = cv.take<double>(1);
    J.d_ssum2 = cv.take<double>(2);  // per compute stream
    J.d_loss = cv.take<float>(1);
    J.d_ws = cv.take<char>(ws_bytes);
    J.d_ws2 = cv.take<char>(ws_bytes);
    J.d_u8[0] = real_is_u8 ? cv.take<unsigned char>((size_t)chunk * npix * 4) : nullptr;
    J.d_u8[1] = real_is_u8 ? cv.take<unsigned char>((size_t)chunk * npix * 4) : nullptr;
    if (pass == 0) {
      int rc = ensure_arena(ctx, align_up(cv.off, 256) + 256);
      if (rc != PH_OK) return rc;
    }
  }
  J.batch = batch; J.npix = npix; J.channels = channels; J.bins = bins; J.method = method;
  J.sigma_sqr = sigma_sqr; J.epsilon = epsilon; J.impl = impl; J.chunk = chunk; J.nchunks = nchunks;
  J.ws_bytes = ws_bytes;
  J.with_grad = with_grad;

  static const double kOne = 1.0;
  PH_CUDA_OK(cudaMemcpyAsync(J.d_one, &kOne, sizeof(double), cudaMemcpyHostToDevice, ctx->s_in));
  PH_CUDA_OK(cudaMemcpyAsync(J.d_dom, bin_centers_host, sizeof(float) * bins, cudaMemcpyHostToDevice, ctx->s_in));
  // ---- phase 1: upload + forward + unit-scale backward, chunk by chunk ----
  for (int k = 0; k < nchunks; ++k) {
    const int64_t b0 = J.start[k];
    const int64_t nb = J.start[k + 1] - b0;
    const size_t n = (size_t)nb * npix * channels;
    const int slot = k & 1;
    if (k >= 2) PH_CUDA_OK(cudaStreamWaitEvent(ctx->s_in, ctx->ev_free[slot], 0));
    if (real_is_u8) {
      PH_CUDA_OK(cudaMemcpyAsync(J.d_u8[slot], static_cast<const unsigned char*>(real_host) + (size_t)b0 * npix * 4,
                                 (size_t)nb * npix * 4, cudaMemcpyHostToDevice, ctx->s_in));
    } else {
      PH_CUDA_OK(cudaMemcpyAsync(J.d_real[slot], static_cast<const float*>(real_host) + (size_t)b0 * npix * channels,
                                 n * sizeof(float), cudaMemcpyHostToDevice, ctx->s_in));
    }
    PH_CUDA_OK(cudaMemcpyAsync(J.d_fake + (size_t)b0 * npix * channels, fake_host + (size_t)b0 * npix * channels,
                               n * sizeof(float), cudaMemcpyHostToDevice, ctx->s_in));
    PH_CUDA_OK(cudaEventRecord(ctx->ev_in[k], ctx->s_in));
    // chunks alternate between two compute streams: the kernels of chunk k+1 fill the SMs that the (persistent)
    // kernels of chunk k vacate in their tails, and the small prologue kernels run beside the large ones
    cudaStream_t sc = slot == 0 ? ctx->s_compute : ctx->s_compute2;
    char* ws = slot == 0 ? J.d_ws : J.d_ws2;
    PH_CUDA_OK(cudaStreamWaitEvent(sc, ctx->ev_in[k], 0));
    if (real_is_u8) {  // blacken + normalise on the device (dataset_utils.py:66-77)
      int rcu = ph_u8_to_float_image(J.d_u8[slot], nb * npix, 1, 1, J.d_real[slot], sc);
      if (rcu != PH_OK) return rcu;
    }
    int rc = ph_hist_forward(J.d_real[slot], nb, npix, channels, J.d_dom, bins, method, sigma_sqr, epsilon,
                             J.d_hreal + (size_t)b0 * hist_elems, J.d_denom_r + b0, ws, ws_bytes,
                             impl | PH_IMPL_DEDUP,  // real images are dataset sprites: contract unique colours
                             sc);
    if (rc != PH_OK) return rc;
    PH_CUDA_OK(cudaEventRecord(ctx->ev_free[slot], sc));
    // forward of the fake chunk fused with its share of the Hellinger sum (added into d_ssum[slot])
    rc = ph_hist_forward_ssum(J.d_fake + (size_t)b0 * npix * channels, nb, npix, channels, J.d_dom, bins, method,
                              sigma_sqr, epsilon, J.d_hfake + (size_t)b0 * hist_elems, J.d_denom_f + b0,
                              J.d_hreal + (size_t)b0 * hist_elems, J.d_ssum2 + slot, k >= 2, ws, ws_bytes, impl, sc);
    if (rc != PH_OK) return rc;
    if (with_grad) {
      rc = ph_hist_backward(J.d_fake + (size_t)b0 * npix * channels, nb, npix, channels, J.d_dom, bins, method, sigma_sqr,
                            epsilon, J.d_hfake + (size_t)b0 * hist_elems, J.d_denom_f + b0, nullptr,
                            J.d_hreal + (size_t)b0 * hist_elems, J.d_one, 1, nullptr,
                            J.d_gradfull + (size_t)b0 * npix * channels, ws, ws_bytes, impl, sc);
      if (rc != PH_OK) return rc;
    }
  }
  // ---- the one coupling scalar ----
  PH_CUDA_OK(cudaEventRecord(ctx->ev_join, ctx->s_compute2));
  PH_CUDA_OK(cudaStreamWaitEvent(ctx->s_compute, ctx->ev_join, 0));
  // the shard's sum stays on the device (ph_host_hist_finish_comm all-reduces it there) and is also returned
  sum_parts_kernel<<<1, 1, 0, ctx->s_compute>>>(J.d_ssum2, nchunks > 1 ? 2 : 1, J.d_ssum);
  PH_LAUNCH_OK("sum_parts_kernel");
  if (ssum_local_host != nullptr) {
    PH_CUDA_OK(cudaMemcpyAsync(ssum_local_host, J.d_ssum, sizeof(double), cudaMemcpyDeviceToHost, ctx->s_compute));
    PH_CUDA_OK(cudaStreamSynchronize(ctx->s_compute));
  }
  ctx->job_valid = true;
  return PH_OK;
}

int ph_host_hist_begin(ph_host_ctx* ctx, const float* real_host, const float* fake_host, int64_t batch,
                       int64_t npix, int channels, const float* bin_centers_host, int bins, int method,
                       float sigma_sqr, float epsilon, int impl, double* ssum_local_host) {
  PH_CHECK_ARG(ssum_local_host != nullptr, "NULL pointer argument");
  return host_hist_begin_impl(ctx, real_host, false, fake_host, batch, npix, channels, bin_centers_host, bins, method,
                              sigma_sqr, epsilon, impl, true, ssum_local_host);
}

int ph_host_hist_begin_u8real(ph_host_ctx* ctx, const uint8_t* real_u8_host, const float* fake_host, int64_t batch,
                              int64_t npix, const float* bin_centers_host, int bins, int method, float sigma_sqr,
                              float epsilon, int impl, double* ssum_local_host) {
  PH_CHECK_ARG(ssum_local_host != nullptr, "NULL pointer argument");
  return host_hist_begin_impl(ctx, real_u8_host, true, fake_host, batch, npix, 4, bin_centers_host, bins, method,
                              sigma_sqr, epsilon, impl, true, ssum_local_host);
}

static int host_hist_finish_impl(ph_host_ctx* ctx, ph_comm* comm, double ssum_global, int64_t global_batch,
                                 float* loss_host, float* grad_fake_host, float* grad_fake_device) {
  PH_CHECK_ARG(ctx && loss_host, "NULL pointer argument");
  PH_CHECK_ARG(ctx->job_valid, "ph_host_hist_finish without a preceding successful ph_host_hist_begin");
  PH_CHECK_ARG(global_batch > 0, "global_batch must be positive");
  PH_CUDA_OK(cudaSetDevice(ctx->device));
  ctx->job_valid = false;
  const ph_host_ctx::Job& J = ctx->job;
  int rc;
  if (comm != nullptr) {  // sum over the ranks on the device, over peer memory: no host round trip
    rc = ph_comm_allreduce_sum_f64(comm, J.d_ssum, 1, ctx->s_compute);
    if (rc != PH_OK) return rc;
  } else {
    PH_CUDA_OK(cudaMemcpyAsync(J.d_ssum, &ssum_global, sizeof(double), cudaMemcpyHostToDevice, ctx->s_compute));
  }
  rc = ph_hellinger_finish(J.d_ssum, global_batch, J.d_loss, ctx->s_compute);
  if (rc != PH_OK) return rc;
  PH_CUDA_OK(cudaMemcpyAsync(loss_host, J.d_loss, sizeof(float), cudaMemcpyDeviceToHost, ctx->s_compute));
  // ---- phase 2: scale the stored unit gradient by 1 / (B sqrt(S)); it either stays on the device (the consumer
  //      — the generator's backward — lives there) or is downloaded chunk by chunk behind its scaling pass ----
  PH_CHECK_ARG(J.with_grad || (!grad_fake_device && !grad_fake_host), "this job was started without a gradient");
  const int64_t n_all = J.batch * J.npix * J.channels;
  if (grad_fake_device) {
    rc = launch_grad_rescale(J.d_gradfull, grad_fake_device, n_all, J.d_ssum, global_batch, ctx->s_compute);
    if (rc != PH_OK) return rc;
  }
  if (grad_fake_host) {
    for (int k = 0; k < J.nchunks; ++k) {
      const int64_t b0 = J.start[k];
      const int64_t nb = J.start[k + 1] - b0;
      const int64_t n = nb * J.npix * J.channels;
      float* g = J.d_gradfull + (size_t)b0 * J.npix * J.channels;
      rc = launch_grad_rescale(g, g, n, J.d_ssum, global_batch, ctx->s_compute);
      if (rc != PH_OK) return rc;
      PH_CUDA_OK(cudaEventRecord(ctx->ev_done[k], ctx->s_compute));
      PH_CUDA_OK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_done[k], 0));
      PH_CUDA_OK(cudaMemcpyAsync(grad_fake_host + (size_t)b0 * J.npix * J.channels, g, (size_t)n * sizeof(float),
                                 cudaMemcpyDeviceToHost, ctx->s_out));
    }
  }
  PH_CUDA_OK(cudaStreamSynchronize(ctx->s_compute));
  PH_CUDA_OK(cudaStreamSynchronize(ctx->s_out));
  return PH_OK;
}

int ph_host_hist_finish(ph_host_ctx* ctx, double ssum_global, int64_t global_batch, float* loss_host,
                        float* grad_fake_host, float* grad_fake_device) {
  return host_hist_finish_impl(ctx, nullptr, ssum_global, global_batch, loss_host, grad_fake_host, grad_fake_device);
}

int ph_host_hist_finish_comm(ph_host_ctx* ctx, ph_comm* comm, int64_t global_batch, float* loss_host,
                             float* grad_fake_host, float* grad_fake_device) {
  PH_CHECK_ARG(comm != nullptr, "comm must not be NULL");
  return host_hist_finish_impl(ctx, comm, 0.0, global_batch, loss_host, grad_fake_host, grad_fake_device);
}

int ph_host_hist_loss_sharded(ph_host_ctx* ctx, ph_comm* comm, const void* real_host, int real_is_u8,
                              const float* fake_host, int64_t batch, int64_t npix, int channels,
                              const float* bin_centers_host, int bins, int method, float sigma_sqr, float epsilon,
                              int impl, int64_t global_batch, float* loss_host, float* grad_fake_host,
                              float* grad_fake_device) {
  PH_CHECK_ARG(comm != nullptr, "comm must not be NULL");
  PH_CHECK_ARG(!real_is_u8 || channels == 4, "uint8 real images are RGBA");
  int rc = host_hist_begin_impl(ctx, real_host, real_is_u8 != 0, fake_host, batch, npix, channels, bin_centers_host, bins,
                                method, sigma_sqr, epsilon, impl, grad_fake_host != nullptr || grad_fake_device != nullptr,
                                nullptr);
  if (rc != PH_OK) return rc;
  return host_hist_finish_impl(ctx, comm, 0.0, global_batch, loss_host, grad_fake_host, grad_fake_device);
}

int ph_host_hist_loss(ph_host_ctx* ctx, const float* real_host, const float* fake_host, int64_t batch,
                      int64_t npix, int channels, const float* bin_centers_host, int bins, int method,
                      float sigma_sqr, float epsilon, int impl, float* loss_host, float* grad_fake_host) {
  double ssum = 0.0;
  int rc = host_hist_begin_impl(ctx, real_host, false, fake_host, batch, npix, channels, bin_centers_host, bins, method,
                                sigma_sqr, epsilon, impl, grad_fake_host != nullptr, &ssum);
  if (rc != PH_OK) return rc;
  return ph_host_hist_finish(ctx, ssum, batch, loss_host, grad_fake_host, nullptr);
}

}  // extern "C"

// source / target as int32 (elem_bytes 4) or as the decoded PNG's uint8 (elem_bytes 1: a quarter of the upload,
// widened to int32 on the device)
static int host_load_indexed(ph_host_ctx* ctx, const void* source_host, const void* target_host, int elem_bytes,
                             int64_t batch, int64_t npix, int ordering, const float* shuffle_keys_host,
                             int32_t* source_indexed_host,
                             int32_t* target_indexed_host, int32_t* palette_host, int32_t* ncolors_host,
                             float* target_one_hot_host) {
  PH_CHECK_ARG(ctx && source_host && target_host && source_indexed_host && target_indexed_host && palette_host &&
                   ncolors_host,
               "NULL pointer argument");
  PH_CHECK_ARG(batch > 0 && npix > 0, "bad shape");
  PH_CHECK_ARG(ordering != PH_ORDER_SHUFFLED || shuffle_keys_host != nullptr, "'shuffled' ordering needs shuffle keys");
  PH_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t img = (size_t)batch * npix * 4;
  const int depth = PH_MAX_PALETTE_SIZE;
  for (int pass = 0; pass < 2; ++pass) {
    Carver cv(pass == 0 ? nullptr : ctx->arena);
    // pixels in the element type the caller has them in: the kernel reads uint8 RGBA as it is (4 B per pixel)
    void* d_src = cv.take<char>(img * elem_bytes);
    void* d_tgt = cv.take<char>(img * elem_bytes);
    int32_t* d_sidx = cv.take<int32_t>((size_t)batch * npix);
    int32_t* d_tidx = cv.take<int32_t>((size_t)batch * npix);
    int32_t* d_pal = cv.take<int32_t>((size_t)batch * depth * 4);
    int32_t* d_nc = cv.take<int32_t>((size_t)batch);
    float* d_keys = ordering == PH_ORDER_SHUFFLED ? cv.take<float>((size_t)batch * depth) : nullptr;
    float* d_oh = target_one_hot_host ? cv.take<float>((size_t)batch * npix * depth) : nullptr;
    if (pass == 0) {
      int rc = ensure_arena(ctx, align_up(cv.off, 256) + 256);
      if (rc != PH_OK) return rc;
      continue;
    }
    cudaStream_t st = ctx->s_compute;
    int rc;
    PH_CUDA_OK(cudaMemcpyAsync(d_src, source_host, img * elem_bytes, cudaMemcpyHostToDevice, st));
    PH_CUDA_OK(cudaMemcpyAsync(d_tgt, target_host, img * elem_bytes, cudaMemcpyHostToDevice, st));
    if (d_keys)
      PH_CUDA_OK(cudaMemcpyAsync(d_keys, shuffle_keys_host, (size_t)batch * depth * sizeof(float), cudaMemcpyHostToDevice, st));
    if (elem_bytes == 1)
      rc = ph_load_indexed_images_u8(static_cast<const uint8_t*>(d_src), static_cast<const uint8_t*>(d_tgt), batch, npix,
                                     ordering, d_keys, d_sidx, d_tidx, d_pal, d_nc, st);
    else
      rc = ph_load_indexed_images(static_cast<const int32_t*>(d_src), static_cast<const int32_t*>(d_tgt), batch, npix,
                                  ordering, d_keys, d_sidx, d_tidx, d_pal, d_nc, st);
    if (rc != PH_OK) return rc;
    if (d_oh) {
      rc = ph_one_hot(d_tidx, batch * npix, depth, d_oh, st);
      if (rc != PH_OK) return rc;
    }
    PH_CUDA_OK(cudaMemcpyAsync(source_indexed_host, d_sidx, (size_t)batch * npix * 4, cudaMemcpyDeviceToHost, st));
    PH_CUDA_OK(cudaMemcpyAsync(target_indexed_host, d_tidx, (size_t)batch * npix * 4, cudaMemcpyDeviceToHost, st));
    PH_CUDA_OK(cudaMemcpyAsync(palette_host, d_pal, (size_t)batch * depth * 16, cudaMemcpyDeviceToHost, st));
    PH_CUDA_OK(cudaMemcpyAsync(ncolors_host, d_nc, (size_t)batch * 4, cudaMemcpyDeviceToHost, st));
    if (d_oh)
      PH_CUDA_OK(cudaMemcpyAsync(target_one_hot_host, d_oh, (size_t)batch * npix * depth * 4, cudaMemcpyDeviceToHost, st));
    PH_CUDA_OK(cudaStreamSynchronize(st));
  }
  return PH_OK;
}

extern "C" {

int ph_host_load_indexed_images(ph_host_ctx* ctx, const int32_t* source_host, const int32_t* target_host,
                                int64_t batch, int64_t npix, int ordering, const float* shuffle_keys_host,
                                int32_t* source_indexed_host, int32_t* target_indexed_host, int32_t* palette_host,
                                int32_t* ncolors_host, float* target_one_hot_host) {
  return host_load_indexed(ctx, source_host, target_host, 4, batch, npix, ordering, shuffle_keys_host, source_indexed_host,
                           target_indexed_host, palette_host, ncolors_host, target_one_hot_host);
}

int ph_host_load_indexed_images_u8(ph_host_ctx* ctx, const uint8_t* source_host, const uint8_t* target_host,
                                   int64_t batch, int64_t npix, int ordering, const float* shuffle_keys_host,
                                   int32_t* source_indexed_host, int32_t* target_indexed_host, int32_t* palette_host,
                                   int32_t* ncolors_host, float* target_one_hot_host) {
  return host_load_indexed(ctx, source_host, target_host, 1, batch, npix, ordering, shuffle_keys_host, source_indexed_host,
                           target_indexed_host, palette_host, ncolors_host, target_one_hot_host);
}

}  // extern "C"
