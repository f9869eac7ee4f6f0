// extern "C" surface of libpalhist.so — argument validation and engine dispatch only.
// See include/palhist.h for the contract and the reference file:line each entry point replaces.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "common.cuh"
#include "hist_internal.cuh"

namespace ph {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};  // process-wide: autograd runs backward on its own thread

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
  static const bool on = !(getenv("PH_PDL") && atoi(getenv("PH_PDL")) == 0);
  return on;
}

int cached_sm_count() {
  static int counts[64];
  static std::once_flag once;
  std::call_once(once, [] { memset(counts, 0, sizeof(counts)); });
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (counts[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    counts[dev] = n;
  }
  return counts[dev];
}

// sticky asynchronous status: one word of mapped (zero-copy) host memory per device, written by kernels with plain
// stores and read by the host without any synchronisation
static int* g_status_host[64];
static int* g_status_dev[64];
static std::mutex g_status_mutex;
int* async_status_word() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(g_status_mutex);
  if (g_status_dev[dev] == nullptr) {
    int* h = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&h), sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    *h = 0;
    int* d = nullptr;
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), h, 0) != cudaSuccess) {
      cudaGetLastError();
      cudaFreeHost(h);
      return nullptr;
    }
    g_status_host[dev] = h;
    g_status_dev[dev] = d;
  }
  return g_status_dev[dev];
}
// entry check of the ph_hist_* calls: an earlier launch on this device left the range flag
static int consume_async_status() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return PH_OK;
  volatile int* h = g_status_host[dev];
  if (h == nullptr || *h == 0) return PH_OK;
  const int bits = *h;
  *h = 0;
  if (bits & PH_ASYNC_MIRROR) {
    set_error("an earlier histogram forward was launched with PH_IMPL_MIRROR on bin centres that are not antisymmetric; its "
              "results are off by the asymmetry — re-run that call without the flag");
    return PH_ERR_UNSUPPORTED;
  }
  set_error("an earlier histogram launch on the tensor-core engine met pixels outside its operand range (image far "
            "outside [-1, 1]); its results are inf / NaN — re-run that call with the CUDA-core engine (impl = simt)");
  return PH_ERR_UNSUPPORTED;
}

static int resolve_impl(int impl, int64_t npix, int bins, int method) {
  impl &= PH_IMPL_ENGINE_MASK;
  if (impl == PH_IMPL_AUTO) return tc_supported(npix, bins, method) ? PH_IMPL_TC : PH_IMPL_SIMT;
  return impl;
}

static int check_hist_args(const void* image, int64_t batch, int64_t npix, int channels,
                           const void* dom, int bins, int method, float sigma_sqr) {
  PH_CHECK_ARG(image != nullptr && dom != nullptr, "image / bin_centers must not be NULL");
  PH_CHECK_ARG(batch >= 0 && npix > 0, "bad shape: batch=%lld npix=%lld", (long long)batch, (long long)npix);
  PH_CHECK_ARG(channels == 3 || channels == 4, "channels must be 3 or 4 (got %d)", channels);
  PH_CHECK_ARG(bins >= 1 && bins <= 1024, "bins must be in [1,1024] (got %d)", bins);
  PH_CHECK_ARG(method == PH_METHOD_INVERSE_QUADRATIC || method == PH_METHOD_RBF,
               "unknown histogram method %d (the reference silently mis-computes here, histogram.py:22-27)", method);
  PH_CHECK_ARG(sigma_sqr > 0.f, "sigma_sqr must be positive");
  PH_CHECK_ARG(channels != 4 || (reinterpret_cast<uintptr_t>(image) & 15) == 0,
               "RGBA image pointer must be 16-byte aligned");
  return PH_OK;
}

}  // namespace ph

using namespace ph;

extern "C" {

int ph_abi_version(void) { return PH_ABI_VERSION; }
const char* ph_last_error(void) { return g_error; }
int64_t ph_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void ph_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

int ph_async_status(int device, int clear) {
  if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return 0;
  if (device < 0 || device >= 64) return 0;
  volatile int* h = g_status_host[device];
  if (h == nullptr) return 0;
  const int v = *h;
  if (clear) *h = 0;
  return v;
}

int ph_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
  cudaDeviceProp prop;
  PH_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return PH_OK;
}

int ph_hist256_plan(int64_t batch, int64_t npix, int64_t* plan4) {
  PH_CHECK_ARG(plan4 != nullptr && batch >= 1 && npix >= 1 && npix < (1ll << 31), "bad argument");
  int slices = 1, items = 1, tpi = 1;
  int64_t pps = 0;
  tc_fwd256_plan(batch, npix, &slices, &pps);
  tc_bwd256_plan(batch, npix, &items, &tpi);
  plan4[0] = slices; plan4[1] = pps; plan4[2] = items; plan4[3] = tpi;
  return PH_OK;
}

size_t ph_hist_workspace_bytes(int64_t batch, int64_t npix, int bins, int impl) {
  if (batch <= 0 || npix <= 0 || bins <= 0) return 256;
  size_t s = simt_workspace_bytes(batch, npix, bins);
  if ((impl & PH_IMPL_ENGINE_MASK) != PH_IMPL_SIMT) {
    const size_t t = tc_workspace_bytes(batch, npix, bins);
    if (t > s) s = t;
  }
  return s;
}

static int hist_forward_impl(const float* image, int64_t batch, int64_t npix, int channels, const float* bin_centers,
                             int bins, int method, float sigma_sqr, float epsilon, float* hist, float* denom,
                             const float* hist_true, double* ssum, int accumulate, void* workspace,
                             size_t workspace_bytes, int impl, void* stream) {
  int rc = check_hist_args(image, batch, npix, channels, bin_centers, bins, method, sigma_sqr);
  if (rc != PH_OK) return rc;
  if ((rc = consume_async_status()) != PH_OK) return rc;
  PH_CHECK_ARG(hist != nullptr && denom != nullptr, "hist / denom must not be NULL");
  PH_CHECK_ARG((impl & ~(PH_IMPL_ENGINE_MASK | PH_IMPL_DEDUP | PH_IMPL_MIRROR)) == 0 && (impl & PH_IMPL_ENGINE_MASK) <= PH_IMPL_TC,
               "bad impl %d", impl);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ssum != nullptr && !accumulate) PH_CUDA_OK(cudaMemsetAsync(ssum, 0, sizeof(double), st));
  if (batch == 0) return PH_OK;
  const int eng = resolve_impl(impl, npix, bins, method);
  PH_CHECK_ARG(workspace != nullptr && workspace_bytes >= ph_hist_workspace_bytes(batch, npix, bins, eng),
               "workspace too small: %zu < %zu", workspace_bytes, ph_hist_workspace_bytes(batch, npix, bins, eng));
  if (eng == PH_IMPL_TC) {
    if (!tc_supported(npix, bins, method)) {
      set_error("tensor-core engine does not cover npix=%lld bins=%d method=%d", (long long)npix, bins, method);
      return PH_ERR_UNSUPPORTED;
    }
    return tc_hist_forward(image, batch, npix, channels, bin_centers, bins, method, sigma_sqr, epsilon, hist,
                           denom, workspace, (impl & PH_IMPL_DEDUP) != 0, (impl & PH_IMPL_MIRROR) != 0, hist_true, ssum, st);
  }
  rc = simt_hist_forward(image, batch, npix, channels, bin_centers, bins, method, sigma_sqr, epsilon, hist, denom,
                         workspace, st);
  if (rc != PH_OK || ssum == nullptr) return rc;
  return launch_hellinger_ssum_accumulate(hist_true, hist, batch * (int64_t)bins * bins * 3, ssum, st);
}

int ph_hist_forward(const float* image, int64_t batch, int64_t npix, int channels,
                    const float* bin_centers, int bins, int method, float sigma_sqr, float epsilon,
                    float* hist, float* denom, void* workspace, size_t workspace_bytes, int impl,
                    void* stream) {
  return hist_forward_impl(image, batch, npix, channels, bin_centers, bins, method, sigma_sqr, epsilon, hist, denom,
                           nullptr, nullptr, 0, workspace, workspace_bytes, impl, stream);
}

int ph_hist_forward_ssum(const float* image, int64_t batch, int64_t npix, int channels, const float* bin_centers,
                         int bins, int method, float sigma_sqr, float epsilon, float* hist, float* denom,
                         const float* hist_true, double* ssum, int accumulate, void* workspace,
                         size_t workspace_bytes, int impl, void* stream) {
  PH_CHECK_ARG(hist_true != nullptr && ssum != nullptr, "hist_true / ssum must not be NULL");
  PH_CHECK_ARG((reinterpret_cast<uintptr_t>(hist_true) & 15) == 0 && (reinterpret_cast<uintptr_t>(hist) & 15) == 0,
               "histogram pointers must be 16-byte aligned");
  return hist_forward_impl(image, batch, npix, channels, bin_centers, bins, method, sigma_sqr, epsilon, hist, denom,
                           hist_true, ssum, accumulate, workspace, workspace_bytes, impl, stream);
}

int ph_component_histogram(const float* component, const float* projection1, const float* projection2,
                           const float* color_intensities, int64_t batch, int64_t npix,
                           const float* bin_centers, int bins, int method, float sigma_sqr,
                           float epsilon, float* hist_raw, void* stream) {
  PH_CHECK_ARG(component && projection1 && projection2 && color_intensities && bin_centers && hist_raw,
               "NULL pointer argument");
  PH_CHECK_ARG(batch >= 0 && npix > 0, "bad shape");
  PH_CHECK_ARG(bins >= 1 && bins <= 1024, "bins must be in [1,1024] (got %d)", bins);
  PH_CHECK_ARG(method == PH_METHOD_INVERSE_QUADRATIC || method == PH_METHOD_RBF, "unknown histogram method %d", method);
  PH_CHECK_ARG(sigma_sqr > 0.f, "sigma_sqr must be positive");
  if (batch == 0) return PH_OK;
  return simt_component_histogram(component, projection1, projection2, color_intensities, batch, npix,
                                  bin_centers, bins, method, sigma_sqr, epsilon, hist_raw,
                                  static_cast<cudaStream_t>(stream));
}

int ph_hist_backward(const float* image, int64_t batch, int64_t npix, int channels,
                     const float* bin_centers, int bins, int method, float sigma_sqr, float epsilon,
                     const float* hist_pred, const float* denom_pred, const float* grad_hist,
                     const float* hist_true, const double* ssum, int64_t global_batch,
                     const float* loss_scale, float* grad_image, void* workspace, size_t workspace_bytes,
                     int impl, void* stream) {
  int rc = check_hist_args(image, batch, npix, channels, bin_centers, bins, method, sigma_sqr);
  if (rc != PH_OK) return rc;
  if ((rc = consume_async_status()) != PH_OK) return rc;
  PH_CHECK_ARG(hist_pred && denom_pred && grad_image, "hist_pred / denom_pred / grad_image must not be NULL");
  PH_CHECK_ARG(grad_hist != nullptr || (hist_true != nullptr && ssum != nullptr && global_batch > 0),
               "either grad_hist or (hist_true, ssum, global_batch>0) must be given");
  PH_CHECK_ARG((impl & ~(PH_IMPL_ENGINE_MASK | PH_IMPL_DEDUP | PH_IMPL_MIRROR)) == 0 && (impl & PH_IMPL_ENGINE_MASK) <= PH_IMPL_TC,
               "bad impl %d", impl);
  PH_CHECK_ARG(channels != 4 || (reinterpret_cast<uintptr_t>(grad_image) & 15) == 0,
               "RGBA gradient pointer must be 16-byte aligned");
  if (batch == 0) return PH_OK;
  const int eng = resolve_impl(impl, npix, bins, method);
  PH_CHECK_ARG(workspace != nullptr && workspace_bytes >= ph_hist_workspace_bytes(batch, npix, bins, eng),
               "workspace too small: %zu < %zu", workspace_bytes, ph_hist_workspace_bytes(batch, npix, bins, eng));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (eng == PH_IMPL_TC) {
    if (!tc_supported(npix, bins, method)) {
      set_error("tensor-core engine does not cover npix=%lld bins=%d method=%d", (long long)npix, bins, method);
      return PH_ERR_UNSUPPORTED;
    }
    return tc_hist_backward(image, batch, npix, channels, bin_centers, bins, method, sigma_sqr, epsilon,
                            hist_pred, denom_pred, grad_hist, hist_true, ssum, global_batch, loss_scale,
                            grad_image, workspace, st);
  }
  return simt_hist_backward(image, batch, npix, channels, bin_centers, bins, method, sigma_sqr, epsilon,
                            hist_pred, denom_pred, grad_hist, hist_true, ssum, global_batch, loss_scale,
                            grad_image, workspace, st);
}

int ph_hellinger_ssum(const float* hist_true, const float* hist_pred, int64_t n, double* ssum, void* stream) {
  PH_CHECK_ARG(hist_true && hist_pred && ssum && n >= 0, "bad argument");
  return launch_hellinger_ssum(hist_true, hist_pred, n, ssum, static_cast<cudaStream_t>(stream));
}

int ph_hellinger_finish(const double* ssum, int64_t global_batch, float* loss, void* stream) {
  PH_CHECK_ARG(ssum && loss && global_batch > 0, "bad argument");
  return launch_hellinger_finish(ssum, global_batch, loss, static_cast<cudaStream_t>(stream));
}

int ph_hellinger_backward(const float* hist_true, const float* hist_pred, int64_t n, const double* ssum,
                          int64_t global_batch, const float* loss_scale, float* grad_true, float* grad_pred,
                          void* stream) {
  PH_CHECK_ARG(hist_true && hist_pred && ssum && n >= 0 && global_batch > 0, "bad argument");
  PH_CHECK_ARG(grad_true || grad_pred, "at least one gradient output is required");
  return launch_hellinger_backward(hist_true, hist_pred, n, ssum, global_batch, loss_scale, grad_true, grad_pred,
                                   static_cast<cudaStream_t>(stream));
}

int ph_mean_abs_or_sq_diff(const float* a, const float* b, int64_t n, int kind, float* out, void* stream) {
  PH_CHECK_ARG(a && b && out && n > 0, "bad argument");
  PH_CHECK_ARG(kind == 1 || kind == 2, "kind must be 1 (L1) or 2 (L2)");
  return launch_diff_reduce(a, b, n, kind, out, static_cast<cudaStream_t>(stream));
}

int ph_extract_palette(const int32_t* image, int64_t batch, int64_t rows, int ordering, const float* shuffle_keys,
                       int32_t* palette, int32_t* ncolors, void* stream) {
  PH_CHECK_ARG(image && palette && ncolors, "NULL pointer argument");
  PH_CHECK_ARG(batch >= 0 && rows > 0, "bad shape");
  PH_CHECK_ARG(ordering >= PH_ORDER_TOP2BOTTOM && ordering <= PH_ORDER_SHUFFLED, "bad ordering %d", ordering);
  PH_CHECK_ARG((reinterpret_cast<uintptr_t>(image) & 15) == 0 && (reinterpret_cast<uintptr_t>(palette) & 15) == 0,
               "image / palette must be 16-byte aligned");
  return launch_extract_palette(image, nullptr, batch, rows, ordering, shuffle_keys, palette, ncolors,
                                static_cast<cudaStream_t>(stream));
}

int ph_rgba_to_indexed(const int32_t* image, int64_t batch, int64_t npix, const int32_t* palette,
                       int64_t palette_batch, int mode, int32_t* indexed, float* one_hot, int depth,
                       void* stream) {
  PH_CHECK_ARG(image && palette && indexed, "NULL pointer argument");
  PH_CHECK_ARG(batch >= 0 && npix >= 0, "bad shape");
  PH_CHECK_ARG(palette_batch == 1 || palette_batch == batch, "palette_batch must be 1 or batch");
  PH_CHECK_ARG(mode == PH_INDEX_EXACT_SUM || mode == PH_INDEX_NEAREST, "bad index mode %d", mode);
  PH_CHECK_ARG(one_hot == nullptr || depth > 0, "one-hot depth must be positive");
  PH_CHECK_ARG((reinterpret_cast<uintptr_t>(image) & 15) == 0 && (reinterpret_cast<uintptr_t>(palette) & 15) == 0,
               "image / palette must be 16-byte aligned");
  PH_CHECK_ARG(one_hot == nullptr || (reinterpret_cast<uintptr_t>(one_hot) & 15) == 0, "one_hot must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // gridDim.y limit: split very large batches
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int64_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
    int rc = launch_rgba_to_indexed(image + b0 * npix * 4, nb, npix,
                                    palette + (palette_batch == 1 ? 0 : b0 * PH_MAX_PALETTE_SIZE * 4),
                                    palette_batch == 1 ? 1 : nb, mode, indexed + b0 * npix,
                                    one_hot ? one_hot + b0 * npix * depth : nullptr, depth, st);
    if (rc != PH_OK) return rc;
  }
  return PH_OK;
}

int ph_u8_to_float_image(const uint8_t* image_u8, int64_t npixels, int blacken, int normalize, float* image,
                         void* stream) {
  PH_CHECK_ARG(image_u8 && image && npixels >= 0, "bad argument");
  PH_CHECK_ARG((reinterpret_cast<uintptr_t>(image_u8) & 3) == 0 && (reinterpret_cast<uintptr_t>(image) & 15) == 0,
               "uint8 image must be 4-byte aligned and the float image 16-byte aligned");
  return launch_u8_to_float_image(image_u8, npixels, blacken, normalize, image, static_cast<cudaStream_t>(stream));
}

int ph_augment_pair(const float* first, const float* second, int64_t batch, int height, int width,
                    const float* hue_delta, const float* translation, const uint8_t* apply, int normalize,
                    float* out_first, float* out_second, void* stream) {
  PH_CHECK_ARG(first && out_first, "NULL pointer argument");
  PH_CHECK_ARG((second == nullptr) == (out_second == nullptr), "second and out_second go together");
  PH_CHECK_ARG(batch >= 0 && height > 0 && width > 0, "bad shape");
  PH_CHECK_ARG((int64_t)height * width < (1ll << 31), "image too large");
  PH_CHECK_ARG(out_first != first && out_second != first && (second == nullptr || (out_first != second && out_second != second)),
               "augmentation cannot run in place (the translation gathers)");
  PH_CHECK_ARG(((reinterpret_cast<uintptr_t>(first) | reinterpret_cast<uintptr_t>(second) |
                 reinterpret_cast<uintptr_t>(out_first) | reinterpret_cast<uintptr_t>(out_second)) & 15) == 0,
               "RGBA float32 images must be 16-byte aligned");
  return launch_augment_pair(first, second, batch, height, width, hue_delta, translation, apply, normalize, out_first,
                             out_second, static_cast<cudaStream_t>(stream));
}

int ph_one_hot(const int32_t* indexed, int64_t n, int depth, float* one_hot, void* stream) {
  PH_CHECK_ARG(indexed && one_hot && n >= 0 && depth > 0, "bad argument");
  PH_CHECK_ARG((reinterpret_cast<uintptr_t>(one_hot) & 15) == 0, "one_hot must be 16-byte aligned");
  return launch_one_hot(indexed, n, depth, one_hot, static_cast<cudaStream_t>(stream));
}

int ph_indexed_to_rgba(const int32_t* indexed, int64_t batch, int64_t npix, const int32_t* palette,
                       int64_t palette_batch, int palette_rows, int channels, int32_t* out, void* stream) {
  PH_CHECK_ARG(indexed && palette && out, "NULL pointer argument");
  PH_CHECK_ARG(batch >= 0 && npix >= 0 && palette_rows > 0 && channels > 0, "bad shape");
  PH_CHECK_ARG(palette_batch == 1 || palette_batch == batch, "palette_batch must be 1 or batch");
  PH_CHECK_ARG(channels != 4 || ((reinterpret_cast<uintptr_t>(palette) & 15) == 0 &&
                                 (reinterpret_cast<uintptr_t>(out) & 15) == 0),
               "RGBA palette / out must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int64_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
    int rc = launch_indexed_to_rgba(indexed + b0 * npix, nb, npix,
                                    palette + (palette_batch == 1 ? 0 : b0 * (int64_t)palette_rows * channels),
                                    palette_batch == 1 ? 1 : nb, palette_rows, channels,
                                    out + b0 * npix * channels, st);
    if (rc != PH_OK) return rc;
  }
  return PH_OK;
}

int ph_argmax_indexed(const float* probabilities, int64_t batch, int64_t npix, int depth, const int32_t* palette,
                      int64_t palette_batch, int palette_rows, int32_t* indexed, int32_t* rgba, void* stream) {
  PH_CHECK_ARG(probabilities && (indexed || rgba), "NULL pointer argument");
  PH_CHECK_ARG(batch >= 0 && npix >= 0 && depth > 0, "bad shape");
  PH_CHECK_ARG((reinterpret_cast<uintptr_t>(probabilities) & 15) == 0, "probabilities must be 16-byte aligned");
  if (rgba) {
    PH_CHECK_ARG(palette != nullptr && palette_rows > 0, "rgba output needs a palette");
    PH_CHECK_ARG(palette_batch == 1 || palette_batch == batch, "palette_batch must be 1 or batch");
    PH_CHECK_ARG((reinterpret_cast<uintptr_t>(palette) & 15) == 0 && (reinterpret_cast<uintptr_t>(rgba) & 15) == 0,
                 "RGBA palette / rgba must be 16-byte aligned");
  }
  return launch_argmax_indexed(probabilities, batch, npix, depth, palette, palette_batch, palette_rows, indexed, rgba,
                               static_cast<cudaStream_t>(stream));
}

int ph_load_indexed_images(const int32_t* source, const int32_t* target, int64_t batch, int64_t npix, int ordering,
                           const float* shuffle_keys, int32_t* source_indexed, int32_t* target_indexed,
                           int32_t* palette, int32_t* ncolors, void* stream) {
  PH_CHECK_ARG(source && target && source_indexed && target_indexed && palette && ncolors, "NULL pointer argument");
  PH_CHECK_ARG(batch >= 0 && npix > 0, "bad shape");
  PH_CHECK_ARG(ordering >= PH_ORDER_TOP2BOTTOM && ordering <= PH_ORDER_SHUFFLED, "bad ordering %d", ordering);
  PH_CHECK_ARG((reinterpret_cast<uintptr_t>(source) & 15) == 0 && (reinterpret_cast<uintptr_t>(target) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(palette) & 15) == 0,
               "source / target / palette must be 16-byte aligned");
  // one launch: the palette is extracted and both images are indexed from the same CTA-resident table
  return launch_load_indexed_fused(source, target, 4, batch, npix, ordering, shuffle_keys, source_indexed, target_indexed,
                                   palette, ncolors, static_cast<cudaStream_t>(stream));
}

int ph_load_indexed_images_u8(const uint8_t* source, const uint8_t* target, int64_t batch, int64_t npix, int ordering,
                              const float* shuffle_keys, int32_t* source_indexed, int32_t* target_indexed,
                              int32_t* palette, int32_t* ncolors, void* stream) {
  PH_CHECK_ARG(source && target && source_indexed && target_indexed && palette && ncolors, "NULL pointer argument");
  PH_CHECK_ARG(batch >= 0 && npix > 0, "bad shape");
  PH_CHECK_ARG(ordering >= PH_ORDER_TOP2BOTTOM && ordering <= PH_ORDER_SHUFFLED, "bad ordering %d", ordering);
  PH_CHECK_ARG((reinterpret_cast<uintptr_t>(source) & 3) == 0 && (reinterpret_cast<uintptr_t>(target) & 3) == 0 &&
                   (reinterpret_cast<uintptr_t>(palette) & 15) == 0,
               "uint8 source / target must be 4-byte aligned and the palette 16-byte aligned");
  return launch_load_indexed_fused(source, target, 1, batch, npix, ordering, shuffle_keys, source_indexed, target_indexed,
                                   palette, ncolors, static_cast<cudaStream_t>(stream));
}

int ph_pixel_map(const float* in, int64_t n, int op, float* out, void* stream) {
  PH_CHECK_ARG(in && out && n >= 0, "bad argument");
  PH_CHECK_ARG(op >= PH_MAP_BLACKEN && op <= PH_MAP_DENORMALIZE, "bad op %d", op);
  PH_CHECK_ARG(op != PH_MAP_BLACKEN || n % 4 == 0, "blacken works on RGBA pixels: n must be a multiple of 4");
  return launch_pixel_map(in, n, op, out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
