// Internal launcher prototypes shared between the translation units of libpalhist.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace ph {

int cached_sm_count();  // SM count of the current device (148 on B200), cached per device
// Device-usable pointer to the current device's sticky status word in mapped host memory (ph_async_status), created
// on first use; NULL if the allocation failed.
int* async_status_word();

// ---- hist_simt.cu ---------------------------------------------------------------------------
int simt_fwd_splits(int64_t batch, int64_t npix, int bins);
size_t simt_workspace_bytes(int64_t batch, int64_t npix, int bins);
int simt_hist_forward(const float* image, int64_t batch, int64_t npix, int channels, const float* dom,
                      int bins, int method, float sigma_sqr, float eps, float* hist, float* denom,
                      void* workspace, cudaStream_t st);
int simt_component_histogram(const float* comp, const float* proj1, const float* proj2,
                             const float* inten, int64_t batch, int64_t npix, const float* dom, int bins,
                             int method, float sigma_sqr, float eps, float* hist_raw, cudaStream_t st);
int launch_bwd_prep(const float* hist_pred, const float* denom, const float* grad_hist,
                    const float* hist_true, const double* ssum, int64_t global_batch, const float* loss_scale,
                    int64_t batch, int bins, int transposed, float* ghat, cudaStream_t st);
int simt_hist_backward(const float* image, int64_t batch, int64_t npix, int channels, const float* dom,
                       int bins, int method, float sigma_sqr, float eps, const float* hist_pred,
                       const float* denom, const float* grad_hist, const float* hist_true,
                       const double* ssum, int64_t global_batch, const float* loss_scale, float* grad_image,
                       void* workspace, cudaStream_t st);
int launch_hellinger_ssum(const float* ht, const float* hp, int64_t n, double* ssum, cudaStream_t st);
int launch_hellinger_finish(const double* ssum, int64_t global_batch, float* loss, cudaStream_t st);
int launch_hellinger_backward(const float* ht, const float* hp, int64_t n, const double* ssum,
                              int64_t global_batch, const float* loss_scale, float* grad_true, float* grad_pred,
                              cudaStream_t st);
int launch_diff_reduce(const float* a, const float* b, int64_t n, int kind, float* out, cudaStream_t st);

// ---- hist_tc.cu (tcgen05 engine) --------------------------------------------------------------
bool tc_supported(int64_t npix, int bins, int method);
size_t tc_workspace_bytes(int64_t batch, int64_t npix, int bins);
size_t tc_bwd_workspace_bytes(int64_t batch, int bins);  // hist_tc_bwd.cu: G^ operand tiles, scales (+ float G^ when bins > 64)
int tc_hist_forward(const float* image, int64_t batch, int64_t npix, int channels, const float* dom,
                    int bins, int method, float sigma_sqr, float eps, float* hist, float* denom,
                    void* workspace, bool dedup, bool mirror, const float* hist_true, double* ssum, cudaStream_t st);
// hist_tc_fwd256.cu: dedicated 256-bin forward (whole 256 x 256 histogram of a channel in one CTA's tensor memory)
size_t tc_fwd256_workspace_bytes(int64_t batch, int64_t npix);
void tc_fwd256_plan(int64_t batch, int64_t npix, int* slices_per_image, int64_t* px_per_slice);  // host-only work plan
int tc_fwd256_forward(const float* image, int64_t batch, int64_t npix, int channels, const float* dom, int method,
                      float sigma_sqr, float eps, const float4* ulist, const int* nunique, int dedup_max,
                      float iy_scale, float* hist, float* denom, void* workspace, cudaStream_t st);
// hist_tc_bwd256.cu: dedicated 256-bin backward (128-pixel tile x whole 256 x 256 G^ of a channel per round)
size_t tc_bwd256_workspace_bytes(int64_t batch);
void tc_bwd256_plan(int64_t batch, int64_t npix, int* items_per_image, int* tiles_per_item);  // host-only work plan
int tc_bwd256_backward(const float* image, int64_t batch, int64_t npix, int channels, const float* dom, int method,
                       float sigma_sqr, float eps, const float* hist_pred, const float* denom, const float* grad_hist,
                       const float* hist_true, const double* ssum, int64_t global_batch, const float* loss_scale,
                       float* grad_image, void* workspace, cudaStream_t st);
int launch_hellinger_ssum_accumulate(const float* ht, const float* hp, int64_t n, double* ssum, cudaStream_t st);
int tc_hist_backward(const float* image, int64_t batch, int64_t npix, int channels, const float* dom,
                     int bins, int method, float sigma_sqr, float eps, const float* hist_pred,
                     const float* denom, const float* grad_hist, const float* hist_true,
                     const double* ssum, int64_t global_batch, const float* loss_scale, float* grad_image,
                     void* workspace, cudaStream_t st);

// ---- palette.cu -------------------------------------------------------------------------------
int launch_extract_palette(const int32_t* image, const int32_t* image2, int64_t batch, int64_t rows, int ordering,
                           const float* shuffle_keys, int32_t* palette, int32_t* ncolors, cudaStream_t st);
int launch_load_indexed_fused(const void* source, const void* target, int elem_bytes, int64_t batch, int64_t npix,
                              int ordering, const float* shuffle_keys, int32_t* source_indexed,
                              int32_t* target_indexed, int32_t* palette, int32_t* ncolors, cudaStream_t st);
int launch_pixel_map(const float* in, int64_t n, int op, float* out, cudaStream_t st);
int launch_rgba_to_indexed(const int32_t* image, int64_t batch, int64_t npix, const int32_t* palette,
                           int64_t palette_batch, int mode, int32_t* indexed, float* one_hot, int depth,
                           cudaStream_t st);
int launch_u8_to_float_image(const uint8_t* src, int64_t npixels, int blacken, int normalize, float* dst,
                             cudaStream_t st);
int launch_one_hot(const int32_t* indexed, int64_t n, int depth, float* one_hot, cudaStream_t st);
int launch_augment_pair(const float* first, const float* second, int64_t batch, int height, int width,
                        const float* hue_delta, const float* translation, const uint8_t* apply, int normalize,
                        float* out_first, float* out_second, cudaStream_t st);
int launch_argmax_indexed(const float* probs, int64_t batch, int64_t npix, int depth, const int32_t* palette,
                          int64_t palette_batch, int palette_rows, int32_t* indexed, int32_t* rgba, cudaStream_t st);
int launch_indexed_to_rgba(const int32_t* indexed, int64_t batch, int64_t npix, const int32_t* palette,
                           int64_t palette_batch, int palette_rows, int channels, int32_t* out,
                           cudaStream_t st);

}  // namespace ph
