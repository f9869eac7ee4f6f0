/*
 * palhist.h — C ABI of the B200-native colour kernels (libpalhist.so).
 *
 * Drop-in boundary for the per-pixel colour path of fegemo/palette-and-histo-gan.  The reference
 * has no FFI of its own: its boundary is a set of Python callables over TensorFlow ops.  Each entry
 * point below names the reference callable (file:line under /root/reference) whose arithmetic it
 * replaces; the Python host in `palette_and_histo_gan_b200/` keeps those callables' signatures and
 * forwards to these functions through ctypes (see INTEGRATION.md for the binding).
 *
 * Conventions
 *   - plain pointers and sizes only; every `const T*` / `T*` is a DEVICE pointer unless the
 *     parameter name ends in `_host`;
 *   - all tensors are caller-owned, dense, row-major (C-contiguous) in the layouts stated;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream); kernels are
 *     enqueued on it and the call returns without synchronising unless stated otherwise;
 *   - return value: PH_OK (0) or a negative PH_ERR_* code; `ph_last_error()` returns a
 *     thread-local message for the last failing call on this host thread;
 *   - re-entrant: scratch memory is passed in by the caller (`workspace`, size from the matching
 *     `*_workspace_bytes`); the only process-wide state is the launch counter and the sticky asynchronous
 *     status word of ph_async_status();
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     PH_ERR_CUDA.
 */
#ifndef PALHIST_H_
#define PALHIST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PH_ABI_VERSION 3

#if defined(__GNUC__)
#define PH_API __attribute__((visibility("default")))
#else
#define PH_API
#endif

enum ph_status {
  PH_OK = 0,
  PH_ERR_INVALID = -1,     /* bad argument (shape, enum, NULL pointer, workspace too small) */
  PH_ERR_CUDA = -2,        /* CUDA runtime / launch failure (message has cudaGetErrorString) */
  PH_ERR_UNSUPPORTED = -3, /* valid request this build cannot serve (e.g. tensor-core path, bins not multiple of 64) */
};

/* histogram.py:22-27 — the two bin kernels the reference implements. */
enum ph_method { PH_METHOD_INVERSE_QUADRATIC = 0, PH_METHOD_RBF = 1 };

/* io_utils.py:44-58 — palette orderings.  PH_ORDER_SHUFFLED is the reference's else-branch
 * (`tf.random.shuffle(colors)`, :56-58): the colours in first-occurrence order are permuted by ranking
 * caller-provided independent uniform keys (`shuffle_keys`, (batch,256) float32 on the device) — a uniformly
 * random permutation, reproducible from the caller's seed; TensorFlow's own random stream is not reproduced. */
enum ph_ordering { PH_ORDER_TOP2BOTTOM = 0, PH_ORDER_BOTTOM2TOP = 1, PH_ORDER_GRAYNESS = 2, PH_ORDER_SHUFFLED = 3 };

/* rgba_to_indexed flavour: the reference's exact-match scatter-add (io_utils.py:84-91) or
 * nearest colour (first arg-min of squared RGBA distance; equal to the former whenever every
 * pixel colour occurs exactly once in the palette). */
enum ph_index_mode { PH_INDEX_EXACT_SUM = 0, PH_INDEX_NEAREST = 1 };

/* Contraction engine for the histogram GEMMs. */
enum ph_impl {
  PH_IMPL_AUTO = 0, /* tensor cores when the bin count is a multiple of 64 (64 ... 1024), else SIMT */
  PH_IMPL_SIMT = 1, /* fp32 CUDA-core contraction (any bin count) */
  PH_IMPL_TC = 2,   /* tcgen05 kind::f16 contraction with fp16 hi+lo operand split (fp32-accurate), accumulators in TMEM;
                       more than 64 bins are handled as 64 x 64 blocks */
  PH_IMPL_ENGINE_MASK = 3,
  /* flag, OR-ed into impl for ph_hist_forward: first reduce every image to its unique colours with
   * multiplicities and contract those (exact: the histogram is a sum over pixels of a function of the
   * colour).  Pays off for palette images such as the reference's real sprites (10-54 colours); an
   * image with more than 512 colours is contracted densely.  Ignored by the CUDA-core engine. */
  PH_IMPL_DEDUP = 8,
  /* flag, OR-ed into impl for ph_hist_forward / ph_hist_forward_ssum: the caller asserts that the bin centres are
   * antisymmetric, |c[j] + c[bins-1-j]| <= 2e-5 sigma for every j (tf.linspace(-3, 3, 64), histogram.py:55, is: 3.6e-7).
   * Dense 64-bin batches of images with >= 1024 pixels on the tensor-core engine then run the mirrored-tile forward: the weight vectors of +x and -x
   * are bin-reversed copies of each other around the midpoint centres (c[j] - c[63-j]) / 2, three vectors per pixel
   * instead of six (adds <= 2e-6 to the histogram error at 64 x 64 pixels).  Ignored for every other
   * configuration (other bin counts, PH_IMPL_DEDUP, the CUDA-core engine, the backward).  A launch whose centres break
   * the assertion sets PH_ASYNC_MIRROR. */
  PH_IMPL_MIRROR = 16,
};

/* per-image status written by ph_extract_palette into ncolors[]: >=0 colour count (may exceed 256 =
 * overflow, palette undefined — the reference raises there, io_utils.py:62); PH_PALETTE_BAD_VALUE
 * when a channel value lies outside [0,255]. */
#define PH_PALETTE_BAD_VALUE (-1)
#define PH_MAX_PALETTE_SIZE 256 /* configuration.py:31 */

PH_API int ph_abi_version(void);
PH_API const char* ph_last_error(void);
/* Number of kernel launches issued through this library by this process (all threads) since the
 * last reset (bench.py's `gpu_launches`). */
PH_API int64_t ph_launch_count(void);
PH_API void ph_reset_launch_count(void);
/* Sticky asynchronous status of the kernels launched by this process on `device` (-1: the current device): bit 0 =
 * a histogram forward on the tensor-core engine met a pixel whose intensity-weighted operand does not fit fp16
 * (an image far outside [-1,1]: the reference, histogram.py:58, accepts any float; results of that launch are
 * inf / NaN — re-run it with PH_IMPL_SIMT).  The word lives in mapped host memory, so reading it never
 * synchronises; it is as current as the last completed kernel.  clear != 0 resets it.  Every ph_hist_* call also
 * checks it on entry and fails with PH_ERR_UNSUPPORTED (clearing it), like CUDA's own asynchronous errors. */
#define PH_ASYNC_RANGE 1
/* bit 1 = a forward launched with PH_IMPL_MIRROR found bin centres that are not antisymmetric (its results are off by
 * the asymmetry / sigma; re-run without the flag). */
#define PH_ASYNC_MIRROR 2
PH_API int ph_async_status(int device, int clear);
/* sm_count / compute capability of `device`; PH_ERR_CUDA without a usable GPU. */
PH_API int ph_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* Host-only query (no device needed): how the dedicated 256-bin kernels cut a batch into work
 * items on this device's SM count.  plan4 = { forward: pixel slices per image, pixels per slice (a multiple of the
 * 512-pixel accumulation chain); backward: items (tile ranges) per image, 128-pixel tiles per item }.  The slices /
 * items cover every pixel exactly once and none is empty (tests/test_abi.py). */
PH_API int ph_hist256_plan(int64_t batch, int64_t npix, int64_t* plan4);

/* ------------------------------------------------------------------------------------------
 * RGB-uv histogram  (histogram.py:36-81 `calculate_rgbuv_histogram`)
 * ------------------------------------------------------------------------------------------
 * image        (batch, npix, channels) float32 in [-1,1], channels = 3 or 4 (alpha ignored, :61)
 * bin_centers  (bins) float32 — the `tf.linspace(-3,3,size)` tensor of :55 passed as data
 * hist         (batch, bins, bins, 3) float32, each image sums to 1 (:78-79)
 * denom        (batch) float32 — the per-image normaliser of :78 (needed by the backward)
 */
PH_API size_t ph_hist_workspace_bytes(int64_t batch, int64_t npix, int bins, int impl);

PH_API int ph_hist_forward(const float* image, int64_t batch, int64_t npix, int channels,
                    const float* bin_centers, int bins, int method, float sigma_sqr, float epsilon,
                    float* hist, float* denom, void* workspace, size_t workspace_bytes, int impl,
                    void* stream);

/* The generator-loss call site (pix2pix_model.py:243-245) as one launch per image set: histogram of the `fake`
 * images as above and, while each normalised histogram is still on chip, this batch's share of the Hellinger sum
 * of squares  sum (sqrt(hist) - sqrt(hist_true))^2  (histogram.py:88) against the histograms of the real images,
 * added into the device double *ssum (zeroed first unless accumulate != 0, so a batch can be fed in chunks). */
PH_API int ph_hist_forward_ssum(const float* image, int64_t batch, int64_t npix, int channels, const float* bin_centers,
                         int bins, int method, float sigma_sqr, float epsilon, float* hist, float* denom,
                         const float* hist_true, double* ssum, int accumulate, void* workspace,
                         size_t workspace_bytes, int impl, void* stream);

/* histogram.py:5-32 `calculate_component_histogram`: un-normalised (batch,bins,bins) histogram of one
 * component against two projections, intensities (batch,npix) given. */
PH_API int ph_component_histogram(const float* component, const float* projection1, const float* projection2,
                           const float* color_intensities, int64_t batch, int64_t npix,
                           const float* bin_centers, int bins, int method, float sigma_sqr,
                           float epsilon, float* hist_raw, void* stream);

/* Backward of ph_hist_forward (TF autodiff in the reference, pix2pix_model.py:78).  Exactly one of
 * the two upstream forms is used:
 *   (a) grad_hist != NULL: dL/dhist, (batch,bins,bins,3) float32;
 *   (b) grad_hist == NULL: the Hellinger loss of histogram.py:84-89 is differentiated in the
 *       prologue from hist_true, *ssum (device double, whole-batch sum of squares — after the
 *       all-reduce when the batch is sharded), global_batch and the upstream scalar *loss_scale (device float; NULL = 1).
 * grad_image (batch, npix, channels) float32; the alpha channel (if any) is written as 0. */
PH_API int ph_hist_backward(const float* image, int64_t batch, int64_t npix, int channels,
                     const float* bin_centers, int bins, int method, float sigma_sqr, float epsilon,
                     const float* hist_pred, const float* denom_pred, const float* grad_hist,
                     const float* hist_true, const double* ssum, int64_t global_batch,
                     const float* loss_scale, float* grad_image, void* workspace, size_t workspace_bytes,
                     int impl, void* stream);

/* ------------------------------------------------------------------------------------------
 * Hellinger / L1 / L2 histogram losses  (histogram.py:84-97)
 * ------------------------------------------------------------------------------------------ */
/* *ssum (device double) = sum over n elements of (sqrt(pred)-sqrt(true))^2; overwritten. */
PH_API int ph_hellinger_ssum(const float* hist_true, const float* hist_pred, int64_t n, double* ssum,
                      void* stream);
/* *loss (device float) = (1/sqrt 2) * sqrt(*ssum) / global_batch. */
PH_API int ph_hellinger_finish(const double* ssum, int64_t global_batch, float* loss, void* stream);
/* Backward of the Hellinger loss: d loss / d hist for either argument (NULL = not wanted), n elements,
 * scaled by the upstream scalar *loss_scale (device float; NULL = 1); *ssum is the whole-batch sum
 * of squares. */
PH_API int ph_hellinger_backward(const float* hist_true, const float* hist_pred, int64_t n, const double* ssum,
                          int64_t global_batch, const float* loss_scale, float* grad_true, float* grad_pred,
                          void* stream);
/* *out (device float) = mean |a-b| (kind 1) or mean (a-b)^2 (kind 2) over n elements. */
PH_API int ph_mean_abs_or_sq_diff(const float* a, const float* b, int64_t n, int kind, float* out,
                           void* stream);

/* ------------------------------------------------------------------------------------------
 * Palette helpers  (io_utils.py:25-103, pix2pix_model.py:300-301, dataset_utils.py:131-151)
 * ------------------------------------------------------------------------------------------ */
/* io_utils.py:25-65 `extract_palette`, batched.  image (batch, rows, 4) int32 with values in
 * [0,255]; rows are the pixels in reshape(-1,4) order.  palette (batch,256,4) int32 padded with
 * INVALID_INDEX_COLOR (configuration.py:32); ncolors (batch) int32, see PH_PALETTE_BAD_VALUE.
 * shuffle_keys: (batch,256) float32, required for PH_ORDER_SHUFFLED, ignored (may be NULL) otherwise. */
PH_API int ph_extract_palette(const int32_t* image, int64_t batch, int64_t rows, int ordering,
                       const float* shuffle_keys, int32_t* palette, int32_t* ncolors, void* stream);

/* io_utils.py:78-93 `rgba_to_indexed`, batched, optionally fused with the one-hot of
 * pix2pix_model.py:300-301.  image (batch,npix,4) int32; palette (palette_batch,256,4) int32 with
 * palette_batch == batch or 1 (shared); indexed (batch,npix) int32; one_hot NULL or
 * (batch,npix,depth) float32 (index outside [0,depth) -> all-zero row, TF semantics). */
PH_API int ph_rgba_to_indexed(const int32_t* image, int64_t batch, int64_t npix, const int32_t* palette,
                       int64_t palette_batch, int mode, int32_t* indexed, float* one_hot, int depth,
                       void* stream);

/* dataset_utils.py:66-77 `load_image` after PNG decode: uint8 RGBA (npixels,4) -> float32 (npixels,4) with
 * blacken_transparent_pixels (:11-20, applied when blacken != 0) and normalize (:39-48, x/127.5-1, when
 * normalize != 0).  Lets a host caller upload sprites as uint8 (4 B/pixel instead of 16). */
PH_API int ph_u8_to_float_image(const uint8_t* image_u8, int64_t npixels, int blacken, int normalize, float* image,
                         void* stream);

/* dataset_utils.py:80-102 `augment_two` for a batch of RGBA float32 image pairs (batch,height,width,4), with the
 * probability gate of :109-120 and, optionally, the `normalize` of :39-48 that follows it in `load_rgba_ds`
 * (:220-225) fused in:  hue rotation of channels 0..2 of both images by hue_delta[b] (tf.image.adjust_hue; the
 * caller draws delta in [-0.5, 0.5), :82), then one shared translation by translation[b] = (dx, dy) pixels
 * (keras RandomTranslation((-0.15, 0.075), 0.125, "constant", "nearest"), :89: out[y,x] = in[round(y-dy),
 * round(x-dx)] or 0 outside).  hue_delta (batch) and translation (batch,2) are DEVICE float32 arrays, either
 * may be NULL (step skipped); apply (batch) device uint8 or NULL: images with apply[b] == 0 pass through
 * unchanged; second / out_second may both be NULL (single image, :80-84 or :87-92 alone).  Not in place. */
PH_API int ph_augment_pair(const float* first, const float* second, int64_t batch, int height, int width,
                    const float* hue_delta, const float* translation, const uint8_t* apply, int normalize,
                    float* out_first, float* out_second, void* stream);

/* pix2pix_model.py:300-301 `tf.one_hot(idx, depth)`: indexed (n) int32 -> one_hot (n,depth) float32. */
PH_API int ph_one_hot(const int32_t* indexed, int64_t n, int depth, float* one_hot, void* stream);

/* io_utils.py:96-103 `indexed_to_rgba`: out[b,n,:] = palette[b or 0, indexed[b,n], :]; an index
 * outside [0,palette_rows) is a PH_ERR_INVALID-free no-op row of zeros (TF's GPU gather semantics). */
PH_API int ph_indexed_to_rgba(const int32_t* indexed, int64_t batch, int64_t npix, const int32_t* palette,
                       int64_t palette_batch, int palette_rows, int channels, int32_t* out,
                       void* stream);

/* pix2pix_model.py:283-287 (`generate`: argmax over the softmax channels, int32, first maximum) fused with
 * io_utils.py:96-103 `indexed_to_rgba` (:356, :446-447): probabilities (batch,npix,depth) float32 ->
 * indexed (batch,npix) int32 and/or rgba (batch,npix,4) int32 = palette[b or 0, argmax, :]; either output may
 * be NULL.  An arg-max >= palette_rows gives a row of zeros like ph_indexed_to_rgba. */
PH_API int ph_argmax_indexed(const float* probabilities, int64_t batch, int64_t npix, int depth, const int32_t* palette,
                      int64_t palette_batch, int palette_rows, int32_t* indexed, int32_t* rgba, void* stream);

/* dataset_utils.py:138-151 glue, batched: shared palette of source||target (rows interleaved
 * src px0, tgt px0, src px1, ... as the channel-axis concat + reshape(-1,4) produces), then both
 * index images.  source/target (batch,npix,4) int32. */
PH_API int ph_load_indexed_images(const int32_t* source, const int32_t* target, int64_t batch, int64_t npix,
                           int ordering, const float* shuffle_keys, int32_t* source_indexed,
                           int32_t* target_indexed, int32_t* palette, int32_t* ncolors, void* stream);
/* Same with source/target as the decoded PNG's uint8 RGBA (batch,npix,4) on the device (dataset_utils.py:66-70
 * before the int32 cast of :140-141): 4 B per pixel read instead of 16, identical outputs. */
PH_API int ph_load_indexed_images_u8(const uint8_t* source, const uint8_t* target, int64_t batch, int64_t npix,
                              int ordering, const float* shuffle_keys, int32_t* source_indexed,
                              int32_t* target_indexed, int32_t* palette, int32_t* ncolors, void* stream);

/* The loader's pixel helpers on device tensors (dataset_utils.py:11-20, :39-60), n elements of float32:
 *   PH_MAP_BLACKEN      RGBA groups of four: alpha == 0 -> the whole pixel becomes 0 (n a multiple of 4);
 *   PH_MAP_NORMALIZE    x / 127.5 - 1  (true division and a subtraction, the reference's two roundings);
 *   PH_MAP_DENORMALIZE  (x + 1) * 127.5.
 * In place (out == in) is allowed. */
enum ph_map_op { PH_MAP_BLACKEN = 0, PH_MAP_NORMALIZE = 1, PH_MAP_DENORMALIZE = 2 };
PH_API int ph_pixel_map(const float* in, int64_t n, int op, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sharded batches on one NVLink / NVSwitch box: the one exchange of the path — the sum over ranks of the Hellinger
 * sum of squares (histogram.py:88-89 takes one sqrt over the WHOLE batch) — as a single 32-thread kernel over peer
 * memory instead of a library collective.  Every rank owns a small mailbox in its HBM; the kernel stores this
 * rank's value into the mailbox of every peer (P2P stores over NVLink, release at system scope), waits until the
 * values of all ranks for this round have landed in its own mailbox (acquire polls of local memory) and adds them
 * in rank order, so all ranks obtain the same bits.  ~3 us against ~30 us for an 8-byte NCCL all-reduce, no host
 * round trip, no SM-residency requirement beyond one warp.
 *   ph_comm_create   allocates this rank's mailbox on `device`;
 *   ph_comm_export   its 64-byte CUDA IPC handle (exchange the handles of all ranks by any means, e.g.
 *                    torch.distributed.all_gather_object);
 *   ph_comm_connect  maps the peers' mailboxes (handles_host: world x 64 bytes, rank order; own entry ignored);
 *   ph_comm_allreduce_sum_f64  *value (device double, `count` <= 2 of them) <- sum over ranks, in place, enqueued
 *                    on `stream`; a collective: every rank must call it the same number of times.
 * A rank that waits longer than ~10 s for a peer aborts its kernel (trap -> launch error) instead of hanging.
 * ------------------------------------------------------------------------------------------ */
typedef struct ph_comm ph_comm;
#define PH_COMM_HANDLE_BYTES 64
#define PH_COMM_MAX_WORLD 16
PH_API int ph_comm_create(int device, int rank, int world, ph_comm** comm);
PH_API int ph_comm_export(ph_comm* comm, void* handle_host);
PH_API int ph_comm_connect(ph_comm* comm, const void* handles_host);
PH_API int ph_comm_allreduce_sum_f64(ph_comm* comm, double* value, int count, void* stream);
PH_API void ph_comm_destroy(ph_comm* comm);

/* ------------------------------------------------------------------------------------------
 * Host-buffer convenience (the call a CPU-side caller such as a tf.data worker binds):
 * pinned or pageable HOST pointers in and out; device staging, chunked H2D/compute/D2H overlap and
 * one final stream synchronise happen inside.  `ctx` from ph_host_ctx_create (one per host thread).
 * ------------------------------------------------------------------------------------------ */
typedef struct ph_host_ctx ph_host_ctx;
PH_API int ph_host_ctx_create(int device, ph_host_ctx** ctx);
PH_API void ph_host_ctx_destroy(ph_host_ctx* ctx);

/* pix2pix_model.py:243-245 + :78 in one call: loss and d loss / d fake for host images.
 * real_host / fake_host / grad_fake_host (batch,npix,channels) float32; loss_host 1 float. */
PH_API int ph_host_hist_loss(ph_host_ctx* ctx, const float* real_host, const float* fake_host, int64_t batch,
                      int64_t npix, int channels, const float* bin_centers_host, int bins, int method,
                      float sigma_sqr, float epsilon, int impl, float* loss_host,
                      float* grad_fake_host);

/* The same in two phases for a batch sharded over processes: `begin` uploads, runs both forward
 * passes and returns this shard's sum of squares (host double); the caller all-reduces it; `finish`
 * takes the whole-batch sum and batch size, returns the loss (host float) and this shard's gradient:
 * downloaded into grad_fake_host and/or left in the caller's DEVICE buffer grad_fake_device (either
 * may be NULL; the usual consumer, the generator's backward pass, lives on the device). */
PH_API int ph_host_hist_begin(ph_host_ctx* ctx, const float* real_host, const float* fake_host, int64_t batch,
                       int64_t npix, int channels, const float* bin_centers_host, int bins, int method,
                       float sigma_sqr, float epsilon, int impl, double* ssum_local_host);
/* Same with the real images as uint8 RGBA sprites straight from the decoder (batch,npix,4): they are
 * blackened + normalised on the device (dataset_utils.py:66-77); fake stays float32 RGBA. */
PH_API int ph_host_hist_begin_u8real(ph_host_ctx* ctx, const uint8_t* real_u8_host, const float* fake_host,
                              int64_t batch, int64_t npix, const float* bin_centers_host, int bins, int method,
                              float sigma_sqr, float epsilon, int impl, double* ssum_local_host);
PH_API int ph_host_hist_finish(ph_host_ctx* ctx, double ssum_global, int64_t global_batch, float* loss_host,
                        float* grad_fake_host, float* grad_fake_device);
/* `finish` for ranks connected by a ph_comm: the shard's sum of squares left on the device by `begin` is
 * summed over the ranks by ph_comm_allreduce_sum_f64 on the context's stream — no host round trip between the
 * phases (the host value `begin` returned is not needed). */
PH_API int ph_host_hist_finish_comm(ph_host_ctx* ctx, ph_comm* comm, int64_t global_batch, float* loss_host,
                             float* grad_fake_host, float* grad_fake_device);
/* Both phases in one call for a shard of a batch spread over the ranks of `comm`: upload, forwards, unit-scale
 * backward, sum of the shards' sums of squares over peer memory, loss, gradient — one host synchronisation, at the
 * end.  real_host: float32 (batch,npix,channels) or, with real_is_u8 != 0, uint8 RGBA sprites. */
PH_API int ph_host_hist_loss_sharded(ph_host_ctx* ctx, ph_comm* comm, const void* real_host, int real_is_u8,
                              const float* fake_host, int64_t batch, int64_t npix, int channels,
                              const float* bin_centers_host, int bins, int method, float sigma_sqr, float epsilon,
                              int impl, int64_t global_batch, float* loss_host, float* grad_fake_host,
                              float* grad_fake_device);

/* dataset_utils.py:138-151 for host images (+ optional one-hot of the target indices). */
PH_API int ph_host_load_indexed_images(ph_host_ctx* ctx, const int32_t* source_host, const int32_t* target_host,
                                int64_t batch, int64_t npix, int ordering, const float* shuffle_keys_host,
                                int32_t* source_indexed_host, int32_t* target_indexed_host, int32_t* palette_host,
                                int32_t* ncolors_host, float* target_one_hot_host);

/* Same with the images as the decoded PNG's uint8 RGBA (dataset_utils.py:66-70 `decode_png` before the casts of
 * :72 and :140-141): a quarter of the upload, read as uint8 by the kernel, outputs identical.
 * shuffle_keys_host: (batch,256) float32 for PH_ORDER_SHUFFLED, else NULL. */
PH_API int ph_host_load_indexed_images_u8(ph_host_ctx* ctx, const uint8_t* source_host, const uint8_t* target_host,
                                   int64_t batch, int64_t npix, int ordering, const float* shuffle_keys_host,
                                   int32_t* source_indexed_host, int32_t* target_indexed_host, int32_t* palette_host,
                                   int32_t* ncolors_host, float* target_one_hot_host);

#ifdef __cplusplus
}
#endif
#endif /* PALHIST_H_ */
