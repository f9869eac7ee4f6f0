// Palette kernels: unique-colour extraction, colour -> index, one-hot, index -> colour.
// Integer/byte work bound by HBM bandwidth: 128-bit loads/stores, one CTA-resident hash table per
// image, warp-level match/ballot de-duplication in front of the shared-memory atomics.
//
// Reference semantics: io_utils.py:25-65 (extract_palette), :78-93 (rgba_to_indexed), :96-103
// (indexed_to_rgba), pix2pix_model.py:300-301 (one-hot), dataset_utils.py:138-151 (call site).
#include "common.cuh"
#include "hist_internal.cuh"

namespace ph {

constexpr unsigned long long SLOT_EMPTY = ~0ull;

__device__ __forceinline__ bool in_byte_range(const int4& c) {
  return ((unsigned)c.x | (unsigned)c.y | (unsigned)c.z | (unsigned)c.w) < 256u;
}
__device__ __forceinline__ unsigned pack_rgba(const int4& c) {
  return (unsigned)c.x | ((unsigned)c.y << 8) | ((unsigned)c.z << 16) | ((unsigned)c.w << 24);
}
__device__ __forceinline__ int4 unpack_rgba(unsigned k) {
  return make_int4((int)(k & 255u), (int)((k >> 8) & 255u), (int)((k >> 16) & 255u), (int)(k >> 24));
}
// x / 127.5f, correctly rounded like TensorFlow's true division (dataset_utils.py:39-48), without the ~10-instruction
// IEEE division: q = x y, r = fma(-q, 127.5, x) (exact), q' = fma(r, y, q) with y = RN(1 / 127.5) is the correctly
// rounded quotient (Markstein); tools/div127_5_check.c compares it with x / 127.5f for ALL 2^32 bit patterns: identical
// for 1e-30 <= |x| < 1e38, so zeros (sign kept) are handled apart and anything else takes the IEEE division.
__device__ __forceinline__ float div_127_5(float x) {
  const float y = 1.0f / 127.5f;  // constant-folded, correctly rounded
  const float q = __fmul_rn(x, y);
  const float r = __fmaf_rn(-q, 127.5f, x);
  const float q2 = __fmaf_rn(r, y, q);
  const float ax = fabsf(x);
  if (ax >= 1e-30f && ax < 1e38f) return q2;
  if (x == 0.f) return x;
  return __fdiv_rn(x, 127.5f);
}
__device__ __forceinline__ float normalize_px(float x) { return __fsub_rn(div_127_5(x), 1.0f); }

template <int BITS>
__device__ __forceinline__ unsigned hash_slot(unsigned key) {
  return (key * 2654435761u) >> (32 - BITS);
}

// =============================================================================================
// extract_palette: one CTA (512 threads) per image.
//   1. every row (pixel) is packed to a 32-bit key; lanes holding the same key elect the lane with
//      the earliest row (warp match) and only that lane touches the hash table;
//   2. the table keeps (key, earliest row) per colour  -> first-occurrence order of
//      UniqueWithCountsV2 (io_utils.py:46-57);
//   3. entries are ranked by earliest row, and for "grayness" re-ranked by the float32 key
//      ((r*0.2989+g*0.5870)+b*0.1140)+a*0 with ties broken by first occurrence = stable argsort
//      (io_utils.py:51-55); "shuffled" (io_utils.py:56-58) re-ranks by caller-provided random keys the same way;
//   4. rows n..255 are INVALID_INDEX_COLOR (io_utils.py:61-63, configuration.py:32).
// The pass is a chain load -> match -> hash per row (~100 instructions per pixel), i.e. bound by latency and issue
// rate, not by bytes: every thread issues four independent loads before it touches the table, and when the image
// fits (rows <= 512 x 16 = 8192: a 64 x 64 source||target pair exactly) the packed keys stay in registers, so the
// index pass of the fused variant reads no pixel a second time.  Measured (bench.py, tools/gpu_r2_d.sh): 4096 pairs
// 186 us = 3.6 TB/s of the 20 B per pixel that move (55 % of the HBM copy rate); a cfgB batch of 256 pairs is one
// CTA per pair = 28 warps per SM and takes 24.6 us, of which the dependent chain of one CTA is ~16 us.  Splitting a
// pair over a cluster of 2 or 4 CTAs (tables merged into rank 0's through distributed shared memory, finished
// table copied back for the index pass) was built and measured SLOWER — 28.7 us and 36.9 us: three cluster barriers
// and the merge cost more than the halved row loop saves — and removed again.
// =============================================================================================
constexpr int PAL_THREADS = 512;
constexpr int PAL_HASH_BITS = 11;
constexpr int PAL_HASH_SIZE = 1 << PAL_HASH_BITS;
constexpr int PAL_MAX = PH_MAX_PALETTE_SIZE;
constexpr int PAL_INFLIGHT = 4;                           // loads in flight per thread
constexpr int PAL_KEEP = 16;                              // keys a thread keeps in registers (cached variant)

// U8: pixels are the decoded PNG's uint8 RGBA (4 B, already the packed key); otherwise int32 RGBA (16 B).
template <bool U8>
__device__ __forceinline__ unsigned load_pixel_key(const void* src0, const void* src1, int r, bool& bad) {
  // src1 != nullptr: rows interleave source/target pixels (dataset_utils.py:142-145)
  const void* base = (src1 != nullptr && (r & 1)) ? src1 : src0;
  const int i = src1 != nullptr ? (r >> 1) : r;
  if (U8) return __ldg(static_cast<const unsigned*>(base) + i);
  const int4 c = __ldg(static_cast<const int4*>(base) + i);
  if (!in_byte_range(c)) bad = true;
  return pack_rgba(c);
}

// (key, position) into an open-addressing table that keeps the smallest position per key
__device__ __forceinline__ void palette_table_insert(unsigned long long* table, int* count, unsigned key, unsigned pos) {
  const unsigned long long word = ((unsigned long long)key << 32) | pos;
  unsigned h = hash_slot<PAL_HASH_BITS>(key);
  for (int probe = 0; probe < PAL_HASH_SIZE; ++probe) {
    unsigned long long cur = *(volatile unsigned long long*)&table[h];
    if (cur == SLOT_EMPTY) {
      cur = atomicCAS(&table[h], SLOT_EMPTY, word);
      if (cur == SLOT_EMPTY) { atomicAdd(count, 1); return; }
    }
    if ((unsigned)(cur >> 32) == key) {
      if ((unsigned)cur > pos) atomicMin(&table[h], word);
      return;
    }
    h = (h + 1) & (PAL_HASH_SIZE - 1);
  }
}

// FUSED_INDEX: also index both images of the pair from the same CTA-resident table
// (dataset_utils.py:148-149 in the same launch): after the colours are ranked, each table slot is
// rewritten to (key, final palette index) and every pixel is looked up with one probe.
// CACHED: rows <= PAL_THREADS * PAL_KEEP = 8192, the keys of the thread's rows stay in registers.
template <bool FUSED_INDEX, bool U8, bool CACHED>
__global__ void __launch_bounds__(PAL_THREADS, 3) extract_palette_kernel(
    const void* __restrict__ image, const void* __restrict__ image2, int64_t rows64, int ordering,
    const float* __restrict__ shuffle_keys, int4* __restrict__ palette, int* __restrict__ ncolors,
    int* __restrict__ indexed, int* __restrict__ indexed2) {
  __shared__ unsigned long long table[PAL_HASH_SIZE];
  __shared__ unsigned long long entries[PAL_MAX];
  __shared__ unsigned first_order[PAL_MAX];  // keys in first-occurrence order
  __shared__ float gray[PAL_MAX];            // secondary sort key in first-occurrence order
  __shared__ int final_rank[PAL_MAX];        // palette row of the colour with first-occurrence rank i
  __shared__ int s_count, s_bad, s_n, s_fail;

  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const int rows = (int)rows64;  // < 2^31 (checked by the launcher): 32-bit row arithmetic throughout
  for (int i = tid; i < PAL_HASH_SIZE; i += PAL_THREADS) table[i] = SLOT_EMPTY;
  if (tid == 0) { s_count = 0; s_bad = 0; s_n = 0; s_fail = 0; }
  __syncthreads();

  const int per_image = image2 ? rows / 2 : rows;
  const size_t px_bytes = U8 ? 4 : 16;
  const void* src0 = static_cast<const char*>(image) + (size_t)b * per_image * px_bytes;
  const void* src1 = image2 ? static_cast<const char*>(image2) + (size_t)b * per_image * px_bytes : nullptr;
  int* const idx0 = FUSED_INDEX ? indexed + b * per_image : nullptr;
  int* const idx1 = FUSED_INDEX ? indexed2 + b * per_image : nullptr;
  const bool reversed = ordering == PH_ORDER_BOTTOM2TOP;
  auto row_at = [&](int j) { return j * PAL_THREADS + tid; };  // row handled by this thread in its slot j

  unsigned kept[CACHED ? PAL_KEEP : 1];
  bool bad = false;
  // one batch: PAL_INFLIGHT independent loads per thread, then the match / hash step of each row
  auto insert_batch = [&](const int bt, unsigned* keep) {
    unsigned key[PAL_INFLIGHT];
#pragma unroll
    for (int k = 0; k < PAL_INFLIGHT; ++k) {
      const int r = row_at(bt * PAL_INFLIGHT + k);
      key[k] = r < rows ? load_pixel_key<U8>(src0, src1, r, bad) : 0u;
      if (keep != nullptr) keep[k] = key[k];
    }
#pragma unroll
    for (int k = 0; k < PAL_INFLIGHT; ++k) {
      const int r = row_at(bt * PAL_INFLIGHT + k);
      const bool active = r < rows;
      // warp de-duplication: among lanes with the same colour keep the one with the earliest row
      const unsigned amask = __ballot_sync(0xffffffffu, active);
      if (!active) continue;
      const unsigned peers = __match_any_sync(amask, key[k]);
      const int leader = reversed ? 31 - __clz(peers) : __ffs(peers) - 1;
      if (lane != leader) continue;
      if (*(volatile int*)&s_count > PAL_MAX) continue;  // already overflowed: result is "too many"
      palette_table_insert(table, &s_count, key[k], (unsigned)(reversed ? rows - 1 - r : r));
    }
  };
  if (CACHED) {
#pragma unroll
    for (int bt = 0; bt < PAL_KEEP / PAL_INFLIGHT; ++bt) insert_batch(bt, &kept[bt * PAL_INFLIGHT]);  // rows beyond the image: inactive
  } else {
    const int nbatch = (rows + PAL_THREADS * PAL_INFLIGHT - 1) / (PAL_THREADS * PAL_INFLIGHT);
#pragma unroll 1
    for (int bt = 0; bt < nbatch; ++bt) insert_batch(bt, nullptr);
  }
  if (bad) s_bad = 1;
  __syncthreads();

  {
    const int count = s_count;
    const bool fail = s_bad || count > PAL_MAX;
    const int4 filler = make_int4(255, 0, 220, 255);
    int4* out = palette + b * PAL_MAX;
    if (fail) {
      if (tid == 0) { ncolors[b] = s_bad ? PH_PALETTE_BAD_VALUE : count; s_fail = 1; }
      for (int k = tid; k < PAL_MAX; k += PAL_THREADS) out[k] = filler;
    } else {
      // compact occupied slots
      for (int i = tid; i < PAL_HASH_SIZE; i += PAL_THREADS) {
        const unsigned long long w = table[i];
        if (w != SLOT_EMPTY) entries[atomicAdd(&s_n, 1)] = w;
      }
      __syncthreads();
      const int n = s_n;  // == count
      // rank by earliest row -> first-occurrence order
      if (tid < n) {
        const unsigned long long me = entries[tid];
        const unsigned mypos = (unsigned)me;
        int rank = 0;
        for (int e = 0; e < n; ++e) rank += ((unsigned)entries[e] < mypos) ? 1 : 0;
        const unsigned key = (unsigned)(me >> 32);
        first_order[rank] = key;
        float g = 0.f;
        if (ordering == PH_ORDER_GRAYNESS) {
          const int4 c = unpack_rgba(key);
          // float32, non-fused, left to right: the (n,4)x(4,1) product of io_utils.py:51-52
          g = __fmul_rn((float)c.x, 0.2989f);
          g = __fadd_rn(g, __fmul_rn((float)c.y, 0.5870f));
          g = __fadd_rn(g, __fmul_rn((float)c.z, 0.1140f));
          g = __fadd_rn(g, __fmul_rn((float)c.w, 0.0f));
        } else if (ordering == PH_ORDER_SHUFFLED) {
          // tf.random.shuffle(colors): rank by independent uniform keys = a uniformly random permutation of the rows
          g = __ldg(shuffle_keys + b * PAL_MAX + rank);
        }
        gray[rank] = g;
      }
      __syncthreads();
      if (tid < n) {
        int rank = tid;
        if ((ordering == PH_ORDER_GRAYNESS || ordering == PH_ORDER_SHUFFLED) && n > 1) {
          const float mine = gray[tid];
          rank = 0;
          for (int e = 0; e < n; ++e) {
            const float other = gray[e];
            rank += (other < mine || (other == mine && e < tid)) ? 1 : 0;  // stable
          }
        }
        out[rank] = unpack_rgba(first_order[tid]);
        final_rank[tid] = rank;
      }
      for (int k = n + tid; k < PAL_MAX; k += PAL_THREADS) out[k] = filler;
      if (tid == 0) ncolors[b] = n;
      if (FUSED_INDEX) {
        // final index of each colour -> its table slot (low word)
        __syncthreads();
        if (tid < n) {
          const unsigned key = first_order[tid];
          unsigned h = hash_slot<PAL_HASH_BITS>(key);
          while ((unsigned)(table[h] >> 32) != key || table[h] == SLOT_EMPTY) h = (h + 1) & (PAL_HASH_SIZE - 1);
          table[h] = ((unsigned long long)key << 32) | (unsigned)final_rank[tid];
        }
      }
    }
    __syncthreads();
  }
  if (!FUSED_INDEX) return;
  if (s_fail) {  // the host raises for this image; keep the outputs defined
    for (int j = 0; row_at(j) < rows; ++j) { const int r = row_at(j); ((r & 1) ? idx1 : idx0)[r >> 1] = 0; }
    return;
  }
  // a pixel equal to the filler colour also matches every padding row: scatter_nd adds them (io_utils.py:84-91)
  const int n = s_n;
  const unsigned filler_key = pack_rgba(make_int4(255, 0, 220, 255));
  const int filler_extra = (PAL_MAX * (PAL_MAX - 1) - n * (n - 1)) / 2;  // sum of n..255
  auto lookup = [&](unsigned key) {
    unsigned h = hash_slot<PAL_HASH_BITS>(key);
    unsigned long long w = table[h];
    while ((unsigned)(w >> 32) != key || w == SLOT_EMPTY) { h = (h + 1) & (PAL_HASH_SIZE - 1); w = table[h]; }
    int idx = (int)(unsigned)w;
    if (key == filler_key) idx += filler_extra;
    return idx;
  };
  if (CACHED) {
#pragma unroll
    for (int j = 0; j < PAL_KEEP; ++j) {
      const int r = row_at(j);
      if (r < rows) ((r & 1) ? idx1 : idx0)[r >> 1] = lookup(kept[j]);
    }
  } else {
    bool ignore = false;
    const int nbatch = (rows + PAL_THREADS * PAL_INFLIGHT - 1) / (PAL_THREADS * PAL_INFLIGHT);
#pragma unroll 1
    for (int bt = 0; bt < nbatch; ++bt) {  // the second read of the pixels (L2), four loads in flight again
      unsigned key[PAL_INFLIGHT];
#pragma unroll
      for (int k = 0; k < PAL_INFLIGHT; ++k) {
        const int r = row_at(bt * PAL_INFLIGHT + k);
        key[k] = r < rows ? load_pixel_key<U8>(src0, src1, r, ignore) : 0u;
      }
#pragma unroll
      for (int k = 0; k < PAL_INFLIGHT; ++k) {
        const int r = row_at(bt * PAL_INFLIGHT + k);
        if (r < rows) ((r & 1) ? idx1 : idx0)[r >> 1] = lookup(key[k]);
      }
    }
  }
}

// =============================================================================================
// rgba_to_indexed (+ optional fused one-hot).
// The palette is hashed once per CTA: slot = (key, sum of the row indices carrying that colour)
// for the reference's scatter-add semantics (duplicates — e.g. a pixel equal to the filler colour —
// add up, io_utils.py:84-91; no match -> 0), or (key, lowest row index) for nearest-colour mode
// where an exact hit short-cuts the arg-min scan.  Values outside [0,255] take an exact
// 4x int32 comparison scan, so any int32 input follows the reference.
// =============================================================================================
constexpr int IDX_THREADS = 256;
constexpr int IDX_PX_PER_THREAD = 4;
constexpr int IDX_HASH_BITS = 10;
constexpr int IDX_HASH_SIZE = 1 << IDX_HASH_BITS;

__device__ __forceinline__ void write_one_hot_rows(float* __restrict__ one_hot, int64_t first_px,
                                                   int64_t npix_total, int idx, int depth, int lane) {
  // the warp writes the `depth` floats of each of its 32 pixels; 16 B per lane per store
  for (int p = 0; p < 32; ++p) {
    const int hot = __shfl_sync(0xffffffffu, idx, p);
    const int64_t px = first_px + p;
    if (px >= npix_total) break;  // warp-uniform
    float* row = one_hot + px * depth;
    if ((depth & 3) == 0) {
      for (int c4 = lane * 4; c4 < depth; c4 += 128) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int d = hot - c4;
        if (d == 0) v.x = 1.f; else if (d == 1) v.y = 1.f; else if (d == 2) v.z = 1.f; else if (d == 3) v.w = 1.f;
        __stcs(reinterpret_cast<float4*>(row + c4), v);
      }
    } else {
      for (int c = lane; c < depth; c += 32) __stcs(row + c, c == hot ? 1.f : 0.f);
    }
  }
}

__global__ void __launch_bounds__(IDX_THREADS) rgba_to_indexed_kernel(
    const int4* __restrict__ image, int64_t npix, const int4* __restrict__ palette,
    int64_t palette_batch, int mode, int* __restrict__ indexed, float* __restrict__ one_hot, int depth) {
  __shared__ unsigned long long table[IDX_HASH_SIZE];
  __shared__ int4 pal[PAL_MAX];
  __shared__ int s_wide;  // some palette row has a channel outside [0,255]

  const int64_t b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31;
  const int4* psrc = palette + (palette_batch == 1 ? 0 : b) * PAL_MAX;
  for (int i = tid; i < IDX_HASH_SIZE; i += IDX_THREADS) table[i] = SLOT_EMPTY;
  if (tid == 0) s_wide = 0;
  __syncthreads();
  if (tid < PAL_MAX) {
    const int4 c = __ldg(psrc + tid);
    pal[tid] = c;
    if (in_byte_range(c)) {
      const unsigned key = pack_rgba(c);
      const unsigned long long word = ((unsigned long long)key << 32) | (unsigned)tid;
      unsigned h = hash_slot<IDX_HASH_BITS>(key);
      for (int probe = 0; probe < IDX_HASH_SIZE; ++probe) {
        unsigned long long cur = atomicCAS(&table[h], SLOT_EMPTY, word);
        if (cur == SLOT_EMPTY) break;
        if ((unsigned)(cur >> 32) == key) {
          if (mode == PH_INDEX_EXACT_SUM) atomicAdd(&table[h], (unsigned long long)tid);
          else atomicMin(&table[h], word);
          break;
        }
        h = (h + 1) & (IDX_HASH_SIZE - 1);
      }
    } else {
      s_wide = 1;
    }
  }
  __syncthreads();
  const bool wide_palette = s_wide != 0;

  const int4* img = image + b * npix;
  int* idx_out = indexed + b * npix;
  float* oh = one_hot ? one_hot + b * npix * (int64_t)depth : nullptr;
  const int64_t cta_first = (int64_t)blockIdx.x * (IDX_THREADS * IDX_PX_PER_THREAD);

#pragma unroll
  for (int it = 0; it < IDX_PX_PER_THREAD; ++it) {
    const int64_t px = cta_first + it * IDX_THREADS + tid;
    int idx = 0;
    if (px < npix) {
      const int4 c = __ldg(img + px);
      bool hit = false;
      if (in_byte_range(c)) {
        const unsigned key = pack_rgba(c);
        unsigned h = hash_slot<IDX_HASH_BITS>(key);
        for (int probe = 0; probe < IDX_HASH_SIZE; ++probe) {
          const unsigned long long w = table[h];
          if (w == SLOT_EMPTY) break;
          if ((unsigned)(w >> 32) == key) { idx = (int)(unsigned)w; hit = true; break; }
          h = (h + 1) & (IDX_HASH_SIZE - 1);
        }
      } else if (wide_palette || mode == PH_INDEX_NEAREST) {
        // exact 128-bit comparison against every row (values that do not pack into a byte key)
        int sum = 0, first = -1;
        for (int k = 0; k < PAL_MAX; ++k) {
          const int4 q = pal[k];
          if (q.x == c.x && q.y == c.y && q.z == c.z && q.w == c.w) { sum += k; if (first < 0) first = k; }
        }
        if (first >= 0) { idx = mode == PH_INDEX_EXACT_SUM ? sum : first; hit = true; }
      }
      if (!hit && mode == PH_INDEX_NEAREST) {
        long long best = 0x7fffffffffffffffll;
        for (int k = 0; k < PAL_MAX; ++k) {
          const int4 q = pal[k];
          const long long dx = (long long)q.x - c.x, dy = (long long)q.y - c.y;
          const long long dz = (long long)q.z - c.z, dw = (long long)q.w - c.w;
          const long long d = dx * dx + dy * dy + dz * dz + dw * dw;
          if (d < best) { best = d; idx = k; }
        }
      }
      idx_out[px] = idx;
    }
    if (oh) {
      const int64_t warp_first = cta_first + it * IDX_THREADS + (tid & ~31);
      if (warp_first < npix) write_one_hot_rows(oh, warp_first, npix, idx, depth, lane);
    }
  }
}

// =============================================================================================
// one-hot (pix2pix_model.py:300-301) and indexed_to_rgba (io_utils.py:96-103)
// =============================================================================================
__global__ void __launch_bounds__(256) one_hot_kernel(const int* __restrict__ indexed, int64_t n,
                                                      int depth, float* __restrict__ one_hot) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t nblk = (n + 31) / 32;
  for (int64_t blk = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); blk < nblk; blk += warps) {
    const int64_t px = blk * 32 + lane;
    const int idx = px < n ? __ldg(indexed + px) : -1;
    write_one_hot_rows(one_hot, blk * 32, n, idx, depth, lane);
  }
}

__global__ void __launch_bounds__(256) indexed_to_rgba_kernel(
    const int* __restrict__ indexed, int64_t npix, const int* __restrict__ palette, int64_t palette_batch,
    int palette_rows, int channels, int* __restrict__ out) {
  const int64_t b = blockIdx.y;
  const int* pal = palette + (palette_batch == 1 ? 0 : b) * (int64_t)palette_rows * channels;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t px = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; px < npix; px += stride) {
    const int idx = __ldg(indexed + b * npix + px);
    const bool ok = idx >= 0 && idx < palette_rows;
    int* dst = out + (b * npix + px) * channels;
    if (channels == 4) {
      const int4 c = ok ? __ldg(reinterpret_cast<const int4*>(pal) + idx) : make_int4(0, 0, 0, 0);
      *reinterpret_cast<int4*>(dst) = c;
    } else {
      for (int ch = 0; ch < channels; ++ch) dst[ch] = ok ? __ldg(pal + (int64_t)idx * channels + ch) : 0;
    }
  }
}

// =============================================================================================
// Indexed-model inference (pix2pix_model.py:283-287 `generate`: argmax over the 256 softmax channels, int32;
// :356 / :446-447 `indexed_to_rgba` of the result) in one pass: 1 KiB read per pixel, 4 + 16 B written — the
// HBM-bound twin of the one-hot writer.  One warp per pixel row (two coalesced 512 B loads when depth is a
// multiple of 128), running first-maximum in index order per lane (strict >: ties keep the earlier index and
// a NaN is never selected, as in the reference's CPU arg-max reducer; a row without any value > -inf gives 0),
// then a warp arg-max that prefers the smaller index on equal values.  Each warp handles 32 consecutive pixels and lane p keeps pixel p's
// result, so indices and colours leave as coalesced 128 B / 512 B stores.
// =============================================================================================
__device__ __forceinline__ void argmax_step(float v, int i, float& best, int& bi) {
  if (v > best) { best = v; bi = i; }
}
__device__ __forceinline__ int warp_argmax(float best, int bi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi >= 0 && (ob > best || (ob == best && (bi < 0 || oi < bi)))) { best = ob; bi = oi; }
  }
  return bi < 0 ? 0 : bi;
}
__global__ void __launch_bounds__(256) argmax_indexed_kernel(
    const float* __restrict__ probs, int64_t npix_total, int64_t npix, int depth, const int* __restrict__ palette,
    int64_t palette_batch, int palette_rows, int* __restrict__ indexed, int* __restrict__ rgba) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = warp * 32; base < npix_total; base += nwarps * 32) {
    int mine = 0;
    const int cnt = (int)min((int64_t)32, npix_total - base);
    if (depth == 256) {
      // the reference's depth (MAX_PALETTE_SIZE): four rows = eight 512 B loads in flight per warp
      for (int k0 = 0; k0 < cnt; k0 += 4) {
        float4 q[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float* row = probs + (base + min(k0 + r, cnt - 1)) * 256;
          q[r][0] = __ldcs(reinterpret_cast<const float4*>(row) + lane);
          q[r][1] = __ldcs(reinterpret_cast<const float4*>(row + 128) + lane);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float best = -INFINITY;
          int bi = -1;
          const int i0 = lane * 4;
          argmax_step(q[r][0].x, i0, best, bi); argmax_step(q[r][0].y, i0 + 1, best, bi);
          argmax_step(q[r][0].z, i0 + 2, best, bi); argmax_step(q[r][0].w, i0 + 3, best, bi);
          argmax_step(q[r][1].x, i0 + 128, best, bi); argmax_step(q[r][1].y, i0 + 129, best, bi);
          argmax_step(q[r][1].z, i0 + 130, best, bi); argmax_step(q[r][1].w, i0 + 131, best, bi);
          const int win = warp_argmax(best, bi);
          if (lane == k0 + r) mine = win;
        }
      }
    } else {
      const bool vec = (depth & 127) == 0;
      for (int k = 0; k < cnt; ++k) {
        const float* row = probs + (base + k) * depth;
        float best = -INFINITY;
        int bi = -1;
        if (vec) {
          for (int c0 = 0; c0 < depth; c0 += 128) {
            const float4 q = __ldcs(reinterpret_cast<const float4*>(row + c0) + lane);
            const int i0 = c0 + lane * 4;
            argmax_step(q.x, i0, best, bi); argmax_step(q.y, i0 + 1, best, bi);
            argmax_step(q.z, i0 + 2, best, bi); argmax_step(q.w, i0 + 3, best, bi);
          }
        } else {
          for (int i = lane; i < depth; i += 32) argmax_step(__ldcs(row + i), i, best, bi);
        }
        const int win = warp_argmax(best, bi);
        if (lane == k) mine = win;
      }
    }
    if (lane < cnt) {
      const int64_t px = base + lane;
      if (indexed) indexed[px] = mine;
      if (rgba) {
        const int64_t b = px / npix;
        const int4* pal = reinterpret_cast<const int4*>(palette) + (palette_batch == 1 ? 0 : b) * (int64_t)palette_rows;
        const int4 c = mine < palette_rows ? __ldg(pal + mine) : make_int4(0, 0, 0, 0);
        reinterpret_cast<int4*>(rgba)[px] = c;
      }
    }
  }
}

// =============================================================================================
// Loader-side pixel prep (dataset_utils.py:66-77 after decode_png): uint8 RGBA -> float32 with
// blacken_transparent_pixels (:11-20, alpha == 0 -> the whole pixel becomes 0) and normalize
// (:39-48, x/127.5 - 1) in one pass.  4 B in, 16 B out per pixel; lets the host ship sprites as the
// uint8 they are on disk (4x fewer PCIe bytes than the float32 tensor the reference uploads).
// =============================================================================================
__global__ void __launch_bounds__(256) u8_to_float_image_kernel(const uchar4* __restrict__ src, int64_t npixels,
                                                                int blacken, int normalize,
                                                                float4* __restrict__ dst) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npixels; i += stride) {
    uchar4 q = __ldg(src + i);
    if (blacken && q.w == 0) q = make_uchar4(0, 0, 0, 0);
    float4 o = make_float4((float)q.x, (float)q.y, (float)q.z, (float)q.w);
    if (normalize) {
      // same two roundings as the reference: (image / 127.5) - 1
      o.x = normalize_px(o.x);
      o.y = normalize_px(o.y);
      o.z = normalize_px(o.z);
      o.w = normalize_px(o.w);
    }
    dst[i] = o;
  }
}

// The loader's pixel helpers as standalone tensor ops (dataset_utils.py:11-20 blacken_transparent_pixels, :39-48
// normalize, :51-60 denormalize): float32 RGBA pixels, 16 B in and 16 B out per thread and iteration.
template <int OP>
__device__ __forceinline__ float4 pixel_map_op(float4 q) {
  if (OP == PH_MAP_BLACKEN) {
    if (q.w == 0.f) q = make_float4(0.f, 0.f, 0.f, 0.f);  // tf.where(alpha == 0, zeros, image); -0.0 == 0 as in TF
  } else if (OP == PH_MAP_NORMALIZE) {
    q.x = normalize_px(q.x); q.y = normalize_px(q.y);
    q.z = normalize_px(q.z); q.w = normalize_px(q.w);
  } else {
    q.x = __fmul_rn(__fadd_rn(q.x, 1.0f), 127.5f); q.y = __fmul_rn(__fadd_rn(q.y, 1.0f), 127.5f);
    q.z = __fmul_rn(__fadd_rn(q.z, 1.0f), 127.5f); q.w = __fmul_rn(__fadd_rn(q.w, 1.0f), 127.5f);
  }
  return q;
}
template <int OP>
__global__ void __launch_bounds__(256) pixel_map_kernel(const float* __restrict__ in, int64_t n, float* __restrict__ out) {
  const int64_t n4 = n >> 2, stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec) {
    for (int64_t i = t; i < n4; i += stride)
      reinterpret_cast<float4*>(out)[i] = pixel_map_op<OP>(__ldg(reinterpret_cast<const float4*>(in) + i));
  } else {
    for (int64_t i = t; i < n4; i += stride) {
      const float4 q = pixel_map_op<OP>(make_float4(in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3]));
      out[4 * i] = q.x; out[4 * i + 1] = q.y; out[4 * i + 2] = q.z; out[4 * i + 3] = q.w;
    }
  }
  if (OP != PH_MAP_BLACKEN && t < (n & 3)) {  // tail of an element count that is not a multiple of four
    const float4 q = pixel_map_op<OP>(make_float4(in[4 * n4 + t], 0.f, 0.f, 1.f));
    out[4 * n4 + t] = q.x;
  }
}

int launch_pixel_map(const float* in, int64_t n, int op, float* out, cudaStream_t st) {
  if (n == 0) return PH_OK;
  int64_t grid = ceil_div(ceil_div(n, 4), 256 * 4);
  const int64_t cap = (int64_t)cached_sm_count() * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  if (op == PH_MAP_BLACKEN) pixel_map_kernel<PH_MAP_BLACKEN><<<(unsigned)grid, 256, 0, st>>>(in, n, out);
  else if (op == PH_MAP_NORMALIZE) pixel_map_kernel<PH_MAP_NORMALIZE><<<(unsigned)grid, 256, 0, st>>>(in, n, out);
  else pixel_map_kernel<PH_MAP_DENORMALIZE><<<(unsigned)grid, 256, 0, st>>>(in, n, out);
  PH_LAUNCH_OK("pixel_map_kernel");
  return PH_OK;
}

int launch_u8_to_float_image(const uint8_t* src, int64_t npixels, int blacken, int normalize, float* dst,
                             cudaStream_t st) {
  if (npixels == 0) return PH_OK;
  int64_t grid = ceil_div(npixels, 256 * 4);
  const int64_t cap = (int64_t)cached_sm_count() * 16;
  if (grid > cap) grid = cap;
  u8_to_float_image_kernel<<<(unsigned)grid, 256, 0, st>>>(reinterpret_cast<const uchar4*>(src), npixels, blacken,
                                                           normalize, reinterpret_cast<float4*>(dst));
  PH_LAUNCH_OK("u8_to_float_image_kernel");
  return PH_OK;
}

// =============================================================================================
// Augmentation of an image pair (dataset_utils.py:80-102 `augment_two`): hue rotation of both
// images by the same delta (tf.image.adjust_hue's (h, v_min, v_max) algorithm, float32 without fused multiply-add
// so that it matches an op-for-op evaluation bit for bit) and one shared nearest-neighbour translation with
// constant fill 0 (keras RandomTranslation -> ImageProjectiveTransformV3), optionally followed by `normalize`
// (:39-48).  One thread per output pixel and image: gather 16 B, write 16 B.  Translation moves pixels, hue maps
// each pixel on its own, so hue(translate(x)) == translate(hue(x)) with the fill applied last.
// =============================================================================================
__device__ __forceinline__ void adjust_hue_pixel(float& r, float& g, float& b, float shift6) {
  float vmax, vmid, vmin;
  int cat;
  // rgb_to_hv_range: the comparison tree decides ties
  if (r < g) {
    if (b < r) { vmax = g; vmid = r; vmin = b; cat = 1; }
    else if (b > g) { vmax = b; vmid = g; vmin = r; cat = 3; }
    else { vmax = g; vmid = b; vmin = r; cat = 2; }
  } else {
    if (b < g) { vmax = r; vmid = g; vmin = b; cat = 0; }
    else if (b > r) { vmax = b; vmid = r; vmin = g; cat = 4; }
    else { vmax = r; vmid = b; vmin = g; cat = 5; }
  }
  const float span = __fsub_rn(vmax, vmin);
  float h = 0.f;
  if (vmax != vmin) {
    const float ratio = __fdiv_rn(__fsub_rn(vmid, vmin), span);
    h = __fadd_rn((float)cat, (cat & 1) ? __fsub_rn(1.0f, ratio) : ratio);
  }
  h = __fadd_rn(h, shift6);
  for (int k = 0; k < 4 && h < 0.f; ++k) h = __fadd_rn(h, 6.0f);   // |delta| <= 1 is enforced by the caller
  for (int k = 0; k < 4 && h >= 6.f; ++k) h = __fsub_rn(h, 6.0f);
  int c2 = (int)h;
  float ratio2 = __fsub_rn(h, (float)c2);
  if (c2 & 1) ratio2 = __fsub_rn(1.0f, ratio2);
  const float mid = __fadd_rn(vmin, __fmul_rn(ratio2, span));
  switch (c2) {
    case 0: r = vmax; g = mid; b = vmin; break;
    case 1: r = mid; g = vmax; b = vmin; break;
    case 2: r = vmin; g = vmax; b = mid; break;
    case 3: r = vmin; g = mid; b = vmax; break;
    case 4: r = mid; g = vmin; b = vmax; break;
    default: r = vmax; g = vmin; b = mid; break;
  }
}

constexpr int AUG_PX_PER_THREAD = 4;

// grid.x = image (first images, then second images), grid.y = chunk of 1024 pixels; 32-bit index arithmetic, four
// independent 16-byte gathers in flight per thread
__global__ void __launch_bounds__(256) augment_pair_kernel(const float4* __restrict__ first,
                                                           const float4* __restrict__ second, int batch,
                                                           int height, int width,
                                                           const float* __restrict__ hue_delta,
                                                           const float* __restrict__ translation,
                                                           const uint8_t* __restrict__ apply, int normalize,
                                                           float4* __restrict__ out_first,
                                                           float4* __restrict__ out_second) {
  const int npix = height * width;
  const int which = (int)blockIdx.x >= batch;
  const int b = (int)blockIdx.x - which * batch;
  const float4* src = (which ? second : first) + (int64_t)b * npix;
  float4* dst = (which ? out_second : out_first) + (int64_t)b * npix;
  const bool on = apply == nullptr || apply[b] != 0;
  const bool move = on && translation != nullptr, rotate = on && hue_delta != nullptr;
  const float ndx = move ? -__ldg(translation + 2 * b) : 0.f, ndy = move ? -__ldg(translation + 2 * b + 1) : 0.f;
  const float shift6 = rotate ? __fmul_rn(__ldg(hue_delta + b), 6.0f) : 0.f;
  const int px0 = blockIdx.y * (256 * AUG_PX_PER_THREAD) + threadIdx.x;
  float4 q[AUG_PX_PER_THREAD];
  bool inside[AUG_PX_PER_THREAD];
#pragma unroll
  for (int k = 0; k < AUG_PX_PER_THREAD; ++k) {
    const int px = px0 + k * 256;
    const int y = px / width, x = px - y * width;
    int s = px;
    inside[k] = px < npix;
    if (move) {
      // input = 1*x + 0*y + (-dx), rounded half away from zero (std::round); outside -> fill value 0
      const float fx = roundf(__fadd_rn((float)x, ndx)), fy = roundf(__fadd_rn((float)y, ndy));
      inside[k] = inside[k] && fx >= 0.f && fx < (float)width && fy >= 0.f && fy < (float)height;
      s = inside[k] ? (int)fy * width + (int)fx : 0;
    }
    q[k] = inside[k] ? __ldg(src + s) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < AUG_PX_PER_THREAD; ++k) {
    const int px = px0 + k * 256;
    if (px >= npix) continue;
    if (rotate && inside[k]) adjust_hue_pixel(q[k].x, q[k].y, q[k].z, shift6);
    if (normalize) {
      q[k].x = normalize_px(q[k].x);
      q[k].y = normalize_px(q[k].y);
      q[k].z = normalize_px(q[k].z);
      q[k].w = normalize_px(q[k].w);
    }
    __stcs(dst + px, q[k]);
  }
}

int launch_augment_pair(const float* first, const float* second, int64_t batch, int height, int width,
                        const float* hue_delta, const float* translation, const uint8_t* apply, int normalize,
                        float* out_first, float* out_second, cudaStream_t st) {
  const int64_t images = batch * (second ? 2 : 1);
  PH_CHECK_ARG(images < (1ll << 31), "batch too large");
  if (images == 0) return PH_OK;
  const int64_t chunks = ceil_div((int64_t)height * width, 256 * AUG_PX_PER_THREAD);
  PH_CHECK_ARG(chunks <= 65535, "image too large");
  augment_pair_kernel<<<dim3((unsigned)images, (unsigned)chunks), 256, 0, st>>>(
      reinterpret_cast<const float4*>(first), reinterpret_cast<const float4*>(second), (int)batch, height, width,
      hue_delta, translation, apply, normalize, reinterpret_cast<float4*>(out_first),
      reinterpret_cast<float4*>(out_second));
  PH_LAUNCH_OK("augment_pair_kernel");
  return PH_OK;
}

// =============================================================================================
// launchers
// =============================================================================================
template <bool FUSED, bool U8>
static int launch_extract_any(const void* image, const void* image2, int64_t batch, int64_t rows, int ordering,
                              const float* shuffle_keys, int32_t* palette, int32_t* ncolors, int32_t* idx1,
                              int32_t* idx2, cudaStream_t st) {
  PH_CHECK_ARG(batch < (1ll << 31), "batch too large");
  PH_CHECK_ARG(rows < (1ll << 31) - 2 * PAL_THREADS * PAL_INFLIGHT, "too many rows per image (%lld)", (long long)rows);
  PH_CHECK_ARG(ordering != PH_ORDER_SHUFFLED || shuffle_keys != nullptr, "'shuffled' ordering needs shuffle keys");
  if (batch == 0) return PH_OK;
  if (rows <= (int64_t)PAL_THREADS * PAL_KEEP)
    extract_palette_kernel<FUSED, U8, true><<<(unsigned)batch, PAL_THREADS, 0, st>>>(
        image, image2, rows, ordering, shuffle_keys, reinterpret_cast<int4*>(palette), ncolors, idx1, idx2);
  else
    extract_palette_kernel<FUSED, U8, false><<<(unsigned)batch, PAL_THREADS, 0, st>>>(
        image, image2, rows, ordering, shuffle_keys, reinterpret_cast<int4*>(palette), ncolors, idx1, idx2);
  PH_LAUNCH_OK(FUSED ? "extract_palette_kernel<fused index>" : "extract_palette_kernel");
  return PH_OK;
}

int launch_extract_palette(const int32_t* image, const int32_t* image2, int64_t batch, int64_t rows, int ordering,
                           const float* shuffle_keys, int32_t* palette, int32_t* ncolors, cudaStream_t st) {
  return launch_extract_any<false, false>(image, image2, batch, rows, ordering, shuffle_keys, palette, ncolors, nullptr,
                                          nullptr, st);
}

// dataset_utils.py:138-151 in ONE launch: shared palette of source||target and both index images.  elem_bytes 4:
// int32 RGBA pixels; 1: the decoded PNG's uint8 RGBA (a quarter of the bytes, same results)
int launch_load_indexed_fused(const void* source, const void* target, int elem_bytes, int64_t batch, int64_t npix,
                              int ordering, const float* shuffle_keys, int32_t* source_indexed,
                              int32_t* target_indexed, int32_t* palette, int32_t* ncolors, cudaStream_t st) {
  PH_CHECK_ARG(2 * npix < (1ll << 31), "too many pixels per image (%lld)", (long long)npix);
  if (elem_bytes == 1)
    return launch_extract_any<true, true>(source, target, batch, 2 * npix, ordering, shuffle_keys, palette, ncolors,
                                          source_indexed, target_indexed, st);
  return launch_extract_any<true, false>(source, target, batch, 2 * npix, ordering, shuffle_keys, palette, ncolors,
                                         source_indexed, target_indexed, st);
}

int launch_rgba_to_indexed(const int32_t* image, int64_t batch, int64_t npix, const int32_t* palette,
                           int64_t palette_batch, int mode, int32_t* indexed, float* one_hot, int depth,
                           cudaStream_t st) {
  PH_CHECK_ARG(batch <= 65535, "rgba_to_indexed: batch %lld > 65535 per call", (long long)batch);
  if (batch == 0 || npix == 0) return PH_OK;
  const dim3 grid((unsigned)ceil_div(npix, IDX_THREADS * IDX_PX_PER_THREAD), (unsigned)batch);
  rgba_to_indexed_kernel<<<grid, IDX_THREADS, 0, st>>>(
      reinterpret_cast<const int4*>(image), npix, reinterpret_cast<const int4*>(palette), palette_batch,
      mode, indexed, one_hot, depth);
  PH_LAUNCH_OK("rgba_to_indexed_kernel");
  return PH_OK;
}

int launch_one_hot(const int32_t* indexed, int64_t n, int depth, float* one_hot, cudaStream_t st) {
  if (n == 0) return PH_OK;
  int64_t grid = ceil_div(n, 256);
  const int64_t cap = (int64_t)cached_sm_count() * 16;
  if (grid > cap) grid = cap;
  one_hot_kernel<<<(unsigned)grid, 256, 0, st>>>(indexed, n, depth, one_hot);
  PH_LAUNCH_OK("one_hot_kernel");
  return PH_OK;
}

int launch_argmax_indexed(const float* probs, int64_t batch, int64_t npix, int depth, const int32_t* palette,
                          int64_t palette_batch, int palette_rows, int32_t* indexed, int32_t* rgba, cudaStream_t st) {
  const int64_t total = batch * npix;
  if (total == 0) return PH_OK;
  int64_t grid = ceil_div(ceil_div(total, 32), 8);  // 8 warps per CTA, 32 pixels per warp and round
  const int64_t cap = (int64_t)cached_sm_count() * 8;
  if (grid > cap) grid = cap;
  argmax_indexed_kernel<<<(unsigned)grid, 256, 0, st>>>(probs, total, npix, depth, palette, palette_batch, palette_rows,
                                                        indexed, rgba);
  PH_LAUNCH_OK("argmax_indexed_kernel");
  return PH_OK;
}

int launch_indexed_to_rgba(const int32_t* indexed, int64_t batch, int64_t npix, const int32_t* palette,
                           int64_t palette_batch, int palette_rows, int channels, int32_t* out,
                           cudaStream_t st) {
  PH_CHECK_ARG(batch <= 65535, "indexed_to_rgba: batch %lld > 65535 per call", (long long)batch);
  if (batch == 0 || npix == 0) return PH_OK;
  int64_t gx = ceil_div(npix, 256);
  if (gx > 1024) gx = 1024;
  indexed_to_rgba_kernel<<<dim3((unsigned)gx, (unsigned)batch), 256, 0, st>>>(
      indexed, npix, palette, palette_batch, palette_rows, channels, out);
  PH_LAUNCH_OK("indexed_to_rgba_kernel");
  return PH_OK;
}

}  // namespace ph
