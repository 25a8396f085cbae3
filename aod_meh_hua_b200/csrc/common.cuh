// Shared device/host helpers for the mehhua kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include "../../include/mehhua.h"

namespace mehhua {

constexpr int kMaxLevels = MEHHUA_MAX_LEVELS;
// Capture mode of K1 (sparse top-k levels: n >= kCapMinRatio * k, or 16 n <= N; build_plan decides): a pre-pass over 1/kCapStride of the level's
// warps estimates the key of rank kCapTargetNum / kCapTargetDen * k; the streaming pass parks the score row of every prior at or above
// it (at most kCapRows per (image, level)); the select then runs over the parked rows only.  A level
// whose estimate misses (fewer than k or more than kCapRows parked rows) falls back to the select
// over all keys and the strided gather - results are identical either way.
constexpr int kCapRows = 4096;
// floats of a parked row: the C exponentials, 1/sum, the score normaliser, padded to a multiple of 4 (16-byte rows:
// vector stores in K1a, one bulk-async copy per row in K1c)
__host__ __device__ constexpr int cap_row_floats(int C) { return (C + 2 + 3) & ~3; }
#ifndef MEHHUA_CAP_STRIDE
#define MEHHUA_CAP_STRIDE 32
#endif
#ifndef MEHHUA_CAP_TARGET_NUM
#define MEHHUA_CAP_TARGET_NUM 7
#endif
#ifndef MEHHUA_CAP_MIN_RATIO
#define MEHHUA_CAP_MIN_RATIO 8
#endif
constexpr int kCapStride = MEHHUA_CAP_STRIDE;
constexpr int kCapTargetNum = MEHHUA_CAP_TARGET_NUM, kCapTargetDen = 4;
constexpr int kCapMinRatio = MEHHUA_CAP_MIN_RATIO;
constexpr int kCapMinWarps = 16;      // sampled warps per (image, level) the stride is reduced to reach on small levels
constexpr int kCapSampleMax = 8192;   // sampled priors per (image, level) the threshold kernel can hold
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// One FPN / SSD level of the current batch, as the kernels see it.
struct LevelDev {
  const float* logits;   // [B, A*C, H, W]
  const float* deltas;   // [B, A*4, H, W]
  const float* lam;      // [B, A, H, W]
  const float* anchors;  // [N, 4]
  int H, W, A, HW;
  int n;      // priors in the level (H*W*A)
  int k;      // rows kept after the per-level top-k
  int n_off;  // first prior of the level inside one image's key array
  int k_off;  // first row of the level inside one image's K_tot rows
  int tpp;    // K1a tiles per (image, anchor) plane
  int tile0;  // first K1a tile of the level inside one image's tile list
  int topk;   // 1 when n > k (the per-level top-k is active)
  int rescan; // 1 when the level's rows are produced by the coalesced rescan kernel (no top-k, or 2 k >= n)
  int rtile0; // first rescan tile of the level inside one image's rescan tile list
  int cap;    // >= 0: the level's rows are CAPTURED while K1a streams it (index of the level among the capture levels); -1: not
};

// Everything a kernel needs about the batch, passed by value (__grid_constant__).
struct Plan {
  LevelDev lv[kMaxLevels];
  int S, B, C, head, num_fg;
  int N;                // priors per image
  int K;                // rows per image (K_tot)
  int row_stride;       // rows per image in score_rows / lam_rows (K; pair_cap in Entropy_ALL mode)
  int tiles_per_image;  // K1a tiles per image
  int rtiles_per_image; // K1c rescan tiles per image
  int n_cap_levels;     // levels in capture mode
  int nms_pre, max_per_img, pair_cap, n_samples;
  int use_lambda, agg_object, agg_scale, agg_class, cls_w, rescale;
  int act;              // MEHHUA_ACT_*
  int mode;             // MEHHUA_MODE_*
  float score_thr, nms_iou, fg_thr, obj_thr, cluster_iou, lambda_scale, lambda_eps;
  float means[4], stds[4];
  float max_ratio;      // |ln(wh_ratio_clip)|
  unsigned long long seed;
};

// Internal scratch carved out of the caller's workspace.
struct Workspace {
  float* keys;                 // [B, N]   ranking key of every prior, anchor-major inside a level
  unsigned long long* cand;    // [B, K*num_fg] NMS candidates (score bits << 32 | ~flat)
  int* cand_cnt;               // [B]
  unsigned* cand_maxc;         // [B]      max candidate box coordinate, ordered-uint encoded
  unsigned* status;            // [1]
  int* work_counter;           // [4]      dynamic work queues
  int* k2_done;                // [kK2SplitPairs] arrival counters of split K2 pairs
  float* k2_part;              // [kK2SplitPairs, kK2Sub, C+1] partial class / entropy sums of split K2 pairs
  int* inv_map;                // [B, N]   position -> row of the dense top-k levels (-1 = not kept)
  unsigned* fg_mask;           // [B, tiles_per_image, 4] Entropy_ALL: one foreground bit per prior (a ballot word per warp of a tile)
  int* tile_cnt;               // [B, tiles_per_image] foreground priors of each tile
  int* tile_pref;              // [B, tiles_per_image] exclusive prefix of tile_cnt inside the tile's (level, anchor) plane
  float* lam_part;             // [B, tiles_per_image] per-tile lambda sums (Entropy_ALL)
  float* tau;                  // [B, S]   capture threshold of a capture level (K1t)
  int* cap_cnt;                // [B, S]   rows captured so far / in total
  unsigned long long* cap_comp;// [B, n_cap_levels, kCapRows] composite (key bits << 32 | ~position) of each captured row
  float* cap_scores;           // [B, n_cap_levels, kCapRows, cap_row_floats(C)] exponentials + normalisers of each captured row
  int* row_slot;               // [B, K]   capture slot of a kept row, -1 = take it from the logits (gather)
  size_t bytes;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// order-preserving float <-> uint (for atomicMax on floats of either sign)
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// level that owns row r of an image (rows are level-major)
__device__ __forceinline__ int level_of_row(const Plan& p, int r) {
  int s = 0;
#pragma unroll
  for (int i = 1; i < kMaxLevels; ++i)
    if (i < p.S && r >= p.lv[i].k_off) s = i;
  return s;
}

// level that owns pair q of an image: pair_off[s] = first pair of level s (levels are contiguous)
__device__ __forceinline__ int level_of_pair(const int* __restrict__ poff, int S, int q) {
  int s = 0;
  for (int i = 1; i < S; ++i)
    if (q >= poff[i]) s = i;
  return s;
}

// Inclusive prefix sum over the threads of a block (thread order).  `warp_sums` holds >= 32 ints.
template <int THREADS>
__device__ __forceinline__ int block_incl_scan(int v, int* warp_sums) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  __syncthreads();  // protect warp_sums reuse
  if (lane == 31) warp_sums[w] = v;
  __syncthreads();
  if (w == 0) {
    int s = (lane < THREADS / 32) ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    warp_sums[lane] = s;
  }
  __syncthreads();
  if (w > 0) v += warp_sums[w - 1];
  return v;
}

// Bitonic sort, descending, of buf[0..n2) (n2 a power of two) by the whole block.
// Compare-exchanges at distance < 32 never leave a warp: element e is held by thread e mod THREADS, so its
// partner e ^ stride sits in the same warp and the exchange is a shuffle - all stages of the sizes 2..32 and the
// last five stages of every larger size run in registers without a barrier (n2 = 2048: 28 barriers instead of 66).
__device__ __forceinline__ unsigned long long bitonic_xchg(unsigned long long v, const int e, const int size, const int stride) {
  const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, stride);
  const bool keep_max = ((e & stride) == 0) == ((e & size) == 0);
  return (keep_max == (v > o)) ? v : o;
}
template <int THREADS>
__device__ __forceinline__ void block_bitonic_desc(unsigned long long* buf, int n2) {
  if (n2 < 64) {      // tiny inputs: every stage through shared memory
    for (int size = 2; size <= n2; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = threadIdx.x; i < (n2 >> 1); i += THREADS) {
          const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));   // stride is a power of two
          const int hi = lo + stride;
          const bool desc = (lo & size) == 0;
          const unsigned long long a = buf[lo], b = buf[hi];
          if ((a < b) == desc) { buf[lo] = b; buf[hi] = a; }
        }
        __syncthreads();
      }
    }
    return;
  }
  // sizes 2..32 entirely in registers (n2 and THREADS are multiples of 32: whole warps take every trip together)
  for (int e = threadIdx.x; e < n2; e += THREADS) {
    unsigned long long v = buf[e];
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) v = bitonic_xchg(v, e, size, stride);
    buf[e] = v;
  }
  __syncthreads();
  for (int size = 64; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride >= 32; stride >>= 1) {
      for (int i = threadIdx.x; i < (n2 >> 1); i += THREADS) {
        const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = buf[lo], b = buf[hi];
        if ((a < b) == desc) { buf[lo] = b; buf[hi] = a; }
      }
      __syncthreads();
    }
    for (int e = threadIdx.x; e < n2; e += THREADS) {
      unsigned long long v = buf[e];
#pragma unroll
      for (int stride = 16; stride > 0; stride >>= 1) v = bitonic_xchg(v, e, size, stride);
      buf[e] = v;
    }
    __syncthreads();
  }
}

constexpr int kSelUnroll = 4;    // loads in flight per thread in the select passes (they are latency-bound)

// Radix-select digit schedules (MSB first) over a 64-bit composite.
// schedule 0: composites whose two top bits are zero (positive floats < 2.0 in the high word):
//             the first 12-bit digit is float bits 29..18 = exponent + 5 mantissa bits.
// schedule 1: arbitrary 64-bit composites.
__device__ __constant__ const int kDigitShift[2][7] = {{50, 41, 32, 24, 16, 8, 0}, {52, 43, 34, 26, 18, 10, 0}};
__device__ __constant__ const int kDigitBits[2][7]  = {{12,  9,  9,  8,  8, 8, 8}, {12,  9,  9,  8,  8,  8, 10}};

// Block-cooperative top-k collection.  `get(i)` returns element i (i < n) as a 64-bit composite,
// all composites distinct (and < 2^62 for SCHED 0); a composite of 0 marks "not an element".  Collects into buf (capacity CAP, sorted descending) every
// element e < hi that is >= a threshold chosen so that the set holds the k largest such elements
// (all of them when fewer than k exist) and at most CAP elements.  Returns the count.
// hist: 4096 ints of shared memory; sh: 40 ints of shared memory.
template <int THREADS, int CAP, int SCHED, class Get>
__device__ int block_collect_topk(Get get, int n, int k, unsigned long long hi,
                                  unsigned long long* buf, int* hist, int* sh, unsigned* status) {
  unsigned long long prefix = 0, mask = 0;
  int need = k, above = 0;
  for (int pass = 0; pass < 7; ++pass) {
    const int shift = kDigitShift[SCHED][pass], bits = kDigitBits[SCHED][pass], bins = 1 << bits;
    for (int i = threadIdx.x; i < bins; i += THREADS) hist[i] = 0;
    __syncthreads();
    for (int i0 = threadIdx.x; i0 < n; i0 += THREADS * kSelUnroll) {   // kSelUnroll independent loads in flight
      unsigned long long e[kSelUnroll];
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u) { const int i = i0 + u * THREADS; e[u] = (i < n) ? get(i) : 0ull; }
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u)
        if (e[u] != 0ull && e[u] < hi && (e[u] & mask) == prefix)
          atomicAdd(&hist[(int)((e[u] >> shift) & (bins - 1))], 1);
    }
    __syncthreads();
    // thread t owns the t-th highest chunk of bins; scan chunk sums from the top
    const int per = (bins + THREADS - 1) / THREADS;
    const int top = bins - 1 - (int)threadIdx.x * per;  // highest bin of my chunk (may be < 0)
    int csum = 0;
    for (int j = 0; j < per; ++j) { const int b = top - j; if (b >= 0) csum += hist[b]; }
    const int incl = block_incl_scan<THREADS>(csum, sh);
    if (threadIdx.x == THREADS - 1) sh[32] = incl;  // elements matching the prefix
    if (incl >= need && incl - csum < need) {        // the crossing chunk (unique)
      int acc = incl - csum, b = top;
      for (;; --b) { if (acc + hist[b] >= need) break; acc += hist[b]; }
      sh[33] = b; sh[34] = acc; sh[35] = hist[b];
    }
    __syncthreads();
    const int matching = sh[32];
    if (matching < need) {   // fewer than k elements exist: take every match
      __syncthreads();
      break;
    }
    const int d = sh[33], cnt_above = sh[34], within = sh[35];
    __syncthreads();
    above += cnt_above;
    need -= cnt_above;
    prefix |= (unsigned long long)d << shift;
    mask |= (unsigned long long)(bins - 1) << shift;
    if (above + within <= CAP) break;
    if (threadIdx.x == 0 && pass == 0) atomicOr(status, MEHHUA_ST_SELECT_SLOWPATH);
  }
  // compaction of every e in [prefix, hi)
  if (threadIdx.x == 0) sh[36] = 0;
  __syncthreads();
  for (int i0 = threadIdx.x; i0 < n; i0 += THREADS * kSelUnroll) {
    unsigned long long e[kSelUnroll];
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) { const int i = i0 + u * THREADS; e[u] = (i < n) ? get(i) : 0ull; }
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u)
      if (e[u] != 0ull && e[u] < hi && e[u] >= prefix) {
        const int pos = atomicAdd(&sh[36], 1);
        if (pos < CAP) buf[pos] = e[u];
      }
  }
  __syncthreads();
  int cnt = sh[36];
  if (cnt > CAP) cnt = CAP;
  int n2 = 1;
  while (n2 < cnt) n2 <<= 1;
  for (int i = cnt + threadIdx.x; i < n2; i += THREADS) buf[i] = 0ull;
  __syncthreads();
  block_bitonic_desc<THREADS>(buf, n2);
  return cnt;
}

// IoU of mmdet bbox_overlaps (iou2d_calculator.py:206-252): no +1, union floored at 1e-6.
// Every operation is an explicitly rounded fp32 op so the result is bit-identical to the
// reference's ATen sequence on the same inputs.
__device__ __forceinline__ float iou_overlaps(const float4 a, const float area_a, const float4 b) {
  const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
  const float ov = __fmul_rn(w, h);
  const float un = fmaxf(__fsub_rn(__fadd_rn(area_a, area_b), ov), 1e-6f);
  return __fdiv_rn(ov, un);
}

}  // namespace mehhua
