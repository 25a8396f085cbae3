// Entropy_ALL scoring mode (uncertainty_pool = 'Entropy_ALL'): no NMS / objects - every prior whose
// max foreground softmax exceeds fg_thr is sampled, grouped by (level, class) and aggregated with
// one of the four scale/class aggregation types.
//
// Reference semantics: ComputeScaleUnc (mmdet/models/dense_heads/Lambda_L2.py:539-569,
// My_L_ssd_head.py:484-515): p = softmax(logits); FG = max_c p > 0.3 (SSD: foreground classes);
// lambda' = mean(lambda over ALL priors of the level) / (lambda + 1e-7) * 25; alpha = p * lambda';
// class key = argmax_c alpha = argmax_c p.  AggregateScaleUnc: Lambda_L2.py:636-691.
//
// KA1 streams the logits once (same tiling as K1a), appends the foreground priors of each image to
// an unordered list and writes one lambda partial sum per tile; KA2 (one block per image) sorts the
// list by (level, prior) so that the result does not depend on scheduling, reduces the lambda
// partials in tile order, and materialises the foreground rows (softmax p, lambda, class) in the
// same buffers the Entropy_NMS route uses, with rows == pairs and a single pseudo-object - K2 and
// K3c then run unchanged.
#pragma once
#include "common.cuh"
#include "k1_alpha_topk.cuh"

namespace mehhua {

constexpr int kAllSortCap = 16384;    // foreground priors per image that KA2 can order in shared memory
constexpr int kAllThreads = 512;
constexpr size_t kAllSmem = (size_t)kAllSortCap * 8 + 64 * 4;

template <int C, int HEAD>
__global__ void __launch_bounds__(kK1aThreads)
ka_fg_kernel(const __grid_constant__ Plan p, unsigned* __restrict__ fg_list, int* __restrict__ fg_cnt,
             float* __restrict__ lam_part, unsigned* __restrict__ status, unsigned* __restrict__ level_maxconf) {
  __shared__ float wsum[kK1aThreads / 32];
  const int t = blockIdx.x;
  const int b = t / p.tiles_per_image;
  const int ti = t - b * p.tiles_per_image;
  int s = 0;
#pragma unroll
  for (int i = 1; i < kMaxLevels; ++i)
    if (i < p.S && ti >= p.lv[i].tile0) s = i;
  const LevelDev& L = p.lv[s];
  const int lt = ti - L.tile0;
  const int a = lt / L.tpp;
  const int hw = (lt - a * L.tpp) * kK1aThreads + threadIdx.x;
  const bool live = hw < L.HW;
  const int CC = (C > 0) ? C : p.C;
  float lam = 0.f;
  if (live) {
    const float* __restrict__ src = L.logits + ((size_t)(b * L.A + a) * CC) * L.HW + hw;
    float inv, den, pfg;
    if constexpr (C > 0) {
      float x[C];
#pragma unroll
      for (int c = 0; c < C; ++c) x[c] = __ldg(src + (size_t)c * L.HW);
      softmax_regs<C, HEAD>(x, inv, den, pfg);
    } else {
      float m;
      softmax_stream<HEAD>(src, (size_t)L.HW, CC, m, inv, den, pfg);
    }
    lam = __ldg(L.lam + (size_t)(b * L.A + a) * L.HW + hw);
    if (level_maxconf) level_maxconf_update(level_maxconf + b * p.S + s, inv);
    if (pfg > p.fg_thr) {
      const int pos = atomicAdd(fg_cnt + b, 1);
      if (pos < p.pair_cap) fg_list[(size_t)b * p.pair_cap + pos] = ((unsigned)s << 28) | (unsigned)(hw * L.A + a);
      else atomicOr(status, MEHHUA_ST_PAIR_OVERFLOW);
    }
  }
  // lambda partial sum of the tile, fixed reduction tree
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lam += __shfl_xor_sync(0xffffffffu, lam, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = lam;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kK1aThreads / 32; ++w) v += wsum[w];
    lam_part[(size_t)b * p.tiles_per_image + ti] = v;
  }
}

template <int C, int HEAD>
__global__ void __launch_bounds__(kAllThreads)
ka_finalize_kernel(const __grid_constant__ Plan p, const unsigned* __restrict__ fg_list,
                   const int* __restrict__ fg_cnt, const float* __restrict__ lam_part,
                   float* __restrict__ score_rows, float* __restrict__ lam_rows, int* __restrict__ topk_idx,
                   float* __restrict__ row_max, int* __restrict__ row_argmax, int* __restrict__ level_fg,
                   int* __restrict__ pair_row, int* __restrict__ pair_obj, int* __restrict__ pair_cls,
                   int* __restrict__ pair_off, float* __restrict__ lam_mean, int* __restrict__ n_obj,
                   int* __restrict__ n_det, unsigned* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char ka_smem[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ka_smem);   // kAllSortCap
  int* sh = reinterpret_cast<int*>(keys + kAllSortCap);                         // 64
  const int b = blockIdx.x;
  int n = min(fg_cnt[b], p.pair_cap);
  if (n > kAllSortCap) {
    if (threadIdx.x == 0) atomicOr(status, MEHHUA_ST_PAIR_OVERFLOW);
    n = kAllSortCap;
  }
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  // ascending (level, prior) order via a descending sort of the inverted key
  for (int i = threadIdx.x; i < n2; i += kAllThreads)
    keys[i] = (i < n) ? (unsigned long long)(~fg_list[(size_t)b * p.pair_cap + i]) : 0ull;
  if (threadIdx.x < kMaxLevels + 1) sh[threadIdx.x] = 0;
  __syncthreads();
  block_bitonic_desc<kAllThreads>(keys, n2);
  // level offsets: count the entries of each level
  for (int i = threadIdx.x; i < n; i += kAllThreads) atomicAdd(&sh[(~(unsigned)keys[i]) >> 28], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int s = 0; s < p.S; ++s) {
      pair_off[b * (p.S + 1) + s] = acc;
      level_fg[b * p.S + s] = sh[s] > 0;
      acc += sh[s];
    }
    pair_off[b * (p.S + 1) + p.S] = acc;
    n_obj[b] = acc > 0 ? 1 : 0;
    n_det[b] = 0;
  }
  // mean lambda over ALL priors of each level: tile partials summed in tile order by one thread
  if (threadIdx.x >= 32 && threadIdx.x < 32 + p.S) {
    const int s = threadIdx.x - 32;
    const LevelDev& L = p.lv[s];
    const float* lp = lam_part + (size_t)b * p.tiles_per_image + L.tile0;
    float v = 0.f;
    for (int i = 0; i < L.tpp * L.A; ++i) v += lp[i];
    lam_mean[b * p.S + s] = __fdiv_rn(v, (float)L.n);
  }
  // foreground rows: softmax p, lambda, class key; rows == pairs, single pseudo-object 0
  const int CC = (C > 0) ? C : p.C;
  for (int q = threadIdx.x; q < n; q += kAllThreads) {
    const unsigned e = ~(unsigned)keys[q];
    const int s = (int)(e >> 28), prior = (int)(e & 0x0fffffffu);
    const LevelDev& L = p.lv[s];
    const int hw = prior / L.A, a = prior - hw * L.A;
    const float* __restrict__ src = L.logits + ((size_t)(b * L.A + a) * CC) * L.HW + hw;
    float* srow = score_rows + ((size_t)b * p.row_stride + q) * CC;
    float best = -1.f;
    int arg = 0;
    if constexpr (C > 0) {
      float x[C];
#pragma unroll
      for (int c = 0; c < C; ++c) x[c] = __ldg(src + (size_t)c * L.HW);
      float inv, den, pfg;
      softmax_regs<C, HEAD>(x, inv, den, pfg);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float pc = __fmul_rn(x[c], inv);
        srow[c] = pc;
        if (pc > best) { best = pc; arg = c; }
      }
    } else {
      float m, inv, den, pfg;
      softmax_stream<HEAD>(src, (size_t)L.HW, CC, m, inv, den, pfg);
      const float nml2 = -__fmul_rn(m, kLog2e);
      for (int c = 0; c < CC; ++c) {
        const float pc = __fmul_rn(ex2_approx(fmaf(__ldg(src + (size_t)c * L.HW), kLog2e, nml2)), inv);
        srow[c] = pc;
        if (pc > best) { best = pc; arg = c; }
      }
    }
    const size_t rq = (size_t)b * p.row_stride + q;
    lam_rows[rq] = __ldg(L.lam + (size_t)(b * L.A + a) * L.HW + hw);
    topk_idx[rq] = prior;
    row_max[rq] = best;
    row_argmax[rq] = arg;
    const size_t pq = (size_t)b * p.pair_cap + q;
    pair_row[pq] = q;
    pair_obj[pq] = 0;
    pair_cls[pq] = arg;
  }
}

}  // namespace mehhua
