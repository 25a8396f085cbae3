// Entropy_ALL family of scoring modes: no NMS / objects - every prior that passes a foreground test is
// sampled and the per-prior epistemic uncertainties are averaged per (level, class) and aggregated.
//
// Reference semantics:
//   uncertainty_pool = 'Entropy_ALL': ComputeScaleUnc (mmdet/models/dense_heads/Lambda_L2.py:539-569,
//     My_L_ssd_head.py:484-515): p = softmax(logits); FG = max_c p > 0.3 (SSD: foreground classes);
//     lambda' = mean(lambda over ALL priors of the level) / (lambda + 1e-7) * 25; alpha = p * lambda';
//     class key = argmax_c alpha = argmax_c p; T = 500.  AggregateScaleUnc: Lambda_L2.py:636-691 (4 types).
//   uncertainty_pool = 'Entropy_Avg' (ablation heads): ComputeAvgUnc (Lambda_L2_ReLU.py:446-474):
//     r = relu(logits); prob = r / (sum r + 1e-9); FG = max_c prob > 0.3; alpha = r * lambda' (zeros allowed);
//     T = 50; per level the mean epistemic uncertainty over ALL FG priors (no class split);
//     AggregateAvgUnc (:532-541): mean over the levels that have one.
//
// Any number of foreground priors per image is handled (the row buffers hold pair_cap of them; more set
// MEHHUA_ST_PAIR_OVERFLOW).  Four kernels, none of which sorts or takes an order-dependent atomic:
//   KA1 ka_fg_kernel     streams the logits once (K1a's tiling): one foreground bit per prior (a ballot
//                        word per warp), a count per tile, a lambda partial sum per tile
//   KA2 ka_scan_kernel   one block per image: exclusive prefix of the tile counts inside every (level,
//                        anchor) plane, level totals -> pair_off, level_fg, mean lambda per level
//   KA3 ka_rows_kernel   every foreground prior finds its row arithmetically - its rank among the level's
//                        foreground priors in PRIOR order n = hw*A + a (the order of the reference's
//                        boolean index) - and materialises row, lambda, class key; rows == pairs, one
//                        pseudo-object, so K2 runs unchanged
//   KA4 ka_reduce_kernel one block per image: per (level, class) sum of epistemic / aleatoric and count in
//                        64-bit fixed point (2^-40 units: exact, so the result does not depend on the
//                        order of the atomics), then the class -> level aggregation.
#pragma once
#include "common.cuh"
#include "k1_alpha_topk.cuh"

namespace mehhua {

constexpr int kAllScanThreads = 256;
constexpr int kAllAggThreads = 512;
constexpr double kAllFixed = 1099511627776.0;      // 2^40

// foreground test of one prior.  ACT softmax: on return x[c] = exp(logit - max), inv = 1/sum; ACT relu: x[c] = relu(logit)
template <int C, int HEAD, int ACT>
__device__ __forceinline__ bool ka_prior_fg(const Plan& p, float (&x)[C > 0 ? C : 1], const float* __restrict__ src,
                                            const size_t stride, const int CC, float& inv) {
  if constexpr (ACT == MEHHUA_ACT_RELU) {
    float sum = 0.f, mx = 0.f;
    if constexpr (C > 0) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        x[c] = fmaxf(__ldg(src + (size_t)c * stride), 0.f);
        sum = __fadd_rn(sum, x[c]);
        mx = fmaxf(mx, x[c]);
      }
    } else {
      for (int c = 0; c < CC; ++c) {
        const float r = fmaxf(__ldg(src + (size_t)c * stride), 0.f);
        sum = __fadd_rn(sum, r);
        mx = fmaxf(mx, r);
      }
    }
    inv = 1.f;
    return __fdiv_rn(mx, __fadd_rn(sum, 1e-9f)) > p.fg_thr;      // max_c (r_c / (S + 1e-9)): the division is monotone
  } else {
    float den, pfg;
    if constexpr (C > 0) {
#pragma unroll
      for (int c = 0; c < C; ++c) x[c] = __ldg(src + (size_t)c * stride);
      softmax_regs<C, HEAD>(x, inv, den, pfg);
    } else {
      float m;
      softmax_stream<HEAD>(src, stride, CC, m, inv, den, pfg);
    }
    return pfg > p.fg_thr;
  }
}

template <int C, int HEAD, int ACT>
__global__ void __launch_bounds__(kKaThreads)
ka_fg_kernel(const __grid_constant__ Plan p, unsigned* __restrict__ fg_mask, int* __restrict__ tile_cnt,
             float* __restrict__ lam_part, unsigned* __restrict__ level_maxconf) {
  __shared__ float wsum[kKaThreads / 32];
  __shared__ int wcnt[kKaThreads / 32];
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
  const int hw = (lt - a * L.tpp) * kKaThreads + threadIdx.x;
  const bool live = hw < L.HW;
  const int CC = (C > 0) ? C : p.C;
  float lam = 0.f;
  bool fg = false;
  if (live) {
    float x[C > 0 ? C : 1];
    float inv;
    fg = ka_prior_fg<C, HEAD, ACT>(p, x, L.logits + ((size_t)(b * L.A + a) * CC) * L.HW + hw, (size_t)L.HW, CC, inv);
    lam = __ldg(L.lam + (size_t)(b * L.A + a) * L.HW + hw);
    if (ACT != MEHHUA_ACT_RELU && level_maxconf) level_maxconf_update(level_maxconf + b * p.S + s, inv);
  }
  __syncwarp();
  const unsigned word = __ballot_sync(0xffffffffu, fg);
  // lambda partial sum of the tile, fixed reduction tree
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lam += __shfl_xor_sync(0xffffffffu, lam, o);
  if ((threadIdx.x & 31) == 0) {
    wsum[threadIdx.x >> 5] = lam;
    wcnt[threadIdx.x >> 5] = __popc(word);
    fg_mask[((size_t)b * p.tiles_per_image + ti) * (kKaThreads / 32) + (threadIdx.x >> 5)] = word;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    int n = 0;
#pragma unroll
    for (int w = 0; w < kKaThreads / 32; ++w) { v += wsum[w]; n += wcnt[w]; }
    lam_part[(size_t)b * p.tiles_per_image + ti] = v;
    tile_cnt[(size_t)b * p.tiles_per_image + ti] = n;
  }
}

__global__ void __launch_bounds__(kAllScanThreads)
ka_scan_kernel(const __grid_constant__ Plan p, const int* __restrict__ tile_cnt, int* __restrict__ tile_pref,
               const float* __restrict__ lam_part, int* __restrict__ pair_off, int* __restrict__ level_fg,
               float* __restrict__ lam_mean, int* __restrict__ n_obj, int* __restrict__ n_det,
               unsigned* __restrict__ status) {
  __shared__ int lvl_total[kMaxLevels];
  const int b = blockIdx.x;
  if (threadIdx.x < kMaxLevels) lvl_total[threadIdx.x] = 0;
  __syncthreads();
  // one thread per (level, anchor) plane: exclusive prefix of the plane's tile counts
  int nplanes = 0;
  for (int s = 0; s < p.S; ++s) nplanes += p.lv[s].A;
  for (int pl = threadIdx.x; pl < nplanes; pl += kAllScanThreads) {
    int s = 0, a = pl;
    while (a >= p.lv[s].A) { a -= p.lv[s].A; ++s; }
    const LevelDev& L = p.lv[s];
    const size_t t0 = (size_t)b * p.tiles_per_image + L.tile0 + (size_t)a * L.tpp;
    int acc = 0;
    for (int i = 0; i < L.tpp; ++i) { tile_pref[t0 + i] = acc; acc += tile_cnt[t0 + i]; }
    atomicAdd(&lvl_total[s], acc);          // integer: order-independent
  }
  // mean lambda over ALL priors of each level: tile partials summed in tile order by one thread
  if (threadIdx.x >= 64 && threadIdx.x < 64 + p.S) {
    const int s = threadIdx.x - 64;
    const LevelDev& L = p.lv[s];
    const float* lp = lam_part + (size_t)b * p.tiles_per_image + L.tile0;
    float v = 0.f;
    for (int i = 0; i < L.tpp * L.A; ++i) v += lp[i];
    lam_mean[b * p.S + s] = __fdiv_rn(v, (float)L.n);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long acc = 0;
    for (int s = 0; s < p.S; ++s) {
      pair_off[b * (p.S + 1) + s] = (int)(acc < p.pair_cap ? acc : p.pair_cap);
      level_fg[b * p.S + s] = lvl_total[s] > 0;
      acc += lvl_total[s];
    }
    if (acc > p.pair_cap) { atomicOr(status, MEHHUA_ST_PAIR_OVERFLOW); acc = p.pair_cap; }
    pair_off[b * (p.S + 1) + p.S] = (int)acc;
    n_obj[b] = acc > 0 ? 1 : 0;
    n_det[b] = 0;
  }
}

template <int C, int HEAD, int ACT>
__global__ void __launch_bounds__(kKaThreads)
ka_rows_kernel(const __grid_constant__ Plan p, const unsigned* __restrict__ fg_mask, const int* __restrict__ tile_cnt,
               const int* __restrict__ tile_pref, const int* __restrict__ pair_off, float* __restrict__ score_rows,
               float* __restrict__ lam_rows, int* __restrict__ topk_idx, float* __restrict__ row_max,
               int* __restrict__ row_argmax, int* __restrict__ pair_row, int* __restrict__ pair_obj,
               int* __restrict__ pair_cls) {
  const int t = blockIdx.x;
  const int b = t / p.tiles_per_image;
  const int ti = t - b * p.tiles_per_image;
  if (tile_cnt[(size_t)b * p.tiles_per_image + ti] == 0) return;
  int s = 0;
#pragma unroll
  for (int i = 1; i < kMaxLevels; ++i)
    if (i < p.S && ti >= p.lv[i].tile0) s = i;
  const LevelDev& L = p.lv[s];
  const int lt = ti - L.tile0;
  const int a = lt / L.tpp;
  const int tl = lt - a * L.tpp;                       // tile of the plane
  const int hw = tl * kKaThreads + threadIdx.x;
  constexpr int W = kKaThreads / 32;
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned* masks = fg_mask + (size_t)b * p.tiles_per_image * W;
  if (!((masks[(size_t)ti * W + wi] >> lane) & 1u)) return;
  // rank among the level's foreground priors in prior order n = hw*A + a:
  //   sum over anchors a' of #{hw' < hw : fg(a', hw')}  +  #{a' < a : fg(a', hw)}
  long long q = pair_off[b * (p.S + 1) + s];
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int a2 = 0; a2 < L.A; ++a2) {
    const size_t t2 = (size_t)L.tile0 + (size_t)a2 * L.tpp + tl;
    q += tile_pref[(size_t)b * p.tiles_per_image + t2];
    for (int w = 0; w < wi; ++w) q += __popc(masks[t2 * W + w]);
    const unsigned word = masks[t2 * W + wi];
    q += __popc(word & lt_mask);
    if (a2 < a) q += (word >> lane) & 1u;
  }
  if (q >= p.pair_cap) return;                         // overflow was flagged by the scan kernel
  const int CC = (C > 0) ? C : p.C;
  const float* __restrict__ src = L.logits + ((size_t)(b * L.A + a) * CC) * L.HW + hw;
  float* srow = score_rows + ((size_t)b * p.row_stride + q) * CC;
  float best = -1.f;
  int arg = 0;
  if constexpr (C > 0) {
    float x[C];
    float inv;
    ka_prior_fg<C, HEAD, ACT>(p, x, src, (size_t)L.HW, CC, inv);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float pc = (ACT == MEHHUA_ACT_RELU) ? x[c] : __fmul_rn(x[c], inv);
      srow[c] = pc;
      if (pc > best) { best = pc; arg = c; }
    }
  } else {
    if (ACT == MEHHUA_ACT_RELU) {
      for (int c = 0; c < CC; ++c) {
        const float pc = fmaxf(__ldg(src + (size_t)c * L.HW), 0.f);
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
  }
  const size_t rq = (size_t)b * p.row_stride + q;
  lam_rows[rq] = __ldg(L.lam + (size_t)(b * L.A + a) * L.HW + hw);
  topk_idx[rq] = hw * L.A + a;
  row_max[rq] = best;
  row_argmax[rq] = arg;
  const size_t pq = (size_t)b * p.pair_cap + q;
  pair_row[pq] = (int)q;
  pair_obj[pq] = 0;
  pair_cls[pq] = arg;
}

// shared memory of the reduce kernel: per (level, class) two 64-bit sums and a count
__host__ __device__ inline size_t ka_reduce_smem_bytes(int S, int C) { return (size_t)S * C * (8 + 8 + 4 + 4) + 64; }

__global__ void __launch_bounds__(kAllAggThreads)
ka_reduce_kernel(const __grid_constant__ Plan p, const int* __restrict__ pair_cls, const int* __restrict__ pair_off,
                 const float* __restrict__ pair_unc, float* __restrict__ image_scores, float* __restrict__ group_unc) {
  extern __shared__ __align__(16) unsigned char ka_smem[];
  const int G = p.S * p.C;
  unsigned long long* sum_e = reinterpret_cast<unsigned long long*>(ka_smem);   // [S*C] epistemic, 2^-40 units (two's complement)
  unsigned long long* sum_a = sum_e + G;                                        // [S*C] aleatoric
  int* cnt = reinterpret_cast<int*>(sum_a + G);                                 // [S*C]
  float* mean_e = reinterpret_cast<float*>(cnt + G);                            // [S*C]
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < G; i += kAllAggThreads) { sum_e[i] = 0ull; sum_a[i] = 0ull; cnt[i] = 0; }
  __syncthreads();
  const int* poff = pair_off + b * (p.S + 1);
  const int np = poff[p.S];
  const int* pcls = pair_cls + (size_t)b * p.pair_cap;
  const float* punc = pair_unc + (size_t)b * p.pair_cap * 3;
  const bool pooled = p.agg_class == MEHHUA_AGG_POOL;
  for (int q = threadIdx.x; q < np; q += kAllAggThreads) {
    const int s = level_of_pair(poff, p.S, q);
    const int g = s * p.C + (pooled ? 0 : pcls[q]);
    atomicAdd(&sum_e[g], (unsigned long long)__double2ll_rn((double)punc[(size_t)q * 3 + 2] * kAllFixed));
    atomicAdd(&sum_a[g], (unsigned long long)__double2ll_rn((double)punc[(size_t)q * 3 + 1] * kAllFixed));
    atomicAdd(&cnt[g], 1);
  }
  __syncthreads();
  // group means
  for (int g = threadIdx.x; g < G; g += kAllAggThreads) {
    const int n = cnt[g];
    float me = 0.f, ma = 0.f;
    if (n > 0) {
      me = (float)((double)(long long)sum_e[g] / kAllFixed / (double)n);
      ma = (float)((double)(long long)sum_a[g] / kAllFixed / (double)n);
    }
    if (group_unc) {
      float* o = group_unc + ((size_t)b * G + g) * 3;
      o[0] = (float)n; o[1] = ma; o[2] = me;
    }
    mean_e[g] = me;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float lacc = 0.f;
    int ln = 0;
    for (int s = 0; s < p.S; ++s) {
      float cacc = (p.agg_class == MEHHUA_AGG_MAX) ? -FLT_MAX : 0.f;
      int cn = 0;
      for (int c = 0; c < (pooled ? 1 : p.C); ++c)
        if (cnt[s * p.C + c] > 0) { cacc = agg_combine(pooled ? MEHHUA_AGG_SUM : p.agg_class, cacc, mean_e[s * p.C + c]); ++cn; }
      if (cn == 0) continue;
      const float cv = pooled ? cacc : agg_finish(p.agg_class, cacc, cn);
      if (pooled && cv == 0.f) continue;              // AggregateAvgUnc keeps a level only `if sUncs` (Lambda_L2_ReLU.py:537)
      lacc = (ln == 0) ? cv : agg_combine(p.agg_scale, lacc, cv);
      ++ln;
    }
    // no foreground prior anywhere: 0 for the Entropy_ALL types (Lambda_L2.py:650-651), mean of an empty list
    // (NaN) for Entropy_Avg (Lambda_L2_ReLU.py:539)
    image_scores[b] = ln > 0 ? agg_finish(p.agg_scale, lacc, ln) : (pooled ? __int_as_float(0x7fc00000) : 0.f);
  }
}

}  // namespace mehhua
