// KM: mutual-information score of the ensemble / MC-dropout baselines.
//
// Reference semantics (mmdet/apis/CalEnsembleUnc.py:166-181 ComputeMI, three members;
// mmdet/apis/CalMCDropoutUnc.py:185-201 ComputeMCDropoutMI, n stochastic passes): per level s and
// image b, with p_m = sigmoid(logits_m) reshaped to [priors, nCls] (channel = a*nCls + c),
//   avg = mean_m p_m;  total = -sum_c avg ln avg;  ale = mean_m(-sum_c p_m ln p_m);  epi = total - ale
//   buffer[b, s] = mean over the level's priors of epi;   score[b] = mean_s buffer[b, s]
// One streaming pass over the M members' logits (HBM-bound for small M, XU-bound for M = 25): tile =
// 128 consecutive (h, w) positions of one (image, anchor) plane, thread = one prior, its classes read
// with stride H*W so every warp load is one coalesced 128-byte line of one class plane of one member.
// Per-tile sums land in the workspace and are added up in tile order by the finishing kernel
// (deterministic).  No special-casing of saturated sigmoids: p = 0 gives 0 * -inf = NaN exactly as in
// the reference.
#pragma once
#include "common.cuh"

namespace mehhua {

constexpr int kMiMaxMembers = 32;
constexpr int kMiThreads = 128;

struct MiPlan {
  const float* logits[kMiMaxMembers][kMaxLevels];   // member m, level s: [B, A*nCls, H, W]
  int HW[kMaxLevels], A[kMaxLevels], tpp[kMaxLevels], tile0[kMaxLevels];
  int S, B, M, n_cls, tiles_per_image;
};

__global__ void __launch_bounds__(kMiThreads)
km_mi_tiles_kernel(const __grid_constant__ MiPlan p, float* __restrict__ tile_sum) {
  const int t = blockIdx.x;
  const int b = t / p.tiles_per_image;
  const int ti = t - b * p.tiles_per_image;
  int s = 0;
#pragma unroll
  for (int i = 1; i < kMaxLevels; ++i)
    if (i < p.S && ti >= p.tile0[i]) s = i;
  const int lt = ti - p.tile0[s];
  const int a = lt / p.tpp[s];
  const int HW = p.HW[s];
  const int hw = (lt - a * p.tpp[s]) * kMiThreads + threadIdx.x;
  float epi = 0.f;
  if (hw < HW) {
    const size_t off = ((size_t)(b * p.A[s] + a) * p.n_cls) * HW + hw;
    const float fM = (float)p.M;
    float total = 0.f, ent = 0.f;
    for (int c = 0; c < p.n_cls; ++c) {
      float ps = 0.f;
      for (int m = 0; m < p.M; ++m) {
        const float x = __ldg(p.logits[m][s] + off + (size_t)c * HW);
        const float pr = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
        ps = __fadd_rn(ps, pr);
        ent = __fmaf_rn(-pr, logf(pr), ent);
      }
      const float avg = __fdiv_rn(ps, fM);
      total = __fmaf_rn(-avg, logf(avg), total);
    }
    epi = __fsub_rn(total, __fdiv_rn(ent, fM));
  }
  // block sum in a fixed order: lanes by shuffle, then the four warps in order
  __shared__ float ws[kMiThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) epi += __shfl_xor_sync(0xffffffffu, epi, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = epi;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = ws[0];
#pragma unroll
    for (int w = 1; w < kMiThreads / 32; ++w) v += ws[w];
    tile_sum[t] = v;
  }
}

// grid = B, block = 32 * S': warp s adds the level's tile sums in tile order (double), lane 0 of warp 0 takes the level mean
__global__ void km_mi_finish_kernel(const __grid_constant__ MiPlan p, const float* __restrict__ tile_sum,
                                    float* __restrict__ level_mi, float* __restrict__ image_scores) {
  __shared__ float lvl[kMaxLevels];
  const int b = blockIdx.x, s = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (s < p.S) {
    const int nt = p.tpp[s] * p.A[s];
    const float* src = tile_sum + (size_t)b * p.tiles_per_image + p.tile0[s];
    double acc = 0.0;
    for (int i = lane; i < nt; i += 32) acc += (double)src[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const float v = (float)(acc / ((double)p.HW[s] * p.A[s]));
      lvl[s] = v;
      if (level_mi) level_mi[b * p.S + s] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int i = 0; i < p.S; ++i) v = __fadd_rn(v, lvl[i]);
    image_scores[b] = __fdiv_rn(v, (float)p.S);
  }
}

}  // namespace mehhua
