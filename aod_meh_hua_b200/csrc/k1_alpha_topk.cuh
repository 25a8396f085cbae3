// K1: logits -> softmax / score -> ranking key (K1a, the HBM-streaming kernel) -> per-level top-k
// (K1b, block radix select) -> the kept rows, lambda, decoded boxes, NMS candidates (K1c).
// Where do the kept rows' scores come from?  Three forms, chosen per level by the plan (build_plan in mehhua.cu):
//   capture : sparse top-k levels (n >= 8 k, or a level that is a small part of the image).  K1t reads 1/32 of the
//             level's warps and estimates the key of rank 1.75 k; K1a parks the record of every prior at or above it
//             while the logits are in registers (C exponentials + the two normalisers, one 16-byte aligned row); K1b
//             sorts the parked composites only; K1c (k1c_parked_kernel, one warp per row) turns the records into
//             score rows.  An (image, level) whose estimate misses falls back to the select over all keys + gather.
//   gather  : the row's C logits re-read with stride H*W (one 32-byte sector per logit): top-k levels that are not
//             captured, and fall-backs.
//   rescan  : levels without a top-k, or with 2 k >= n: a second coalesced pass.
//
// Reference semantics: _get_bboxes per-level block (mmdet/models/dense_heads/Lambda_L2.py:264-326,
// My_L_ssd_head.py:325-361), delta2bbox (core/bbox/coder/delta_xywh_bbox_coder.py:205-267), the
// level-FG test of ComputeObjUnc (Lambda_L2.py:496-502) and the score filter of multiclass_nms
// (core/post_processing/bbox_nms.py:41-66).
#pragma once
#include "common.cuh"

namespace mehhua {

#ifndef MEHHUA_K1A_THREADS
#define MEHHUA_K1A_THREADS 32
#endif
// K1a / KA tile = one warp: a block retires as soon as its warp has parked its rows - with four-warp blocks nine of ten
// blocks waited on some warp's slot atomic (measured: K1a 4.79 -> 4.64 ms per 437 cfg-3 images)
constexpr int kK1aThreads = MEHHUA_K1A_THREADS;   // one prior position per thread, C logits in registers
constexpr int kRescanThreads = 128;               // K1c rescan tile
constexpr int kKaThreads = 128;                   // tile of the Entropy_ALL kernels (their plans are built with it)
constexpr int kSelThreads = 1024;
constexpr int kSelCap = 4096;      // >= MEHHUA_MAX_NMS_PRE
constexpr size_t kSelSmem = kSelCap * 8 + 4096 * 4 + 40 * 4;
constexpr int kGatherThreads = 128;

// Softmax of one prior held in registers.  On return x[c] = exp(logit_c - max) (unnormalised),
// inv = 1/sum, den = 1 / ((sum_c p_c + 1e-20) + 1e-9), the reciprocal of the Retina score denominator
// (1 for SSD) and pfg = max foreground softmax probability.  p_c = x[c]*inv, score_c = p_c*den
// (Retina; within 1 ulp of the reference's division) or p_c.
// Summation is sequential in class order with explicitly rounded ops, so K1a (key) and K1c (row)
// produce bit-identical values for the same prior.
// Internal third head form: the Retina head with the base class's evidential scores (MEHHUA_ACT_RELU_PLUS_ONE,
// L_anchor_head.py:401-406): alpha = relu(logit) + 1, S = sum alpha + 1e-20, score = alpha / S.  In this form
// x[c] = alpha_c, `inv` carries S (the divisor, not a reciprocal) and den = 1.
constexpr int kHeadRpo = 2;

// score of one class from what softmax_regs / softmax_stream left behind
template <int HEAD>
__device__ __forceinline__ float k1_score(const float xc, const float inv, const float den) {
  if (HEAD == kHeadRpo) return __fdiv_rn(xc, inv);
  const float pc = __fmul_rn(xc, inv);
  return (HEAD == MEHHUA_HEAD_RETINA) ? __fmul_rn(pc, den) : pc;
}

template <int C, int HEAD>
__device__ __forceinline__ void softmax_regs(float (&x)[C], float& inv, float& den, float& pfg) {
  if constexpr (HEAD == kHeadRpo) {
    float sum = 0.f, mx = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      x[c] = __fadd_rn(fmaxf(x[c], 0.f), 1.f);
      sum = __fadd_rn(sum, x[c]);
      mx = fmaxf(mx, x[c]);
    }
    inv = __fadd_rn(sum, 1e-20f);
    den = 1.f;
    pfg = __fdiv_rn(mx, inv);
    return;
  }
  constexpr int CF = (HEAD == MEHHUA_HEAD_SSD) ? C - 1 : C;
  float mfg = x[0];
#pragma unroll
  for (int c = 1; c < CF; ++c) mfg = fmaxf(mfg, x[c]);
  const float m = (HEAD == MEHHUA_HEAD_SSD) ? fmaxf(mfg, x[C - 1]) : mfg;
  const float nml2 = -__fmul_rn(m, kLog2e);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    x[c] = ex2_approx(fmaf(x[c], kLog2e, nml2));
    sum = __fadd_rn(sum, x[c]);
  }
  inv = __fdiv_rn(1.f, sum);
  const float efg = ex2_approx(fmaf(mfg, kLog2e, nml2));
  pfg = __fmul_rn(efg, inv);
  if (HEAD == MEHHUA_HEAD_RETINA) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) s = __fadd_rn(s, __fmul_rn(x[c], inv));
    den = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(s, 1e-20f), 1e-9f));
  } else {
    den = 1.f;
  }
}

// Same arithmetic for a class count that is not instantiated: three strided passes over memory.
template <int HEAD>
__device__ __forceinline__ void softmax_stream(const float* __restrict__ src, size_t stride, int C,
                                               float& m_out, float& inv, float& den, float& pfg) {
  if constexpr (HEAD == kHeadRpo) {
    float sum = 0.f, mx = 0.f;
    for (int c = 0; c < C; ++c) {
      const float a = __fadd_rn(fmaxf(__ldg(src + c * stride), 0.f), 1.f);
      sum = __fadd_rn(sum, a);
      mx = fmaxf(mx, a);
    }
    inv = __fadd_rn(sum, 1e-20f);
    den = 1.f;
    pfg = __fdiv_rn(mx, inv);
    m_out = 0.f;
    return;
  }
  const int CF = (HEAD == MEHHUA_HEAD_SSD) ? C - 1 : C;
  float mfg = __ldg(src);
  for (int c = 1; c < CF; ++c) mfg = fmaxf(mfg, __ldg(src + c * stride));
  const float m = (HEAD == MEHHUA_HEAD_SSD) ? fmaxf(mfg, __ldg(src + (size_t)(C - 1) * stride)) : mfg;
  const float nml2 = -__fmul_rn(m, kLog2e);
  float sum = 0.f;
  for (int c = 0; c < C; ++c) sum = __fadd_rn(sum, ex2_approx(fmaf(__ldg(src + c * stride), kLog2e, nml2)));
  inv = __fdiv_rn(1.f, sum);
  pfg = __fmul_rn(ex2_approx(fmaf(mfg, kLog2e, nml2)), inv);
  if (HEAD == MEHHUA_HEAD_RETINA) {
    float s = 0.f;
    for (int c = 0; c < C; ++c)
      s = __fadd_rn(s, __fmul_rn(ex2_approx(fmaf(__ldg(src + c * stride), kLog2e, nml2)), inv));
    den = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(s, 1e-20f), 1e-9f));
  } else {
    den = 1.f;
  }
  m_out = m;
}

// getMaxConf (mmdet/utils/functions.py:467-476): the largest softmax probability of a prior is
// exp(0) * inv = inv (the max-logit class, background included for SSD).  Positive floats order
// like their bit patterns, so the level maximum is one warp REDUX and an atomicMax by one lane -
// issued only when the (possibly stale, hence never too large) cached value would be raised, which
// keeps the same-address atomic traffic to a few updates per (image, level).
__device__ __forceinline__ void level_maxconf_update(unsigned* dst, float inv) {
  const unsigned m = __activemask();
  const unsigned v = __reduce_max_sync(m, __float_as_uint(inv));
  if ((threadIdx.x & 31) == (__ffs(m) - 1) && v > *dst) atomicMax(dst, v);
}

// ------------------------------------------------------------------------------------------
// K1a: stream every logit once.  Tile = 128 consecutive (h,w) positions of one (image, anchor)
// plane; thread t owns position hw and reads its C class logits with stride H*W, so each warp
// load instruction is one coalesced 128-byte line of one class plane.
// C == 0 selects the generic (runtime class count) path.
// ------------------------------------------------------------------------------------------
// The ranking composite of prior position j (= a*HW + hw) with key `key`: exact key ties go to the lower
// position.  Keys outside [0, 2) (negatives, NaN) are clamped so that composites stay below 2^62.
__device__ __forceinline__ unsigned long long k1_composite(const float key, const int j) {
  unsigned kb = __float_as_uint(key);
  if (kb > 0x3fffffffu) kb = (kb & 0x80000000u) ? 0u : 0x3fffffffu;   // negatives / >= 2.0 / NaN
  return ((unsigned long long)kb << 32) | (unsigned long long)(0xffffffffu - (unsigned)j);
}

// key of one prior from its logits; on return x[c] = score_c (C > 0)
template <int C, int HEAD>
__device__ __forceinline__ float k1_key(float (&x)[C > 0 ? C : 1], const float* __restrict__ src, const size_t stride,
                                        const int CC, float& inv, float& pfg, float& den) {
  if constexpr (C > 0) {
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = __ldg(src + (size_t)c * stride);
    softmax_regs<C, HEAD>(x, inv, den, pfg);
  } else {
    float m;
    softmax_stream<HEAD>(src, stride, CC, m, inv, den, pfg);
  }
  return (HEAD == MEHHUA_HEAD_RETINA) ? __fmul_rn(pfg, den) : pfg;
}

template <int C, int HEAD>
__global__ void __launch_bounds__(kK1aThreads)
k1a_keys_kernel(const __grid_constant__ Plan p, float* __restrict__ keys, int* __restrict__ level_fg,
                unsigned* __restrict__ level_maxconf, const float* __restrict__ tau, int* __restrict__ cap_cnt,
                unsigned long long* __restrict__ cap_comp, float* __restrict__ cap_scores) {
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
  const bool live = hw < L.HW;        // lanes past the end of the plane stay for the warp-wide capture step
  const int CC = (C > 0) ? C : p.C;
  float x[C > 0 ? C : 1];
  float inv = 0.f, den = 0.f, pfg = 0.f, key = 0.f;
  if (live) {
    const float* __restrict__ src = L.logits + ((size_t)(b * L.A + a) * CC) * L.HW + hw;
    key = k1_key<C, HEAD>(x, src, (size_t)L.HW, CC, inv, pfg, den);
    keys[(size_t)b * p.N + L.n_off + a * L.HW + hw] = key;
    if (pfg > p.fg_thr) level_fg[b * p.S + s] = 1;
    if (HEAD != kHeadRpo && level_maxconf) level_maxconf_update(level_maxconf + b * p.S + s, inv);
  }
  if constexpr (C > 0) {
    // capture: park the row of every prior at or above the level's threshold (about 1.75 k of them per
    // (image, level)): its C exponentials as they sit in registers (128-bit stores, one contiguous
    // 4*C-byte row per lane) and the two normalisers; K1c turns them into scores with the same two
    // multiplications it applies to a gathered row.
    if (L.cap < 0 || tau == nullptr) return;             // block-uniform
    const unsigned full = 0xffffffffu;
    __syncwarp(full);                                     // reconverge after the divergent updates above
    const bool park = live && key >= __ldg(tau + b * p.S + s);
    const unsigned pm = __ballot_sync(full, park);
    if (pm != 0u) {
      const int lane = threadIdx.x & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(cap_cnt + b * p.S + s, __popc(pm));
      base = __shfl_sync(full, base, 0);
      const int slot = base + __popc(pm & ((1u << lane) - 1u));
      if (park && slot < kCapRows) {
        const size_t e = ((size_t)b * p.n_cap_levels + L.cap) * kCapRows + slot;
        cap_comp[e] = k1_composite(key, a * L.HW + hw);
        float* dst = cap_scores + e * cap_row_floats(C);
        if constexpr (C % 4 == 0) {
#pragma unroll
          for (int c = 0; c < C; c += 4) __stcs(reinterpret_cast<float4*>(dst + c), make_float4(x[c], x[c + 1], x[c + 2], x[c + 3]));
          __stcs(reinterpret_cast<float4*>(dst + C), make_float4(inv, den, 0.f, 0.f));
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) __stcs(dst + c, x[c]);
          __stcs(dst + C, inv);
          __stcs(dst + C + 1, den);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// K1t: capture threshold of every capture level of every image.  grid = (S, B), one block reads the
// logits of 1/stride of the level's warps (whole 128-byte lines per class, as K1a does), computes their
// keys and takes the key of rank ~ kCapTarget * k * (sampled / N) among them.
// ------------------------------------------------------------------------------------------
constexpr int kThrThreads = 512;     // C logits live in registers: 128 registers per thread
constexpr size_t kThrSmem = (size_t)kCapSampleMax * 4 + (size_t)kSelCap * 8 + 4096 * 4 + 40 * 4;

template <int C, int HEAD>
__global__ void __launch_bounds__(kThrThreads)
k1t_threshold_kernel(const __grid_constant__ Plan p, float* __restrict__ tau, unsigned* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char k1t_smem[];
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(k1t_smem);   // kSelCap
  float* skeys = reinterpret_cast<float*>(buf + kSelCap);                       // kCapSampleMax
  int* hist = reinterpret_cast<int*>(skeys + kCapSampleMax);                    // 4096
  int* sh = hist + 4096;                                                        // 40
  const int s = blockIdx.x, b = blockIdx.y;
  const LevelDev& L = p.lv[s];
  if (L.cap < 0) return;
  const int CC = (C > 0) ? C : p.C;
  const int wpp = (L.HW + 31) >> 5;                       // warps per (image, anchor) plane
  const int total = wpp * L.A;
  int stride = kCapStride;
  while ((total + stride - 1) / stride > kCapSampleMax / 32) stride <<= 1;
  while (stride > 1 && total / stride < kCapMinWarps) stride >>= 1;      // small levels: at least kCapMinWarps sampled warps, spread over the planes
  const int phase = (b * 7 + s * 3) % stride;
  const int nsw = total > phase ? (total - phase + stride - 1) / stride : 0;   // sampled warps
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x >> 5; i < nsw; i += kThrThreads / 32) {
    const int g = phase + i * stride;
    const int a = g / wpp, hw = ((g - a * wpp) << 5) + lane;
    float key = -1.f;                                     // not an element
    if (hw < L.HW) {
      float x[C > 0 ? C : 1];
      float inv, den, pfg;
      key = k1_key<C, HEAD>(x, L.logits + ((size_t)(b * L.A + a) * CC) * L.HW + hw, (size_t)L.HW, CC, inv, pfg, den);
    }
    skeys[i * 32 + lane] = key;
  }
  __syncthreads();
  const int n = nsw * 32;
  int valid = 0;     // sampled priors: every sampled warp is full except the last one of a plane
  for (int i = threadIdx.x; i < n; i += kThrThreads) valid += skeys[i] >= 0.f ? 1 : 0;
  valid = block_incl_scan<kThrThreads>(valid, sh);
  if (threadIdx.x == kThrThreads - 1) sh[39] = valid;
  __syncthreads();
  valid = sh[39];
  __syncthreads();
  const int r = (int)(((long long)kCapTargetNum * L.k * valid + (long long)kCapTargetDen * L.n - 1) / ((long long)kCapTargetDen * L.n)) + 1;
  auto get = [&](int j) -> unsigned long long { return skeys[j] >= 0.f ? k1_composite(skeys[j], j) : 0ull; };
  const int cnt = block_collect_topk<kThrThreads, kSelCap, 0>(get, n, min(r, kSelCap), ~0ull, buf, hist, sh, status);
  if (threadIdx.x == 0) {
    // fewer sampled priors than the rank asked for: park everything (the level then falls back if that is too many)
    tau[b * p.S + s] = (cnt >= r) ? __uint_as_float((unsigned)(buf[r - 1] >> 32)) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// K1b: per (image, level) top-k of the keys, sorted descending (exact ties: lower anchor, then lower position).
// grid = (S, B); levels without an active top-k return immediately.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSelThreads)
k1b_select_kernel(const __grid_constant__ Plan p, const float* __restrict__ keys,
                  int* __restrict__ topk_idx, int* __restrict__ inv_map, const int* __restrict__ cap_cnt,
                  const unsigned long long* __restrict__ cap_comp, int* __restrict__ row_slot,
                  unsigned* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char k1b_smem[];
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(k1b_smem);   // kSelCap
  int* hist = reinterpret_cast<int*>(buf + kSelCap);                            // 4096
  int* sh = hist + 4096;                                                        // 40
  const int s = blockIdx.x, b = blockIdx.y;
  const LevelDev& L = p.lv[s];
  if (!L.topk) return;
  int* out = topk_idx + (size_t)b * p.K + L.k_off;
  if (L.cap >= 0) {
    // capture level: the k best are among the parked rows when at least k were parked and none was dropped
    const int nc = cap_cnt[b * p.S + s];
    int* rs = row_slot + (size_t)b * p.K + L.k_off;
    if (nc >= L.k && nc <= kCapRows) {
      const unsigned long long* cc = cap_comp + ((size_t)b * p.n_cap_levels + L.cap) * kCapRows;
      int n2 = 1;
      while (n2 < nc) n2 <<= 1;
      for (int i = threadIdx.x; i < n2; i += kSelThreads) buf[i] = i < nc ? cc[i] : 0ull;
      __syncthreads();
      block_bitonic_desc<kSelThreads>(buf, n2);
      for (int i = threadIdx.x; i < L.k; i += kSelThreads) {
        const int j = (int)(0xffffffffu - (unsigned)(buf[i] & 0xffffffffull));
        const int a = j / L.HW;
        out[i] = (j - a * L.HW) * L.A + a;
      }
      // where did each parked record end up?  composites are distinct: its rank is found by bisection (slot = position in cc)
      for (int i = threadIdx.x; i < nc; i += kSelThreads) {
        const unsigned long long mine = cc[i];
        int lo = 0, hi = n2;         // buf[lo] >= mine > buf[hi] (descending)
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (buf[mid] >= mine) lo = mid; else hi = mid; }
        if (lo < L.k) rs[lo] = i;
      }
      return;
    }
    for (int i = threadIdx.x; i < L.k; i += kSelThreads) rs[i] = -1;      // fall back: select over all keys, rows by gather
    if (threadIdx.x == 0) atomicOr(status, MEHHUA_ST_CAPTURE_FALLBACK);
  }
  const float* kp = keys + (size_t)b * p.N + L.n_off;
  // composite = key bits << 32 | ~position, position j = a*HW + hw (the key array is anchor-major):
  // exact key ties go to the lower position.  The prior index n = hw*A + a is only rebuilt for
  // the k winners.
  auto get = [&](int j) -> unsigned long long { return k1_composite(__ldcg(kp + j), j); };
  const int cnt = block_collect_topk<kSelThreads, kSelCap, 0>(get, L.n, L.k, ~0ull, buf, hist, sh, status);
  const int k = min(L.k, cnt);
  int* inv = inv_map + (size_t)b * p.N + L.n_off;
  if (L.rescan) {       // dense level: the rescan kernel finds rows through the inverse map
    for (int j = threadIdx.x; j < L.n; j += kSelThreads) inv[j] = -1;
    __syncthreads();
  }
  for (int i = threadIdx.x; i < k; i += kSelThreads) {
    const int j = (int)(0xffffffffu - (unsigned)(buf[i] & 0xffffffffull));
    const int a = j / L.HW;
    out[i] = (j - a * L.HW) * L.A + a;
    if (L.rescan) inv[j] = L.k_off + i;
  }
}

// ------------------------------------------------------------------------------------------
// K1c: one thread per kept row: recompute its softmax row (bit-identical to K1a), write scores,
// lambda, decoded box, row max / argmax and append its NMS candidates (score > score_thr).
// grid = (ceil(K / 128), B).
// ------------------------------------------------------------------------------------------
// The C logits of prior (b, a, hw) of level L into registers (stride H*W between classes).
template <int C>
__device__ __forceinline__ void k1_load_logits(const LevelDev& L, const int b, const int a, const int hw,
                                               float (&x)[C > 0 ? C : 1]) {
  if constexpr (C > 0) {
    const float* __restrict__ src = L.logits + ((size_t)(b * L.A + a) * C) * L.HW + hw;
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = __ldg(src + (size_t)c * L.HW);
  }
}

// delta2bbox of prior n = (hw, a) of level L (core/bbox/coder/delta_xywh_bbox_coder.py:205-267): denormalise, clamp
// dw/dh, exp, centre/size -> corners, clip to the image, divide by the scale factor; explicitly rounded fp32 ops.
__device__ __forceinline__ float4 k1_decode_box(const Plan& p, const LevelDev& L, const int b, const int n, const int a,
                                                const int hw, const float* __restrict__ img_shapes,
                                                const float* __restrict__ scale_factors) {
  float4 box;
  const float* __restrict__ dp = L.deltas + ((size_t)(b * L.A + a) * 4) * L.HW + hw;
  const float4 an = __ldg(reinterpret_cast<const float4*>(L.anchors) + n);
  const float dx = __fadd_rn(__fmul_rn(__ldg(dp), p.stds[0]), p.means[0]);
  const float dy = __fadd_rn(__fmul_rn(__ldg(dp + (size_t)L.HW), p.stds[1]), p.means[1]);
  float dw = __fadd_rn(__fmul_rn(__ldg(dp + 2 * (size_t)L.HW), p.stds[2]), p.means[2]);
  float dh = __fadd_rn(__fmul_rn(__ldg(dp + 3 * (size_t)L.HW), p.stds[3]), p.means[3]);
  const float px = __fmul_rn(__fadd_rn(an.x, an.z), 0.5f);
  const float py = __fmul_rn(__fadd_rn(an.y, an.w), 0.5f);
  const float pw = __fsub_rn(an.z, an.x);
  const float ph = __fsub_rn(an.w, an.y);
  const float dxw = __fmul_rn(pw, dx);
  const float dyh = __fmul_rn(ph, dy);
  dw = fminf(fmaxf(dw, -p.max_ratio), p.max_ratio);
  dh = fminf(fmaxf(dh, -p.max_ratio), p.max_ratio);
  const float gw = __fmul_rn(pw, expf(dw));
  const float gh = __fmul_rn(ph, expf(dh));
  const float gx = __fadd_rn(px, dxw);
  const float gy = __fadd_rn(py, dyh);
  const float hgw = __fmul_rn(gw, 0.5f), hgh = __fmul_rn(gh, 0.5f);
  box.x = __fsub_rn(gx, hgw);
  box.y = __fsub_rn(gy, hgh);
  box.z = __fadd_rn(gx, hgw);
  box.w = __fadd_rn(gy, hgh);
  const float imh = __ldg(img_shapes + 2 * b), imw = __ldg(img_shapes + 2 * b + 1);
  box.x = box.x < 0.f ? 0.f : box.x;  box.y = box.y < 0.f ? 0.f : box.y;
  box.z = box.z < 0.f ? 0.f : box.z;  box.w = box.w < 0.f ? 0.f : box.w;
  box.x = box.x > imw ? imw : box.x;  box.y = box.y > imh ? imh : box.y;
  box.z = box.z > imw ? imw : box.z;  box.w = box.w > imh ? imh : box.w;
  if (p.rescale) {
    const float4 sf = __ldg(reinterpret_cast<const float4*>(scale_factors) + b);
    box.x = __fdiv_rn(box.x, sf.x);  box.y = __fdiv_rn(box.y, sf.y);
    box.z = __fdiv_rn(box.z, sf.z);  box.w = __fdiv_rn(box.w, sf.w);
  }
  return box;
}

// What K1c produces for one kept row r = prior n of level L (used by the gather and rescan kernels):
// softmax row (bit-identical to K1a's arithmetic), score row, lambda, decoded box, row max /
// argmax, and the number of NMS candidates of the row.
template <int C, int HEAD>
__device__ __forceinline__ void k1_row_body(const Plan& p, const LevelDev& L, const int b, const int r, const int n,
                                            const float* __restrict__ img_shapes, const float* __restrict__ scale_factors,
                                            float* __restrict__ score_rows, float* __restrict__ lam_rows,
                                            float* __restrict__ boxes, float* __restrict__ row_max,
                                            int* __restrict__ row_argmax, float (&x)[C > 0 ? C : 1],
                                            float* tile_row, int& ncand, float& bmax, float*& srow,
                                            const bool scores_ready = false) {
  const int CC = (C > 0) ? C : p.C;
  const int NF = p.num_fg;
  float4 box;
    const int hw = n / L.A, a = n - hw * L.A;
    const float* __restrict__ src = L.logits + ((size_t)(b * L.A + a) * CC) * L.HW + hw;
    srow = score_rows + ((size_t)b * p.K + r) * CC;
    float best = -1.f;
    int arg = 0;
    if constexpr (C > 0) {      // x[] holds the row's logits (or, scores_ready, its parked scores), loaded by the caller
      float inv = 1.f, den = 1.f, pfg;
      if (!scores_ready) softmax_regs<C, HEAD>(x, inv, den, pfg);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float sc = x[c];
        if (!scores_ready) sc = k1_score<HEAD>(x[c], inv, den);
        x[c] = sc;
        if (sc > best) { best = sc; arg = c; }
        if (c < NF && sc > p.score_thr) ++ncand;
      }
#pragma unroll
      for (int c = 0; c < C; ++c) tile_row[c] = x[c];     // staged in shared memory, flushed coalesced
    } else {
      float m, inv, den, pfg;
      softmax_stream<HEAD>(src, (size_t)L.HW, CC, m, inv, den, pfg);
      const float nml2 = -__fmul_rn(m, kLog2e);
      for (int c = 0; c < CC; ++c) {
        const float lg = __ldg(src + (size_t)c * L.HW);
        const float e = (HEAD == kHeadRpo) ? __fadd_rn(fmaxf(lg, 0.f), 1.f) : ex2_approx(fmaf(lg, kLog2e, nml2));
        const float sc = k1_score<HEAD>(e, inv, den);
        tile_row[c] = sc;
        if (sc > best) { best = sc; arg = c; }
        if (c < NF && sc > p.score_thr) ++ncand;
      }
    }
    row_max[(size_t)b * p.K + r] = best;
    row_argmax[(size_t)b * p.K + r] = arg;
    lam_rows[(size_t)b * p.K + r] = __ldg(L.lam + (size_t)(b * L.A + a) * L.HW + hw);

    box = k1_decode_box(p, L, b, n, a, hw, img_shapes, scale_factors);
    reinterpret_cast<float4*>(boxes)[(size_t)b * p.K + r] = box;
    bmax = fmaxf(fmaxf(box.x, box.y), fmaxf(box.z, box.w));
}

// warp-aggregated append of each lane's NMS candidates (score > score_thr): one atomic per warp
__device__ __forceinline__ void k1_append_candidates(const Plan& p, const int b, const int r, const int ncand,
                                                     const float bmax, const float* srow /* this lane's scores */,
                                                     unsigned long long* __restrict__ cand, int* __restrict__ cand_cnt,
                                                     unsigned* __restrict__ cand_maxc) {
  const int NF = p.num_fg;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int incl = ncand;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(full, incl, o);
    if (lane >= o) incl += t;
  }
  const int total = __shfl_sync(full, incl, 31);
  if (total == 0) return;
  int base = 0;
  if (lane == 31) base = atomicAdd(cand_cnt + b, total);
  base = __shfl_sync(full, base, 31);
  float wmax = (ncand > 0) ? bmax : -FLT_MAX;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(full, wmax, o));
  if (lane == 0) atomicMax(cand_maxc + b, f2ord(wmax));
  if (ncand > 0) {
    unsigned long long* dst = cand + (size_t)b * p.K * NF + base + (incl - ncand);
    for (int c = 0; c < NF; ++c) {
      const float sc = srow[c];   // written by this thread above
      if (sc > p.score_thr)
        *dst++ = ((unsigned long long)__float_as_uint(sc) << 32) |
                 (unsigned long long)(0xffffffffu - (unsigned)(r * NF + c));
    }
  }
}

// Row staging: a lane's C scores go to its row of a shared-memory tile (odd stride: conflict-free),
// then the warp writes the 32 rows out with coalesced stores (3 instructions per 80-float row
// instead of 80 instructions x 32 scattered sectors).
template <int C> struct K1Tile { static constexpr int stride = (C % 2 == 0) ? C + 1 : C + 2; };

template <int C>
__device__ __forceinline__ void k1_flush_rows(const float* tile /* warp's [32][stride] */, float* srow) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned long long mine = reinterpret_cast<unsigned long long>(srow);
  __syncwarp();
  // only the lanes that hold a row (all of them in the gather kernel, a few in the rescan kernel)
  for (unsigned m = __ballot_sync(full, mine != 0ull); m != 0u; m &= m - 1u) {
    const int i = __ffs(m) - 1;
    float* dst = reinterpret_cast<float*>(__shfl_sync(full, mine, i));
    const float* src = tile + i * K1Tile<C>::stride;
    for (int c = lane; c < C; c += 32) dst[c] = src[c];
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// K1c (gather form): one thread per kept row of the SPARSE top-k levels (k << N): re-reads the
// row's C logits with stride H*W (sector-granular traffic, 8*k*C*4 bytes per level).
// grid = (ceil(K / 128), B); rows of dense / direct levels are left to the rescan kernel.
// ------------------------------------------------------------------------------------------
template <int C, int HEAD>
__global__ void __launch_bounds__(kGatherThreads, 3)
k1c_gather_kernel(const __grid_constant__ Plan p, const float* __restrict__ img_shapes,
                  const float* __restrict__ scale_factors, const int* __restrict__ topk_idx,
                  float* __restrict__ score_rows, float* __restrict__ lam_rows,
                  float* __restrict__ boxes, float* __restrict__ row_max, int* __restrict__ row_argmax,
                  unsigned long long* __restrict__ cand, int* __restrict__ cand_cnt,
                  unsigned* __restrict__ cand_maxc, const int* __restrict__ row_slot,
                  const float* __restrict__ cap_scores, const int skip_parked) {
  const int b = blockIdx.y;
  const int r = blockIdx.x * kGatherThreads + threadIdx.x;
  int ncand = 0;
  float bmax = 0.f;
  float* srow = nullptr;
  __shared__ float tile[(C > 0) ? (kGatherThreads * K1Tile<C>::stride) : 1];
  float* tile_row = (C > 0) ? tile + threadIdx.x * K1Tile<C>::stride : nullptr;
  if (r < p.K) {
    const LevelDev& L = p.lv[level_of_row(p, r)];
    const int slot0 = (C > 0 && L.rescan == 0 && L.cap >= 0) ? row_slot[(size_t)b * p.K + r] : -1;
    if (L.rescan == 0 && !(skip_parked && slot0 >= 0)) {      // parked rows: k1c_parked_kernel's, when it runs
      const int n = topk_idx[(size_t)b * p.K + r];
      const int hw = n / L.A;
      float x[C > 0 ? C : 1];
      bool parked = false;
      if constexpr (C > 0) {
        const int slot = slot0;
        if (slot >= 0) {        // capture level: the row was parked by K1a (one contiguous read): exponentials + normalisers
          const float* src = cap_scores + (((size_t)b * p.n_cap_levels + L.cap) * kCapRows + slot) * cap_row_floats(C);
          float inv, den;
          if constexpr (C % 4 == 0) {
#pragma unroll
            for (int c = 0; c < C; c += 4) {
              const float4 v = __ldcs(reinterpret_cast<const float4*>(src + c));
              x[c] = v.x; x[c + 1] = v.y; x[c + 2] = v.z; x[c + 3] = v.w;
            }
            const float4 nd = __ldcs(reinterpret_cast<const float4*>(src + C));
            inv = nd.x; den = nd.y;
          } else {
#pragma unroll
            for (int c = 0; c < C; ++c) x[c] = __ldcs(src + c);
            inv = __ldcs(src + C); den = __ldcs(src + C + 1);
          }
#pragma unroll
          for (int c = 0; c < C; ++c) x[c] = k1_score<HEAD>(x[c], inv, den);
          parked = true;
        }
      }
      if (!parked) k1_load_logits<C>(L, b, n - hw * L.A, hw, x);
      if (C == 0) tile_row = score_rows + ((size_t)b * p.K + r) * p.C;
      k1_row_body<C, HEAD>(p, L, b, r, n, img_shapes, scale_factors, score_rows, lam_rows, boxes, row_max,
                           row_argmax, x, tile_row, ncand, bmax, srow, parked);
    }
  }
  if constexpr (C > 0) k1_flush_rows<C>(tile + (threadIdx.x & ~31) * K1Tile<C>::stride, srow);
  k1_append_candidates(p, b, r, ncand, bmax, tile_row, cand, cand_cnt, cand_maxc);
}

// ------------------------------------------------------------------------------------------
// K1c (parked form): the kept rows of the capture levels.  K1a left every such row as one contiguous,
// 16-byte aligned record in the workspace (C exponentials, 1/sum, score normaliser), so nothing is
// recomputed from the logits.  One warp owns 32 consecutive rows of an image:
//   stage : lane i issues ONE bulk-async copy (cp.async.bulk, the TMA engine's linear form) of its row's
//           record into the warp's shared-memory slab; all 32 copies are in flight at once and signal a
//           per-warp mbarrier with their byte count - no registers are tied up while the rows travel;
//   rows  : the warp walks its rows one at a time, lane = class (c = lane + 32 j): scores with the same
//           two multiplications K1a's composite was built from (row_max == key bit for bit), coalesced
//           store of the score row, max / first argmax by two warp REDUX, NMS candidates appended to a
//           per-warp queue in shared memory (one atomic per flush, not per row);
//   boxes : lane = row again: lambda, delta2bbox, row max / argmax, one coalesced store each.
// Rows that were not parked (a level whose capture estimate missed) are left to k1c_gather_kernel.
// grid = (ceil(K / (32 * kParkWarps)), B).
// ------------------------------------------------------------------------------------------
constexpr int kParkWarps = 4;
constexpr int kParkQueue = 128;       // NMS candidates a warp queues between two flushes
__host__ __device__ inline size_t k1c_parked_smem(int C, bool bulk) {
  return (size_t)kParkWarps * ((bulk ? 32 * (size_t)cap_row_floats(C) * sizeof(float) : 0) + kParkQueue * 8 + 16);
}

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (ok == 0u);
}
// one linear bulk copy global -> shared, completion counted in bytes on the mbarrier (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// BULK = true : rows staged by bulk-async copies as described above;
// BULK = false: no staging - the warp reads each record straight from global memory (three coalesced 128-byte
//               loads per 80-class row), four rows in flight per warp.
template <int HEAD, bool BULK>
__global__ void __launch_bounds__(kParkWarps * 32)
k1c_parked_kernel(const __grid_constant__ Plan p, const float* __restrict__ img_shapes,
                  const float* __restrict__ scale_factors, const int* __restrict__ topk_idx,
                  float* __restrict__ score_rows, float* __restrict__ lam_rows, float* __restrict__ boxes,
                  float* __restrict__ row_max, int* __restrict__ row_argmax, unsigned long long* __restrict__ cand,
                  int* __restrict__ cand_cnt, unsigned* __restrict__ cand_maxc, const int* __restrict__ row_slot,
                  const float* __restrict__ cap_scores) {
  extern __shared__ __align__(128) unsigned char park_smem[];
  const unsigned full = 0xffffffffu;
  const int C = p.C, NF = p.num_fg, RF = cap_row_floats(C);     // typed class counts only: RF <= 96
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const size_t per_warp = (BULK ? 32 * (size_t)RF * sizeof(float) : 0) + kParkQueue * 8 + 16;
  unsigned char* wbase = park_smem + warp * per_warp;
  const float* slab = reinterpret_cast<const float*>(wbase);                                   // [32][RF] (BULK)
  unsigned long long* queue = reinterpret_cast<unsigned long long*>(wbase + (BULK ? 32 * (size_t)RF * sizeof(float) : 0));
  const unsigned bar = (unsigned)__cvta_generic_to_shared(queue + kParkQueue);
  const unsigned a_slab = (unsigned)__cvta_generic_to_shared(wbase);

  const int r0 = (blockIdx.x * kParkWarps + warp) * 32;
  const int r = r0 + lane;
  int slot = -1, lv = 0;
  if (r < p.K) {
    lv = level_of_row(p, r);
    if (p.lv[lv].rescan == 0 && p.lv[lv].cap >= 0) slot = row_slot[(size_t)b * p.K + r];
  }
  const unsigned pm = __ballot_sync(full, slot >= 0);
  if (pm == 0u) return;                                                                          // warp-uniform
  const float* src = nullptr;
  if (slot >= 0) src = cap_scores + (((size_t)b * p.n_cap_levels + p.lv[lv].cap) * kCapRows + slot) * RF;
  if constexpr (BULK) {
    // ---- stage: 32 bulk copies in flight, one mbarrier phase
    if (lane == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar, (unsigned)(__popc(pm) * RF * (int)sizeof(float)));
    }
    __syncwarp(full);
    if (slot >= 0) bulk_g2s(a_slab + (unsigned)(lane * RF * (int)sizeof(float)), src, (unsigned)(RF * sizeof(float)), bar);
  }
  // what the box phase needs (while the rows travel)
  float lam = 0.f;
  float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
  if (slot >= 0) {
    const LevelDev& L = p.lv[lv];
    const int n = topk_idx[(size_t)b * p.K + r];
    const int hw = n / L.A, a = n - hw * L.A;
    lam = __ldg(L.lam + (size_t)(b * L.A + a) * L.HW + hw);
    box = k1_decode_box(p, L, b, n, a, hw, img_shapes, scale_factors);
  }
  if constexpr (BULK) mbar_wait(bar, 0);
  // ---- rows: lane = class
  const unsigned lt = (1u << lane) - 1u;
  int qn = 0;
  float my_best = -1.f;
  int my_arg = 0;
  bool my_has = false;
  auto flush = [&]() {
    if (qn == 0) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(cand_cnt + b, qn);
    base = __shfl_sync(full, base, 0);
    __syncwarp(full);
    unsigned long long* dst = cand + (size_t)b * p.K * NF + base;
    for (int i = lane; i < qn; i += 32) dst[i] = queue[i];
    __syncwarp(full);
    qn = 0;
  };
  // one row whose record sits in v[] (chunk j = floats 32 j + lane), 1/sum and the score normaliser at floats C, C + 1
  auto do_row = [&](const int i, const float (&v)[3]) {
    const int jn = C >> 5, ln = C & 31;                   // chunk / lane that hold float C (ln <= 30: C is never 31 mod 32 here)
    const float vn = jn == 0 ? v[0] : (jn == 1 ? v[1] : v[2]);
    const float inv = __shfl_sync(full, vn, ln), den = __shfl_sync(full, vn, ln + 1);
    float* dst = score_rows + ((size_t)b * p.K + r0 + i) * C;
    float best = -1.f;
    int arg = 0;
    bool any = false;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int c = 32 * j + lane;
      if (32 * j >= C) break;
      float sc = 0.f;
      if (c < C) {
        sc = k1_score<HEAD>(v[j], inv, den);
        dst[c] = sc;
        if (sc > best) { best = sc; arg = c; }
      }
      const bool is_cand = c < NF && sc > p.score_thr;
      const unsigned cm = __ballot_sync(full, is_cand);
      if (cm != 0u) {
        if (qn + 32 > kParkQueue) flush();
        if (is_cand)
          queue[qn + __popc(cm & lt)] = ((unsigned long long)__float_as_uint(sc) << 32) |
                                       (unsigned long long)(0xffffffffu - (unsigned)((r0 + i) * NF + c));
        qn += __popc(cm);
        any = true;
      }
    }
    // row maximum and its first class: order-preserving bits, two REDUX
    const unsigned ob = f2ord(best);
    const unsigned om = __reduce_max_sync(full, ob);
    const int am = (int)__reduce_min_sync(full, ob == om ? (unsigned)arg : 0x7fffffffu);
    if (lane == i) { my_best = ord2f(om); my_arg = am; my_has = any; }
  };
  const unsigned long long my_src = reinterpret_cast<unsigned long long>(src);
  for (unsigned m = pm; m != 0u;) {
    int idx[4];
    float v[4][3];
#pragma unroll
    for (int u = 0; u < 4; ++u) {                          // up to four rows in flight
      idx[u] = -1;
      if (m != 0u) { idx[u] = __ffs(m) - 1; m &= m - 1u; }
      const int i = idx[u] < 0 ? 0 : idx[u];
      const float* row = BULK ? slab + i * RF : reinterpret_cast<const float*>(__shfl_sync(full, my_src, i));
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int c = 32 * j + lane;
        v[u][j] = (idx[u] >= 0 && c < RF) ? (BULK ? row[c] : __ldcs(row + c)) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (idx[u] >= 0) do_row(idx[u], v[u]);
  }
  flush();
  // ---- boxes: lane = row
  if (slot >= 0) {
    const size_t o = (size_t)b * p.K + r;
    row_max[o] = my_best;
    row_argmax[o] = my_arg;
    lam_rows[o] = lam;
    reinterpret_cast<float4*>(boxes)[o] = box;
  }
  float wmax = my_has ? fmaxf(fmaxf(box.x, box.y), fmaxf(box.z, box.w)) : -FLT_MAX;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(full, wmax, o));
  if (lane == 0 && wmax > -FLT_MAX) atomicMax(cand_maxc + b, f2ord(wmax));
}

// ------------------------------------------------------------------------------------------
// Same flush for rows that sit anywhere in the block's tile: lane i's scores are at tile_row_i.
template <int C>
__device__ __forceinline__ void k1_flush_rows_at(const float* my_tile_row, float* srow) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned long long dst_mine = reinterpret_cast<unsigned long long>(srow);
  const unsigned long long src_mine = reinterpret_cast<unsigned long long>(my_tile_row);
  __syncwarp();
  for (unsigned m = __ballot_sync(full, dst_mine != 0ull); m != 0u; m &= m - 1u) {
    const int i = __ffs(m) - 1;
    float* dst = reinterpret_cast<float*>(__shfl_sync(full, dst_mine, i));
    const float* src = reinterpret_cast<const float*>(__shfl_sync(full, src_mine, i));
    for (int c = lane; c < C; c += 32) dst[c] = src[c];
  }
  __syncwarp();
}

// K1c (rescan form): levels where many priors are kept (no top-k at all, or 2k >= N) are walked a
// second time with K1a's coalesced tiling; a prior finds its row arithmetically (no top-k:
// row = k_off + n) or through the inverse map K1b scattered (row or -1).  Traffic = the level's
// logits once more, instead of ~25x the kept rows for a sector-granular gather.
// Two phases per 128-position tile (positions j = a*HW + hw, flattened over the level's planes): (1) every thread loads its C logits (full coalesced lines, as in
// K1a) and the kept ones park them in their row of the shared-memory tile and join a list;
// (2) the listed rows are handed out to the first threads of the block, so the row arithmetic
// (softmax, argmax, box decode, candidates) runs once per KEPT row, not once per warp that happens
// to hold one.
// ------------------------------------------------------------------------------------------
template <int C, int HEAD>
__global__ void __launch_bounds__(kRescanThreads, 4)
k1c_rescan_kernel(const __grid_constant__ Plan p, const float* __restrict__ img_shapes,
                  const float* __restrict__ scale_factors, const int* __restrict__ inv_map,
                  int* __restrict__ topk_idx, float* __restrict__ score_rows, float* __restrict__ lam_rows,
                  float* __restrict__ boxes, float* __restrict__ row_max, int* __restrict__ row_argmax,
                  unsigned long long* __restrict__ cand, int* __restrict__ cand_cnt,
                  unsigned* __restrict__ cand_maxc) {
  const int t = blockIdx.x;
  const int b = t / p.rtiles_per_image;
  const int ti = t - b * p.rtiles_per_image;
  int s = -1;
#pragma unroll
  for (int i = 0; i < kMaxLevels; ++i)
    if (i < p.S && p.lv[i].rescan && ti >= p.lv[i].rtile0) s = i;
  const LevelDev& L = p.lv[s];
  const int lt = ti - L.rtile0;
  // tile = kRescanThreads consecutive positions j = a*HW + hw of the level (the order of the key array): planes
  // smaller than a tile (SSD's last levels have 1 to 64 positions) share tiles instead of owning a mostly empty one
  const int j = lt * kRescanThreads + threadIdx.x;
  const bool live = j < L.n;
  const int a = live ? j / L.HW : 0;
  const int hw = live ? j - a * L.HW : 0;
  int ncand = 0, r = -1;
  float bmax = 0.f;
  float* srow = nullptr;
  __shared__ float tile[(C > 0) ? (kRescanThreads * K1Tile<C>::stride) : 1];
  if constexpr (C == 0) {
    // generic class count: rows are recomputed by streaming from global memory, one thread per prior
    float* tile_row = nullptr;
    if (live) {
      const int n = hw * L.A + a;
      float x[1];
      if (L.topk) {
        r = inv_map[(size_t)b * p.N + L.n_off + a * L.HW + hw];
      } else {
        r = L.k_off + n;
        topk_idx[(size_t)b * p.K + r] = n;
      }
      if (r >= 0) {
        tile_row = score_rows + ((size_t)b * p.K + r) * p.C;
        k1_row_body<C, HEAD>(p, L, b, r, n, img_shapes, scale_factors, score_rows, lam_rows, boxes, row_max,
                             row_argmax, x, tile_row, ncand, bmax, srow);
      }
    }
    k1_append_candidates(p, b, r, ncand, bmax, tile_row, cand, cand_cnt, cand_maxc);
  } else {
    if (!L.topk) {
      // every prior of the level is kept (block-uniform): each thread turns its own logits into its row, straight from
      // registers - no parking in shared memory, no hand-over
      float* tile_row = tile + threadIdx.x * K1Tile<C>::stride;
      if (live) {
        float x[C];
        k1_load_logits<C>(L, b, a, hw, x);
        const int n = hw * L.A + a;
        r = L.k_off + n;
        topk_idx[(size_t)b * p.K + r] = n;
        k1_row_body<C, HEAD>(p, L, b, r, n, img_shapes, scale_factors, score_rows, lam_rows, boxes, row_max,
                             row_argmax, x, tile_row, ncand, bmax, srow);
      }
      k1_flush_rows<C>(tile + (threadIdx.x & ~31) * K1Tile<C>::stride, srow);
      k1_append_candidates(p, b, r, ncand, bmax, tile_row, cand, cand_cnt, cand_maxc);
      return;
    }
    __shared__ int s_cnt;
    __shared__ short s_lane[kRescanThreads];      // position inside the tile of each kept prior
    __shared__ int s_row[kRescanThreads];         // its output row
    __shared__ int s_n[kRescanThreads];           // its prior index n = hw*A + a
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    if (live) {
      float x[C];
      k1_load_logits<C>(L, b, a, hw, x);      // every thread loads: coalesced lines wherever a warp stays inside one plane
      if (L.topk) {
        r = inv_map[(size_t)b * p.N + L.n_off + a * L.HW + hw];
      } else {
        r = L.k_off + hw * L.A + a;
        topk_idx[(size_t)b * p.K + r] = hw * L.A + a;
      }
      if (r >= 0) {
        float* mine = tile + threadIdx.x * K1Tile<C>::stride;
#pragma unroll
        for (int c = 0; c < C; ++c) mine[c] = x[c];
        const int pos = atomicAdd(&s_cnt, 1);
        s_lane[pos] = (short)threadIdx.x;
        s_row[pos] = r;
        s_n[pos] = hw * L.A + a;
      }
    }
    __syncthreads();
    const int cnt = s_cnt;
    r = -1;
    float* tile_row = nullptr;
    if ((int)threadIdx.x < cnt) {
      const int src = s_lane[threadIdx.x];
      r = s_row[threadIdx.x];
      tile_row = tile + src * K1Tile<C>::stride;
      float x[C];
#pragma unroll
      for (int c = 0; c < C; ++c) x[c] = tile_row[c];
      const int n = s_n[threadIdx.x];
      k1_row_body<C, HEAD>(p, L, b, r, n, img_shapes, scale_factors, score_rows, lam_rows, boxes, row_max,
                           row_argmax, x, tile_row, ncand, bmax, srow);
    }
    if (cnt > (int)(threadIdx.x & ~31)) {      // warp-uniform: this warp holds rows
      k1_flush_rows_at<C>(tile_row, srow);
      k1_append_candidates(p, b, r, ncand, bmax, tile_row, cand, cand_cnt, cand_maxc);
    }
  }
}

}  // namespace mehhua
