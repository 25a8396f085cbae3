// K1: logits -> softmax / score -> ranking key (K1a, the HBM-streaming kernel) -> per-level top-k
// (K1b, block radix select) -> gather of the kept rows, lambda, decoded boxes, NMS candidates (K1c).
//
// Reference semantics: _get_bboxes per-level block (mmdet/models/dense_heads/Lambda_L2.py:264-326,
// My_L_ssd_head.py:325-361), delta2bbox (core/bbox/coder/delta_xywh_bbox_coder.py:205-267), the
// level-FG test of ComputeObjUnc (Lambda_L2.py:496-502) and the score filter of multiclass_nms
// (core/post_processing/bbox_nms.py:41-66).
#pragma once
#include "common.cuh"

namespace mehhua {

constexpr int kK1aThreads = 128;   // one prior position per thread, C logits in registers
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
template <int C, int HEAD>
__device__ __forceinline__ void softmax_regs(float (&x)[C], float& inv, float& den, float& pfg) {
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

// ------------------------------------------------------------------------------------------
// K1a: stream every logit once.  Tile = 128 consecutive (h,w) positions of one (image, anchor)
// plane; thread t owns position hw and reads its C class logits with stride H*W, so each warp
// load instruction is one coalesced 128-byte line of one class plane.
// C == 0 selects the generic (runtime class count) path.
// ------------------------------------------------------------------------------------------
template <int C, int HEAD>
__global__ void __launch_bounds__(kK1aThreads)
k1a_keys_kernel(const __grid_constant__ Plan p, float* __restrict__ keys, int* __restrict__ level_fg) {
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
  if (hw >= L.HW) return;
  const int CC = (C > 0) ? C : p.C;
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
  const float key = (HEAD == MEHHUA_HEAD_RETINA) ? __fmul_rn(pfg, den) : pfg;
  keys[(size_t)b * p.N + L.n_off + a * L.HW + hw] = key;
  if (pfg > p.fg_thr) level_fg[b * p.S + s] = 1;
}

// ------------------------------------------------------------------------------------------
// K1b: per (image, level) top-k of the keys, sorted descending (exact ties: lower anchor, then lower position).
// grid = (S, B); levels without an active top-k return immediately.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSelThreads)
k1b_select_kernel(const __grid_constant__ Plan p, const float* __restrict__ keys,
                  int* __restrict__ topk_idx, unsigned* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char k1b_smem[];
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(k1b_smem);   // kSelCap
  int* hist = reinterpret_cast<int*>(buf + kSelCap);                            // 4096
  int* sh = hist + 4096;                                                        // 40
  const int s = blockIdx.x, b = blockIdx.y;
  const LevelDev& L = p.lv[s];
  if (!L.topk) return;
  const float* kp = keys + (size_t)b * p.N + L.n_off;
  // composite = key bits << 32 | ~position, position j = a*HW + hw (the key array is anchor-major):
  // exact key ties go to the lower position.  The prior index n = hw*A + a is only rebuilt for
  // the k winners.
  auto get = [&](int j) -> unsigned long long {
    unsigned kb = __float_as_uint(__ldcg(kp + j));
    if (kb > 0x3fffffffu) kb = (kb & 0x80000000u) ? 0u : 0x3fffffffu;   // negatives / >= 2.0 / NaN
    return ((unsigned long long)kb << 32) | (unsigned long long)(0xffffffffu - (unsigned)j);
  };
  const int cnt = block_collect_topk<kSelThreads, kSelCap, 0>(get, L.n, L.k, ~0ull, buf, hist, sh, status);
  int* out = topk_idx + (size_t)b * p.K + L.k_off;
  const int k = min(L.k, cnt);
  for (int i = threadIdx.x; i < k; i += kSelThreads) {
    const int j = (int)(0xffffffffu - (unsigned)(buf[i] & 0xffffffffull));
    const int a = j / L.HW;
    out[i] = (j - a * L.HW) * L.A + a;
  }
}

// ------------------------------------------------------------------------------------------
// K1c: one thread per kept row: recompute its softmax row (bit-identical to K1a), write scores,
// lambda, decoded box, row max / argmax and append its NMS candidates (score > score_thr).
// grid = (ceil(K / 128), B).
// ------------------------------------------------------------------------------------------
template <int C, int HEAD>
__global__ void __launch_bounds__(kGatherThreads)
k1c_gather_kernel(const __grid_constant__ Plan p, const float* __restrict__ img_shapes,
                  const float* __restrict__ scale_factors, int* __restrict__ topk_idx,
                  float* __restrict__ score_rows, float* __restrict__ lam_rows,
                  float* __restrict__ boxes, float* __restrict__ row_max, int* __restrict__ row_argmax,
                  unsigned long long* __restrict__ cand, int* __restrict__ cand_cnt,
                  unsigned* __restrict__ cand_maxc) {
  const int b = blockIdx.y;
  const int r = blockIdx.x * kGatherThreads + threadIdx.x;
  const bool live = r < p.K;
  const int CC = (C > 0) ? C : p.C;
  const int NF = p.num_fg;
  int ncand = 0;
  float bmax = 0.f;
  float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
  float* srow = nullptr;
  if (live) {
    const int s = level_of_row(p, r);
    const LevelDev& L = p.lv[s];
    int n;
    if (L.topk) {
      n = topk_idx[(size_t)b * p.K + r];
    } else {
      n = r - L.k_off;
      topk_idx[(size_t)b * p.K + r] = n;
    }
    const int hw = n / L.A, a = n - hw * L.A;
    const float* __restrict__ src = L.logits + ((size_t)(b * L.A + a) * CC) * L.HW + hw;
    srow = score_rows + ((size_t)b * p.K + r) * CC;
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
        const float sc = (HEAD == MEHHUA_HEAD_RETINA) ? __fmul_rn(pc, den) : pc;
        x[c] = sc;
        if (sc > best) { best = sc; arg = c; }
        if (c < NF && sc > p.score_thr) ++ncand;
      }
#pragma unroll
      for (int c = 0; c < C; ++c) srow[c] = x[c];
    } else {
      float m, inv, den, pfg;
      softmax_stream<HEAD>(src, (size_t)L.HW, CC, m, inv, den, pfg);
      const float nml2 = -__fmul_rn(m, kLog2e);
      for (int c = 0; c < CC; ++c) {
        const float e = ex2_approx(fmaf(__ldg(src + (size_t)c * L.HW), kLog2e, nml2));
        const float pc = __fmul_rn(e, inv);
        const float sc = (HEAD == MEHHUA_HEAD_RETINA) ? __fmul_rn(pc, den) : pc;
        srow[c] = sc;
        if (sc > best) { best = sc; arg = c; }
        if (c < NF && sc > p.score_thr) ++ncand;
      }
    }
    row_max[(size_t)b * p.K + r] = best;
    row_argmax[(size_t)b * p.K + r] = arg;
    lam_rows[(size_t)b * p.K + r] = __ldg(L.lam + (size_t)(b * L.A + a) * L.HW + hw);

    // delta2bbox: denormalise, clamp dw/dh, exp, centre/size -> corners, clip, rescale
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
    reinterpret_cast<float4*>(boxes)[(size_t)b * p.K + r] = box;
    bmax = fmaxf(fmaxf(box.x, box.y), fmaxf(box.z, box.w));
  }
  // warp-aggregated append of this row's candidates (one atomic per warp)
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

}  // namespace mehhua
