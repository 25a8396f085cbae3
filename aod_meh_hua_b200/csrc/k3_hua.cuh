// K3: hierarchical uncertainty aggregation.
//   K3a  multi-class NMS -> detections / objects          (one block per image)
//   K3b  IoU clustering of kept boxes onto objects -> ordered (box, object) pair list + lambda mean
//   K3c  segmented means keyed by (object, level, class), class -> level -> object aggregation
//
// Reference semantics: multiclass_nms (mmdet/core/post_processing/bbox_nms.py:7-93) over
// mmcv.ops.nms.batched_nms (mmcv 1.3.8, not vendored: class offset = label*(max coord + 1), greedy
// NMS in descending score, suppress IoU > thr); GetObjectIdx + bbox_overlaps
// (models/dense_heads/Lambda_L2.py:343-349, core/bbox/iou_calculators/iou2d_calculator.py:206-252);
// ComputeObjUnc prologue / grouping (Lambda_L2.py:503-515, 526-536); AggregateObjScaleUnc
// (Lambda_L2.py:597-619).
#pragma once
#include "common.cuh"

namespace mehhua {

constexpr int kNmsThreads = 512;
constexpr int kNmsCap = 2048;     // candidates held in shared memory per chunk
constexpr int kNmsChunk = 1024;   // candidates requested per chunk
constexpr size_t kNmsSmem = kNmsCap * 8 + 4096 * 4 + 40 * 4 + kNmsCap * 16 + kNmsCap + MEHHUA_MAX_DETS * 16;

// IoU of the greedy NMS: inter / (area_i + area_j - inter), areas without +1, all fp32-rounded.
__device__ __forceinline__ float iou_nms(const float4 a, const float4 b) {
  const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
  const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

// ------------------------------------------------------------------------------------------
// K3a.  Candidates are visited in descending (score, then ascending flat index) order, in sorted
// chunks pulled out of the unordered candidate list by block radix-select; the greedy pass stops
// as soon as max_per_img detections are kept, so normally one chunk is enough.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNmsThreads)
k3a_nms_kernel(const __grid_constant__ Plan p, const unsigned long long* __restrict__ cand,
               const int* __restrict__ cand_cnt, const unsigned* __restrict__ cand_maxc,
               const float* __restrict__ boxes, float* __restrict__ dets, int* __restrict__ det_labels,
               int* __restrict__ det_flat, int* __restrict__ n_det, int* __restrict__ n_obj,
               unsigned* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char k3a_smem[];
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(k3a_smem);     // kNmsCap
  float4* sbox = reinterpret_cast<float4*>(buf + kNmsCap);                        // kNmsCap
  float4* kept = sbox + kNmsCap;                                                  // MAX_DETS
  int* hist = reinterpret_cast<int*>(kept + MEHHUA_MAX_DETS);                     // 4096
  int* sh = hist + 4096;                                                          // 40
  unsigned char* alive = reinterpret_cast<unsigned char*>(sh + 40);               // kNmsCap

  const int b = blockIdx.x;
  const int NF = p.num_fg;
  const int nc = cand_cnt[b];
  const unsigned long long* cp = cand + (size_t)b * p.K * NF;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (size_t)b * p.K;
  const float off_unit = (nc > 0) ? __fadd_rn(ord2f(cand_maxc[b]), 1.0f) : 0.f;
  auto get = [&](int i) -> unsigned long long { return __ldcg(cp + i); };

  int nk = 0, nobj = 0, processed = 0;
  unsigned long long hi = ~0ull;
  while (nk < p.max_per_img && processed < nc) {
    const int cnt = block_collect_topk<kNmsThreads, kNmsCap, 0>(get, nc, kNmsChunk, hi, buf, hist, sh, status);
    if (cnt == 0) break;
    for (int i = threadIdx.x; i < cnt; i += kNmsThreads) {
      const unsigned flat = 0xffffffffu - (unsigned)(buf[i] & 0xffffffffull);
      const int r = flat / NF, c = flat - r * NF;
      const float4 o = bx[r];
      const float off = __fmul_rn((float)c, off_unit);
      float4 sb = make_float4(__fadd_rn(o.x, off), __fadd_rn(o.y, off), __fadd_rn(o.z, off), __fadd_rn(o.w, off));
      unsigned char ok = 1;
      for (int q = 0; q < nk; ++q)
        if (iou_nms(kept[q], sb) > p.nms_iou) { ok = 0; break; }
      sbox[i] = sb;
      alive[i] = ok;
    }
    __syncthreads();
    for (int i = 0; i < cnt; ++i) {
      if (!alive[i]) continue;          // uniform: alive[] only changes before a barrier
      const float4 bi = sbox[i];
      if (threadIdx.x == 0) {
        const unsigned long long e = buf[i];
        const unsigned flat = 0xffffffffu - (unsigned)(e & 0xffffffffull);
        const int r = flat / NF, c = flat - r * NF;
        const float sc = __uint_as_float((unsigned)(e >> 32));
        const float4 o = bx[r];
        float* d = dets + ((size_t)b * p.max_per_img + nk) * 5;
        d[0] = o.x; d[1] = o.y; d[2] = o.z; d[3] = o.w; d[4] = sc;
        det_labels[(size_t)b * p.max_per_img + nk] = c;
        det_flat[(size_t)b * p.max_per_img + nk] = (int)flat;
        kept[nk] = bi;
        if (sc > p.obj_thr) ++nobj;
      }
      ++nk;
      if (nk >= p.max_per_img) break;
      for (int j = i + 1 + threadIdx.x; j < cnt; j += kNmsThreads)
        if (alive[j] && iou_nms(bi, sbox[j]) > p.nms_iou) alive[j] = 0;
      __syncthreads();
    }
    processed += cnt;
    hi = buf[cnt - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    n_det[b] = nk;
    n_obj[b] = nobj;
  }
}

// ------------------------------------------------------------------------------------------
// K3b.  One block per image.  For every foreground row (row max > fg_thr, level flagged FG) test
// IoU > cluster_iou against each object; pairs are emitted in row-major (row, object) order - the
// order of FG_pos_bbox.nonzero() - by an ordered block scan, level by level.
// ------------------------------------------------------------------------------------------
constexpr int kPairThreads = 256;

__global__ void __launch_bounds__(kPairThreads)
k3b_pairs_kernel(const __grid_constant__ Plan p, const float* __restrict__ boxes,
                 const float* __restrict__ row_max, const int* __restrict__ row_argmax,
                 const float* __restrict__ lam_rows, const int* __restrict__ level_fg,
                 const float* __restrict__ dets, const int* __restrict__ n_obj,
                 int* __restrict__ pair_row, int* __restrict__ pair_obj, int* __restrict__ pair_cls,
                 int* __restrict__ pair_off, float* __restrict__ lam_mean, unsigned* __restrict__ status) {
  __shared__ float4 obox[MEHHUA_MAX_DETS];
  __shared__ int wsum[32];
  __shared__ float fsum[32];
  __shared__ int s_total;
  const int b = blockIdx.x;
  const int nobj = n_obj[b];
  for (int o = threadIdx.x; o < nobj; o += kPairThreads) {
    const float* d = dets + ((size_t)b * p.max_per_img + o) * 5;
    obox[o] = make_float4(d[0], d[1], d[2], d[3]);
  }
  __syncthreads();
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (size_t)b * p.K;
  int running = 0;
  for (int s = 0; s < p.S; ++s) {
    const LevelDev& L = p.lv[s];
    const int start = running;
    float lam_acc = 0.f;
    if (nobj > 0 && level_fg[b * p.S + s]) {
      for (int base = 0; base < L.k; base += kPairThreads) {
        const int i = base + threadIdx.x;
        const int r = L.k_off + i;
        unsigned words[MEHHUA_MAX_DETS / 32];
#pragma unroll
        for (int w = 0; w < MEHHUA_MAX_DETS / 32; ++w) words[w] = 0u;
        int cnt = 0;
        if (i < L.k && row_max[(size_t)b * p.K + r] > p.fg_thr) {
          const float4 rb = bx[r];
          const float area = __fmul_rn(__fsub_rn(rb.z, rb.x), __fsub_rn(rb.w, rb.y));
#pragma unroll
          for (int w = 0; w < MEHHUA_MAX_DETS / 32; ++w) {
            unsigned bits = 0u;
            for (int j = 0; j < 32; ++j) {
              const int o = w * 32 + j;
              if (o < nobj && iou_overlaps(rb, area, obox[o]) > p.cluster_iou) bits |= 1u << j;
            }
            words[w] = bits;
            cnt += __popc(bits);
          }
        }
        const int incl = block_incl_scan<kPairThreads>(cnt, wsum);
        if (threadIdx.x == kPairThreads - 1) s_total = incl;
        if (cnt > 0) {
          int pos = running + incl - cnt;
          const int cls = row_argmax[(size_t)b * p.K + r];
#pragma unroll
          for (int w = 0; w < MEHHUA_MAX_DETS / 32; ++w) {
            unsigned bits = words[w];
            while (bits) {
              const int j = __ffs(bits) - 1;
              bits &= bits - 1;
              if (pos < p.pair_cap) {
                const size_t q = (size_t)b * p.pair_cap + pos;
                pair_row[q] = r; pair_obj[q] = w * 32 + j; pair_cls[q] = cls;
              }
              ++pos;
            }
          }
          lam_acc = __fmaf_rn((float)cnt, lam_rows[(size_t)b * p.K + r], lam_acc);
        }
        __syncthreads();
        running += s_total;
        __syncthreads();
      }
    }
    // mean lambda over the level's pairs (duplicates counted), fixed reduction tree
    float v = lam_acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) fsum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < kPairThreads / 32; ++w) tot += fsum[w];
      const int np = running - start;
      lam_mean[b * p.S + s] = np > 0 ? __fdiv_rn(tot, (float)np) : 0.f;
      pair_off[b * (p.S + 1) + s] = min(start, p.pair_cap);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    pair_off[b * (p.S + 1) + p.S] = min(running, p.pair_cap);
    if (running > p.pair_cap) atomicOr(status, MEHHUA_ST_PAIR_OVERFLOW);
  }
}

// ------------------------------------------------------------------------------------------
// K3c.  One block per image, one warp per object at a time.  Lane l owns the (level, class) cells
// with class % 32 == l and walks the image's ordered pair list, so every cell is accumulated by a
// single thread in pair order (deterministic, no atomics).  Then class -> level -> object
// aggregation with Sum / Avg / Max selected per level of the hierarchy.
// ------------------------------------------------------------------------------------------
constexpr int kHuaThreads = 256;

__device__ __forceinline__ float agg_combine(int op, float acc, float v) {
  return op == MEHHUA_AGG_MAX ? fmaxf(acc, v) : acc + v;
}
__device__ __forceinline__ float agg_finish(int op, float acc, int n) {
  return op == MEHHUA_AGG_AVG ? __fdiv_rn(acc, (float)n) : acc;
}

__global__ void __launch_bounds__(kHuaThreads)
k3c_hua_kernel(const __grid_constant__ Plan p, const int* __restrict__ pair_row,
               const int* __restrict__ pair_obj, const int* __restrict__ pair_cls,
               const int* __restrict__ pair_off, const float* __restrict__ pair_unc,
               const int* __restrict__ n_obj, float* __restrict__ image_scores) {
  extern __shared__ __align__(16) unsigned char k3c_smem[];
  const int cells = p.S * p.C;
  float* csum = reinterpret_cast<float*>(k3c_smem);                 // [warps][cells]
  int* ccnt = reinterpret_cast<int*>(csum + (kHuaThreads / 32) * cells);
  float* oval = reinterpret_cast<float*>(ccnt + (kHuaThreads / 32) * cells);   // [MAX_DETS]
  int* ohas = reinterpret_cast<int*>(oval + MEHHUA_MAX_DETS);                   // [MAX_DETS]
  unsigned* cls_seen = reinterpret_cast<unsigned*>(ohas + MEHHUA_MAX_DETS);     // [(C+31)/32]

  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nobj = n_obj[b];
  const int np = pair_off[b * (p.S + 1) + p.S];
  const int* prow = pair_row + (size_t)b * p.pair_cap;
  const int* pobj = pair_obj + (size_t)b * p.pair_cap;
  const int* pcls = pair_cls + (size_t)b * p.pair_cap;
  const float* punc = pair_unc + (size_t)b * p.pair_cap * 3;
  for (int i = threadIdx.x; i < (p.C + 31) / 32; i += kHuaThreads) cls_seen[i] = 0u;
  for (int i = threadIdx.x; i < MEHHUA_MAX_DETS; i += kHuaThreads) { oval[i] = 0.f; ohas[i] = 0; }
  __syncthreads();
  float* ms = csum + w * cells;
  int* mc = ccnt + w * cells;
  for (int o = w; o < nobj; o += kHuaThreads / 32) {
    for (int i = lane; i < cells; i += 32) { ms[i] = 0.f; mc[i] = 0; }
    __syncwarp();
    for (int q = 0; q < np; ++q) {
      if (pobj[q] != o) continue;                 // warp-uniform
      const int cls = pcls[q];
      if ((cls & 31) == lane) {
        const int cell = level_of_row(p, prow[q]) * p.C + cls;
        ms[cell] += punc[q * 3 + 2];              // epistemic
        mc[cell] += 1;
      }
    }
    __syncwarp();
    // class aggregation per level, then level aggregation (lane 0 keeps the running value)
    float lvl_acc = 0.f;
    int lvl_n = 0;
    for (int s = 0; s < p.S; ++s) {
      float acc = (p.agg_class == MEHHUA_AGG_MAX) ? -FLT_MAX : 0.f;
      int n = 0;
      for (int c = lane; c < p.C; c += 32) {
        const int k = mc[s * p.C + c];
        if (k > 0) {
          acc = agg_combine(p.agg_class, acc, __fdiv_rn(ms[s * p.C + c], (float)k));
          ++n;
          atomicOr(&cls_seen[c >> 5], 1u << (c & 31));
        }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const float oa = __shfl_xor_sync(0xffffffffu, acc, d);
        const int on = __shfl_xor_sync(0xffffffffu, n, d);
        acc = agg_combine(p.agg_class, acc, oa);
        n += on;
      }
      if (n > 0) {
        const float v = agg_finish(p.agg_class, acc, n);
        lvl_acc = (lvl_n == 0) ? v : agg_combine(p.agg_scale, lvl_acc, v);
        ++lvl_n;
      }
    }
    if (lane == 0 && lvl_n > 0) {
      oval[o] = agg_finish(p.agg_scale, lvl_acc, lvl_n);
      ohas[o] = 1;
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float acc = 0.f;
    int n = 0;
    for (int o = 0; o < nobj; ++o)
      if (ohas[o]) { acc = (n == 0) ? oval[o] : agg_combine(p.agg_object, acc, oval[o]); ++n; }
    float out = n > 0 ? agg_finish(p.agg_object, acc, n) : 0.f;
    if (p.cls_w) {
      int k = 0;
      for (int i = 0; i < (p.C + 31) / 32; ++i) k += __popc(cls_seen[i]);
      out = __fmul_rn(out, (float)k);
    }
    image_scores[b] = out;
  }
}

}  // namespace mehhua
