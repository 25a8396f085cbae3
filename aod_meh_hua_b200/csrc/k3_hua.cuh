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
#ifndef MEHHUA_NMS_CHUNK
#define MEHHUA_NMS_CHUNK 512
#endif
constexpr int kNmsChunk = MEHHUA_NMS_CHUNK;   // candidates requested per chunk
constexpr int kNmsSub = 256;      // candidates resolved per suppression-matrix block
// buf | sbox | kept | hist (aliased by the 256x8-word suppression matrix) | sh | class ids | alive
constexpr size_t kNmsSmem = kNmsCap * 8 + kNmsCap * 16 + MEHHUA_MAX_DETS * 16 + 4096 * 4 + 48 * 4 + kNmsCap * 2 + kNmsCap + MEHHUA_MAX_DETS * 4;

// IoU of the greedy NMS: inter / (area_i + area_j - inter), areas without +1, all fp32-rounded.
__device__ __forceinline__ float iou_nms(const float4 a, const float4 b) {
  const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
  const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

// Does candidate box `bj` of class `cj` survive the kept detections [q_lo, q_hi)?  Boxes of different classes are a class
// offset apart, so only detections of the same class can overlap it: their classes are scanned four at a time (one
// 64-bit shared-memory load, a zero-halfword test) and the IoU is evaluated on a match only.
__device__ __forceinline__ bool nms_clear_of_kept(const float4* kept, const unsigned short* kept_cls, const int q_lo,
                                                  const int q_hi, const unsigned short cj, const float4 bj, const float thr) {
  const unsigned long long cj4 = 0x0001000100010001ull * (unsigned long long)cj;
  for (int q = q_lo & ~3; q < q_hi; q += 4) {
    const unsigned long long x = *reinterpret_cast<const unsigned long long*>(kept_cls + q) ^ cj4;
    if (((x - 0x0001000100010001ull) & ~x & 0x8000800080008000ull) == 0ull) continue;     // no halfword is zero: no class match
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int qq = q + u;
      if (qq >= q_lo && qq < q_hi && kept_cls[qq] == cj && iou_nms(kept[qq], bj) > thr) return false;
    }
  }
  return true;
}


// ------------------------------------------------------------------------------------------
// K3a.  Candidates are visited in descending (score, then ascending flat index) order, in sorted
// chunks pulled out of the unordered candidate list by block radix-select.  Inside a chunk the
// greedy pass works on blocks of 256 candidates: all threads fill a 256x256-bit suppression matrix
// (boxes carry the per-class offset, so only same-class pairs can overlap), one warp then resolves
// keep / suppress sequentially with word-wide mask updates, and the block's new detections
// suppress the rest of the chunk in parallel.  The pass stops at max_per_img detections.
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
  unsigned* rowmask = reinterpret_cast<unsigned*>(hist);                          // [kNmsSub][8], aliases hist
  int* sh = hist + 4096;                                                          // 48
  unsigned short* scls = reinterpret_cast<unsigned short*>(sh + 48);              // kNmsCap
  unsigned char* alive = reinterpret_cast<unsigned char*>(scls + kNmsCap);        // kNmsCap
  unsigned short* kept_idx = reinterpret_cast<unsigned short*>(alive + kNmsCap);  // MAX_DETS
  unsigned short* kept_cls = kept_idx + MEHHUA_MAX_DETS;                          // MAX_DETS: class of each kept detection

  const int b = blockIdx.x;
  const int NF = p.num_fg;
  const int nc = cand_cnt[b];
  const unsigned long long* cp = cand + (size_t)b * p.K * NF;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (size_t)b * p.K;
  const float off_unit = (nc > 0) ? __fadd_rn(ord2f(cand_maxc[b]), 1.0f) : 0.f;
  auto get = [&](int i) -> unsigned long long { return __ldcg(cp + i); };
  const int lane = threadIdx.x & 31;

  int nk = 0, processed = 0;
  unsigned long long hi = ~0ull;
  if (threadIdx.x == 0) sh[45] = 0;      // objects (detections above obj_thr); sh[0..36] belong to the select
  __syncthreads();
  while (nk < p.max_per_img && processed < nc) {
    const int cnt = block_collect_topk<kNmsThreads, kNmsCap, 0>(get, nc, kNmsChunk, hi, buf, hist, sh, status);
    if (cnt == 0) break;
    // offset boxes, class ids, suppression by detections kept from earlier chunks
    for (int i = threadIdx.x; i < cnt; i += kNmsThreads) {
      const unsigned flat = 0xffffffffu - (unsigned)(buf[i] & 0xffffffffull);
      const int r = flat / NF, c = flat - r * NF;
      const float4 o = bx[r];
      const float off = __fmul_rn((float)c, off_unit);
      const float4 sb = make_float4(__fadd_rn(o.x, off), __fadd_rn(o.y, off), __fadd_rn(o.z, off), __fadd_rn(o.w, off));
      const unsigned char ok = nms_clear_of_kept(kept, kept_cls, 0, nk, (unsigned short)c, sb, p.nms_iou) ? 1 : 0;
      sbox[i] = sb;
      scls[i] = (unsigned short)c;
      alive[i] = ok;
    }
    __syncthreads();
    for (int sb0 = 0; sb0 < cnt && nk < p.max_per_img; sb0 += kNmsSub) {
      const int nsb = min(kNmsSub, cnt - sb0);
      // (1) suppression matrix of the block: thread = (row, half of the columns)
      {
        const int ri = threadIdx.x & (kNmsSub - 1), half = threadIdx.x >> 8;
        unsigned wd[4] = {0u, 0u, 0u, 0u};
        if (ri < nsb && alive[sb0 + ri]) {
          const float4 bi = sbox[sb0 + ri];
          const unsigned short ci = scls[sb0 + ri];
#pragma unroll
          for (int wq = 0; wq < 4; ++wq) {
            const int j0 = (half * 4 + wq) * 32;
            unsigned bits = 0u;
            for (int j = max(j0, ri + 1); j < min(j0 + 32, nsb); ++j)
              if (scls[sb0 + j] == ci && iou_nms(bi, sbox[sb0 + j]) > p.nms_iou) bits |= 1u << (j - j0);
            wd[wq] = bits;
          }
        }
#pragma unroll
        for (int wq = 0; wq < 4; ++wq) rowmask[ri * 8 + half * 4 + wq] = wd[wq];
      }
      __syncthreads();
      // (2) sequential resolve by warp 0: lane l < 8 owns alive word l of the block
      const int nk_before = nk;
      if (threadIdx.x < 32) {
        unsigned aw = 0u;
        if (lane < 8)
          for (int j = 0; j < 32; ++j) {
            const int i = lane * 32 + j;
            if (i < nsb && alive[sb0 + i]) aw |= 1u << j;
          }
        int k = nk;
        // walk the alive candidates in order, jumping over suppressed ones word by word: one step per KEPT candidate
        for (int w = 0; w < 8 && w * 32 < nsb && k < p.max_per_img; ++w) {
          unsigned cur = __shfl_sync(0xffffffffu, aw, w);
          while (cur != 0u && k < p.max_per_img) {
            const int j = __ffs(cur) - 1;
            const int i = w * 32 + j;
            if (lane < 8) aw &= ~rowmask[i * 8 + lane];                 // rows only hold later candidates (j' > i)
            if (lane == 0) { kept[k] = sbox[sb0 + i]; kept_idx[k] = (unsigned short)(sb0 + i); kept_cls[k] = scls[sb0 + i]; }
            ++k;
            cur = __shfl_sync(0xffffffffu, aw, w) & ~((2u << j) - 1u);  // what is still alive behind i in this word
          }
        }
        if (lane == 0) sh[44] = k;
      }
      __syncthreads();
      nk = sh[44];
      // the new detections are written out by one thread each (their box loads overlap)
      for (int k = nk_before + threadIdx.x; k < nk; k += kNmsThreads) {
        const unsigned long long e = buf[kept_idx[k]];
        const unsigned flat = 0xffffffffu - (unsigned)(e & 0xffffffffull);
        const int r = flat / NF;
        const float4 o = bx[r];
        float* d = dets + ((size_t)b * p.max_per_img + k) * 5;
        d[0] = o.x; d[1] = o.y; d[2] = o.z; d[3] = o.w; d[4] = __uint_as_float((unsigned)(e >> 32));
        det_labels[(size_t)b * p.max_per_img + k] = (int)(flat - r * NF);
        det_flat[(size_t)b * p.max_per_img + k] = (int)flat;
        if (d[4] > p.obj_thr) atomicAdd(&sh[45], 1);
      }
      // (3) the block's new detections suppress the rest of the chunk
      if (nk < p.max_per_img && nk > nk_before) {
        for (int j = sb0 + nsb + threadIdx.x; j < cnt; j += kNmsThreads) {
          if (!alive[j]) continue;
          if (!nms_clear_of_kept(kept, kept_cls, nk_before, nk, scls[j], sbox[j], p.nms_iou)) alive[j] = 0;
        }
      }
      __syncthreads();
    }
    processed += cnt;
    hi = buf[cnt - 1];
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    n_det[b] = nk;
    n_obj[b] = sh[45];
  }
}

// ------------------------------------------------------------------------------------------
// K3b.  One block per image, rows processed in chunks of 1024 (row order = level order):
//   (a) ordered compaction of the chunk's foreground rows (row max > fg_thr, level flagged FG);
//   (b) one thread per foreground row walks the objects: IoU > cluster_iou -> bit mask;
//   (c) block scan of the per-row pair counts;  (d) pairs emitted in row-major (row, object)
//   order - the order of FG_pos_bbox.nonzero().
// Afterwards warp s averages lambda over level s's pairs (duplicates counted) in a fixed order.
// ------------------------------------------------------------------------------------------
constexpr int kPairThreads = 512;
constexpr int kPairRpt = 2;                       // rows per thread and chunk: 61 KB of shared memory, three blocks per SM
constexpr int kPairChunk = kPairRpt * kPairThreads;
constexpr int kPairWords = MEHHUA_MAX_DETS / 32;
constexpr size_t kPairSmem = MEHHUA_MAX_DETS * 16 + kPairChunk * 16 + kPairChunk * kPairWords * 4 + kPairChunk * 8;

__global__ void __launch_bounds__(kPairThreads)
k3b_pairs_kernel(const __grid_constant__ Plan p, const float* __restrict__ boxes,
                 const float* __restrict__ row_max, const int* __restrict__ row_argmax,
                 const float* __restrict__ lam_rows, const int* __restrict__ level_fg,
                 const float* __restrict__ dets, const int* __restrict__ n_obj,
                 int* __restrict__ pair_row, int* __restrict__ pair_obj, int* __restrict__ pair_cls,
                 int* __restrict__ pair_off, float* __restrict__ lam_mean, unsigned* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char k3b_smem[];
  float4* obox = reinterpret_cast<float4*>(k3b_smem);                       // [MAX_DETS]
  float4* fg_box = obox + MEHHUA_MAX_DETS;                                  // [kPairChunk]
  unsigned* mask = reinterpret_cast<unsigned*>(fg_box + kPairChunk);        // [kPairChunk][kPairWords]
  int* fg_idx = reinterpret_cast<int*>(mask + kPairChunk * kPairWords);     // [kPairChunk]
  int* fg_cnt = fg_idx + kPairChunk;                                        // [kPairChunk]
  __shared__ int wsum[40];
  __shared__ int lvl_cnt[kMaxLevels];
  __shared__ int lvl_fg[kMaxLevels];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nobj = n_obj[b];
  const int nwords = (nobj + 31) >> 5;
  for (int o = threadIdx.x; o < nobj; o += kPairThreads) {
    const float* d = dets + ((size_t)b * p.max_per_img + o) * 5;
    obox[o] = make_float4(d[0], d[1], d[2], d[3]);
  }
  if (threadIdx.x < kMaxLevels) {
    lvl_cnt[threadIdx.x] = 0;
    lvl_fg[threadIdx.x] = (threadIdx.x < p.S) ? level_fg[b * p.S + threadIdx.x] : 0;
  }
  __syncthreads();
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (size_t)b * p.K;
  const float* rmax = row_max + (size_t)b * p.K;
  int running = 0;
  if (nobj > 0) {
    for (int base = 0; base < p.K; base += kPairChunk) {
      // (a) ordered compaction of foreground rows
      const int r0 = base + kPairRpt * threadIdx.x;
      int flags = 0;
#pragma unroll
      for (int j = 0; j < kPairRpt; ++j) {
        const int r = r0 + j;
        if (r < p.K && rmax[r] > p.fg_thr && lvl_fg[level_of_row(p, r)]) flags |= 1 << j;
      }
      const int nf = __popc(flags);
      const int incl = block_incl_scan<kPairThreads>(nf, wsum);
      if (threadIdx.x == kPairThreads - 1) wsum[32] = incl;
      {
        int pos = incl - nf;
#pragma unroll
        for (int j = 0; j < kPairRpt; ++j)
          if (flags & (1 << j)) { fg_idx[pos] = r0 + j; fg_box[pos] = bx[r0 + j]; ++pos; }
      }
      __syncthreads();
      const int nfg = wsum[32];
      // (b) IoU masks: thread = foreground row, objects walked 32 at a time (their boxes are shared-memory broadcasts)
      for (int e = threadIdx.x; e < nfg; e += kPairThreads) {
        const float4 rb = fg_box[e];
        const float area = __fmul_rn(__fsub_rn(rb.z, rb.x), __fsub_rn(rb.w, rb.y));
        int cnt = 0;
        for (int gq = 0; gq < nwords; ++gq) {
          unsigned bits = 0u;
          const int on = min(32, nobj - gq * 32);
          for (int j = 0; j < on; ++j)
            if (iou_overlaps(rb, area, obox[gq * 32 + j]) > p.cluster_iou) bits |= 1u << j;
          mask[e * kPairWords + gq] = bits;
          cnt += __popc(bits);
        }
        fg_cnt[e] = cnt;
      }
      __syncthreads();
      // (c) scan of pair counts, kPairRpt consecutive entries per thread
      const int e0 = kPairRpt * threadIdx.x;
      int c4 = 0;
#pragma unroll
      for (int j = 0; j < kPairRpt; ++j)
        if (e0 + j < nfg) c4 += fg_cnt[e0 + j];
      const int incl2 = block_incl_scan<kPairThreads>(c4, wsum);
      if (threadIdx.x == kPairThreads - 1) wsum[33] = incl2;
      // (d) emit
      int pos = running + incl2 - c4;
      for (int j = 0; j < kPairRpt; ++j) {
        const int e = e0 + j;
        if (e >= nfg) break;
        const int cnt = fg_cnt[e];
        if (cnt == 0) continue;
        const int r = fg_idx[e];
        const int cls = row_argmax[(size_t)b * p.K + r];
        atomicAdd(&lvl_cnt[level_of_row(p, r)], cnt);
        for (int gq = 0; gq < nwords; ++gq) {
          unsigned bits = mask[e * kPairWords + gq];
          while (bits) {
            const int jj = __ffs(bits) - 1;
            bits &= bits - 1;
            if (pos < p.pair_cap) {
              const size_t q = (size_t)b * p.pair_cap + pos;
              pair_row[q] = r; pair_obj[q] = gq * 32 + jj; pair_cls[q] = cls;
            }
            ++pos;
          }
        }
      }
      __syncthreads();
      running += wsum[33];
      __syncthreads();
    }
  }
  // per-level offsets (levels are contiguous in row order, so pairs are too)
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int s = 0; s < p.S; ++s) {
      pair_off[b * (p.S + 1) + s] = min(acc, p.pair_cap);
      acc += lvl_cnt[s];
    }
    pair_off[b * (p.S + 1) + p.S] = min(acc, p.pair_cap);
    if (acc > p.pair_cap) atomicOr(status, MEHHUA_ST_PAIR_OVERFLOW);
  }
  __syncthreads();
  // mean lambda over each level's pairs: warp s, lane-strided partial sums + fixed shuffle tree
  for (int s = w; s < p.S; s += kPairThreads / 32) {
    const int a = pair_off[b * (p.S + 1) + s], e = pair_off[b * (p.S + 1) + s + 1];
    float v = 0.f;
    for (int q = a + lane; q < e; q += 32)
      v += lam_rows[(size_t)b * p.K + pair_row[(size_t)b * p.pair_cap + q]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) lam_mean[b * p.S + s] = (e > a) ? __fdiv_rn(v, (float)(e - a)) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// K3c.  One block per image.  The image's ordered pair list is grouped by (object, level, class):
// pairs are sorted by that key (bitonic sort in shared memory, pair ordinal as tie-break so every
// group is summed in pair order - deterministic), run leaders produce the group means, then one
// thread per OBJECT walks that object's stretch of the sorted group table (class -> level
// aggregation with Sum / Avg / Max selected per level of the hierarchy) and thread 0 folds the
// object values in ascending object order.  The shared-memory sort holds CAP pairs; an image
// with more pairs is processed in several rounds over ascending object ranges (each range's pairs
// fit; one object never has more pairs than there are rows), the fold carrying the object-level
// accumulator across rounds.
// Two instantiations are launched: CAP = kHuaCapSmall serves the images whose pairs fit it (16 KB of
// shared memory: every image of a batch is resident at once), CAP = kHuaCap the others; a block
// whose image belongs to the other instantiation returns at once (the pair count is only known on
// the device).  Same arithmetic either way.
// group key = object << 11 | level << 8 | class  (object <= 256, level < 8, class < 256)
// ------------------------------------------------------------------------------------------
constexpr int kHuaThreads = 256;
constexpr int kHuaCap = 8192;
constexpr int kHuaCapSmall = 1024;

__device__ __forceinline__ float agg_combine(int op, float acc, float v) {
  return op == MEHHUA_AGG_MAX ? fmaxf(acc, v) : acc + v;
}
__device__ __forceinline__ float agg_finish(int op, float acc, int n) {
  return op == MEHHUA_AGG_AVG ? __fdiv_rn(acc, (float)n) : acc;
}

__host__ __device__ inline size_t k3c_smem_bytes(int cap, int C) {
  return (size_t)cap * 16 + (MEHHUA_MAX_DETS + 1) * 4 * 3 + ((C + 31) / 32) * 4 + 64 * 4;
}

template <int CAP>
__global__ void __launch_bounds__(kHuaThreads)
k3c_hua_kernel(const __grid_constant__ Plan p, const int* __restrict__ pair_row,
               const int* __restrict__ pair_obj, const int* __restrict__ pair_cls,
               const int* __restrict__ pair_off, const float* __restrict__ pair_unc,
               const int* __restrict__ n_obj, float* __restrict__ image_scores, unsigned* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char k3c_smem[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(k3c_smem);           // [CAP]
  unsigned* gkey = reinterpret_cast<unsigned*>(keys + CAP);                              // [CAP]
  float* gval = reinterpret_cast<float*>(gkey + CAP);                                    // [CAP]
  int* ocnt = reinterpret_cast<int*>(gval + CAP);                                        // [MAX_DETS + 1] pairs per object
  float* oval = reinterpret_cast<float*>(ocnt + MEHHUA_MAX_DETS + 1);                    // [MAX_DETS + 1] object values of a round
  int* ohas = reinterpret_cast<int*>(oval + MEHHUA_MAX_DETS + 1);                        // [MAX_DETS + 1] object has groups
  unsigned* cls_seen = reinterpret_cast<unsigned*>(ohas + MEHHUA_MAX_DETS + 1);          // [(C+31)/32]
  int* sh = reinterpret_cast<int*>(cls_seen + (p.C + 31) / 32);                          // 64 ints

  const int b = blockIdx.x;
  const int np = pair_off[b * (p.S + 1) + p.S];
  if ((np <= kHuaCapSmall) != (CAP == kHuaCapSmall)) return;      // the other instantiation's image
  const int nobj = min(n_obj[b], MEHHUA_MAX_DETS);
  const int* pobj = pair_obj + (size_t)b * p.pair_cap;
  const int* pcls = pair_cls + (size_t)b * p.pair_cap;
  const float* punc = pair_unc + (size_t)b * p.pair_cap * 3;
  const int* poff = pair_off + b * (p.S + 1);
  for (int i = threadIdx.x; i < (p.C + 31) / 32; i += kHuaThreads) cls_seen[i] = 0u;
  for (int i = threadIdx.x; i <= MEHHUA_MAX_DETS; i += kHuaThreads) ocnt[i] = 0;
  __syncthreads();
  if (np > CAP) {   // pairs per object, needed to cut the object axis into ranges that fit
    for (int q = threadIdx.x; q < np; q += kHuaThreads) atomicAdd(&ocnt[min(pobj[q], MEHHUA_MAX_DETS)], 1);
    __syncthreads();
  }
  // object-level accumulator of the fold (thread 0 only)
  float oacc = 0.f;
  int on = 0;
  int o_lo = 0;
  while (o_lo < max(nobj, 1)) {
    // range [o_lo, o_hi): everything when the image fits, else as many whole objects as fit
    int o_hi = max(nobj, 1);
    if (np > CAP) {
      if (threadIdx.x == 0) {
        int acc = 0, o = o_lo;
        while (o < nobj && acc + ocnt[o] <= CAP) { acc += ocnt[o]; ++o; }
        if (o == o_lo) { atomicOr(status, MEHHUA_ST_PAIR_OVERFLOW); ++o; }   // one object alone overflows the sort
        sh[41] = o;
      }
      __syncthreads();
      o_hi = sh[41];
    }
    if (threadIdx.x == 0) sh[42] = 0;
    for (int o = o_lo + threadIdx.x; o < o_hi && o <= MEHHUA_MAX_DETS; o += kHuaThreads) ohas[o] = 0;
    __syncthreads();
    // collect the range's pairs as inverted composites (descending sort -> ascending (key, ordinal))
    for (int q = threadIdx.x; q < np; q += kHuaThreads) {
      const int o = pobj[q];
      if (o >= o_lo && o < o_hi) {
        const int pos = atomicAdd(&sh[42], 1);
        if (pos < CAP) {
          const unsigned gk = ((unsigned)o << 11) | ((unsigned)level_of_pair(poff, p.S, q) << 8) | (unsigned)pcls[q];
          keys[pos] = ~(((unsigned long long)gk << 32) | (unsigned)q);
        }
      }
    }
    __syncthreads();
    const int m = min(sh[42], CAP);
    int n2 = 1;
    while (n2 < m) n2 <<= 1;
    for (int i = m + threadIdx.x; i < n2; i += kHuaThreads) keys[i] = 0ull;
    __syncthreads();
    block_bitonic_desc<kHuaThreads>(keys, n2);
    // run leaders -> group means, written in key order
    int running = 0;
    for (int base = 0; base < m; base += kHuaThreads) {
      const int i = base + threadIdx.x;
      unsigned gk = 0u;
      bool leader = false;
      if (i < m) {
        gk = (unsigned)((~keys[i]) >> 32);
        leader = (i == 0) || gk != (unsigned)((~keys[i - 1]) >> 32);
      }
      const int incl = block_incl_scan<kHuaThreads>(leader ? 1 : 0, sh);
      if (threadIdx.x == kHuaThreads - 1) sh[40] = incl;
      if (leader) {
        float sum = 0.f;
        int cnt = 0;
        for (int j = i; j < m; ++j) {
          const unsigned long long e = ~keys[j];
          if ((unsigned)(e >> 32) != gk) break;
          sum += punc[(unsigned)(e & 0xffffffffull) * 3 + 2];     // epistemic, in pair order
          ++cnt;
        }
        const int gi = running + incl - 1;
        gkey[gi] = gk;
        gval[gi] = __fdiv_rn(sum, (float)cnt);
        const unsigned cls = gk & 255u;
        atomicOr(&cls_seen[cls >> 5], 1u << (cls & 31));
      }
      __syncthreads();
      running += sh[40];
      __syncthreads();
    }
    // thread o walks the stretch of the (object, level, class)-sorted group table that belongs to object o
    const int ng = running;
    for (int o = o_lo + threadIdx.x; o < o_hi; o += kHuaThreads) {
      int lo = -1, hi = ng;            // first group g with object(g) >= o
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((gkey[mid] >> 11) >= (unsigned)o) hi = mid; else lo = mid; }
      int g = hi;
      if (g < ng && (gkey[g] >> 11) == (unsigned)o) {
        float lacc = 0.f;
        int ln = 0;
        while (g < ng && (gkey[g] >> 11) == (unsigned)o) {
          const unsigned ol = gkey[g] >> 8;
          float cacc = (p.agg_class == MEHHUA_AGG_MAX) ? -FLT_MAX : 0.f;
          int cn = 0;
          while (g < ng && (gkey[g] >> 8) == ol) { cacc = agg_combine(p.agg_class, cacc, gval[g]); ++cn; ++g; }
          const float cv = agg_finish(p.agg_class, cacc, cn);
          lacc = (ln == 0) ? cv : agg_combine(p.agg_scale, lacc, cv);
          ++ln;
        }
        oval[o] = agg_finish(p.agg_scale, lacc, ln);
        ohas[o] = 1;
      }
    }
    __syncthreads();
    // fold the objects of the range in ascending order
    if (threadIdx.x == 0) {
      for (int o = o_lo; o < o_hi; ++o)
        if (ohas[o]) { oacc = (on == 0) ? oval[o] : agg_combine(p.agg_object, oacc, oval[o]); ++on; }
    }
    __syncthreads();
    o_lo = o_hi;
  }
  if (threadIdx.x == 0) {
    float out = on > 0 ? agg_finish(p.agg_object, oacc, on) : 0.f;
    if (p.cls_w) {
      int k = 0;
      for (int i = 0; i < (p.C + 31) / 32; ++i) k += __popc(cls_seen[i]);
      out = __fmul_rn(out, (float)k);
    }
    image_scores[b] = out;
  }
}

}  // namespace mehhua
