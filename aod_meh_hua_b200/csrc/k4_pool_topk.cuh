// K4: top-k over the unlabelled pool's image scores (block radix select).
//
// Reference semantics: update_X_L (mmdet/utils/active_datasets.py:105-107, 124):
//   arg = uncertainty[all_X_U].argsort();  picked = all_X_U[arg[-n_top:]]
// i.e. the n_top largest scores among unlabelled images.  numpy's default argsort is not stable,
// so ties are unpinned in the reference; here ties go to the larger index (what a stable
// ascending argsort followed by [-k:] keeps), which makes the selected set a pure function of
// the scores.
#pragma once
#include "common.cuh"

namespace mehhua {

constexpr int kPoolThreads = 1024;
constexpr int kPoolCap = 4096;
constexpr size_t kPoolSmem = kPoolCap * 8 + 4096 * 4 + 40 * 4;

__global__ void __launch_bounds__(kPoolThreads)
k4_pool_topk_kernel(const float* __restrict__ scores, const unsigned char* __restrict__ mask,
                    long long n, int k, long long* __restrict__ idx_out, int* __restrict__ n_sel,
                    unsigned* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char k4_smem[];
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(k4_smem);
  int* hist = reinterpret_cast<int*>(buf + kPoolCap);
  int* sh = hist + 4096;
  // composite: order-preserving score bits (any sign) << 32 | index; 0 = not a candidate
  auto get = [&](int i) -> unsigned long long {
    if (mask != nullptr && mask[i] == 0) return 0ull;
    const float v = __ldg(scores + i);
    if (v != v) return 0ull;   // NaN scores are never selected
    return ((unsigned long long)f2ord(v) << 32) | (unsigned long long)(unsigned)i;
  };
  int written = 0;
  unsigned long long hi = ~0ull;
  while (written < k) {
    const int want = min(k - written, kPoolCap);
    const int cnt = block_collect_topk<kPoolThreads, kPoolCap, 1>(get, (int)n, want, hi, buf, hist, sh, status);
    const int take = min(cnt, k - written);
    for (int i = threadIdx.x; i < take; i += kPoolThreads)
      idx_out[written + i] = (long long)(buf[i] & 0xffffffffull);
    written += take;
    if (cnt < want) break;   // pool exhausted
    hi = buf[cnt - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_sel = written;
}

// ------------------------------------------------------------------------------------------
// K4 for large pools (n >= kPoolMultiMin; cfg 5: 10^6 scores, k = 25 000): the same selection spread
// over the whole GPU.  Radix select over the 64-bit composites (order-preserving score bits << 32 |
// index, all distinct) with grid-wide histograms: every pass narrows the digit that holds the k-th
// composite until at most kPoolCap candidates share the decided prefix (ties in the score - e.g. the
// exact zeros of object-less images - are separated by the index digits); then everything at or above
// the prefix is compacted, sorted (4096-element bitonic chunks in shared memory, then merge-path
// merges) and the first k indices are written.  All counts stay on the device: the launch sequence is
// fixed, later passes return at once when an earlier one finished the job.
// ------------------------------------------------------------------------------------------
constexpr long long kPoolMultiMin = 32768;
constexpr int kPoolPasses = 6;
constexpr int kPoolHistThreads = 1024;
constexpr int kPoolChunk = 4096;              // elements per shared-memory sort chunk
constexpr int kPoolMergeVT = 8;               // outputs per thread of a merge pass
__device__ __constant__ const int kPoolShift[kPoolPasses] = {52, 41, 30, 19, 8, 0};
__device__ __constant__ const int kPoolBits[kPoolPasses]  = {12, 11, 11, 11, 11, 8};

struct PoolState {
  unsigned long long prefix, mask;   // decided digits of the k-th composite
  int need;                          // how many of the elements that match the prefix are still wanted
  int above;                         // elements above the prefix (all selected)
  int done;                          // 1: prefix final (compaction can run); passes that follow return
  int all;                           // 1: fewer than k candidates: everything is selected
  int m;                             // compacted elements (filled by the compaction)
  int arrived[kPoolPasses];          // blocks that have flushed their histogram, per pass
  int pad;
};
// workspace: PoolState | hist[kPoolPasses][4096] | two composite buffers of cap elements each
__host__ __device__ inline size_t k4m_buf_elems(long long n, int k) {
  const long long c = (long long)k + kPoolCap;
  return (size_t)(((c < n ? c : n) + kPoolChunk - 1) / kPoolChunk) * kPoolChunk;
}
__host__ __device__ inline size_t k4m_state_bytes() { return 256 + (size_t)kPoolPasses * 4096 * sizeof(int); }
__host__ __device__ inline size_t k4m_workspace_bytes(long long n, int k) {
  return k4m_state_bytes() + 2 * k4m_buf_elems(n, k) * sizeof(unsigned long long);
}

__device__ __forceinline__ unsigned long long k4_composite(const float* __restrict__ scores,
                                                           const unsigned char* __restrict__ mask, long long i) {
  if (mask != nullptr && mask[i] == 0) return 0ull;
  const float v = __ldg(scores + i);
  if (v != v) return 0ull;   // NaN scores are never selected
  return ((unsigned long long)f2ord(v) << 32) | (unsigned long long)(unsigned)i;
}

__global__ void __launch_bounds__(kPoolHistThreads)
k4m_hist_kernel(const float* __restrict__ scores, const unsigned char* __restrict__ mask, long long n, int k,
                int pass, PoolState* __restrict__ stp, int* __restrict__ ghist) {
  __shared__ int hist[4096];
  __shared__ int sh[40];
  PoolState& st = *stp;
  if (pass > 0 && st.done) return;
  const int shift = kPoolShift[pass], bits = kPoolBits[pass], bins = 1 << bits;
  const unsigned long long prefix = pass ? st.prefix : 0ull, pmask = pass ? st.mask : 0ull;
  for (int i = threadIdx.x; i < bins; i += kPoolHistThreads) hist[i] = 0;
  __syncthreads();
  for (long long i0 = (long long)blockIdx.x * kPoolHistThreads * kSelUnroll + threadIdx.x; i0 < n;
       i0 += (long long)gridDim.x * kPoolHistThreads * kSelUnroll) {
    unsigned long long e[kSelUnroll];
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) { const long long i = i0 + u * kPoolHistThreads; e[u] = (i < n) ? k4_composite(scores, mask, i) : 0ull; }
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u)
      if (e[u] != 0ull && (e[u] & pmask) == prefix) atomicAdd(&hist[(int)((e[u] >> shift) & (bins - 1))], 1);
  }
  __syncthreads();
  int* gh = ghist + pass * 4096;
  for (int i = threadIdx.x; i < bins; i += kPoolHistThreads)
    if (hist[i] != 0) atomicAdd(gh + i, hist[i]);
  // the last block to arrive picks the digit of the k-th composite from the complete histogram
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) sh[39] = atomicAdd(&st.arrived[pass], 1);
  __syncthreads();
  if (sh[39] != (int)gridDim.x - 1) return;
  __threadfence();
  const int need = pass ? st.need : k;
  const int per = (bins + kPoolHistThreads - 1) / kPoolHistThreads;
  const int top = bins - 1 - (int)threadIdx.x * per;
  int csum = 0;
  for (int j = 0; j < per; ++j) { const int b = top - j; if (b >= 0) csum += __ldcg(gh + b); }
  const int incl = block_incl_scan<kPoolHistThreads>(csum, sh);
  if (threadIdx.x == kPoolHistThreads - 1) sh[32] = incl;
  if (incl >= need && incl - csum < need) {
    int acc = incl - csum, b = top;
    for (;; --b) { const int h = __ldcg(gh + b); if (acc + h >= need) break; acc += h; }
    sh[33] = b; sh[34] = acc; sh[35] = __ldcg(gh + b);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (sh[32] < need) {          // only possible in pass 0: fewer candidates than k
      st.all = 1; st.done = 1; st.prefix = 0ull; st.mask = 0ull; st.need = sh[32]; st.above = 0;
    } else {
      const int d = sh[33], cnt_above = sh[34], within = sh[35];
      st.prefix = prefix | ((unsigned long long)d << shift);
      st.mask = pmask | ((unsigned long long)(bins - 1) << shift);
      st.above = (pass ? st.above : 0) + cnt_above;
      st.need = need - cnt_above;
      st.all = 0;
      st.done = (within <= kPoolCap || pass == kPoolPasses - 1) ? 1 : 0;
    }
  }
}

// every element at or above the decided prefix -> buf (unordered); at most above + kPoolCap of them
__global__ void __launch_bounds__(kPoolHistThreads)
k4m_compact_kernel(const float* __restrict__ scores, const unsigned char* __restrict__ mask, long long n,
                   PoolState* __restrict__ stp, unsigned long long* __restrict__ buf, long long cap) {
  PoolState& st = *stp;
  const unsigned long long prefix = st.prefix, pmask = st.mask;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  for (long long w0 = (long long)blockIdx.x * kPoolHistThreads + (threadIdx.x & ~31); w0 < n;      // warp-uniform trip count
       w0 += (long long)gridDim.x * kPoolHistThreads) {
    const long long i0 = w0 + lane;
    const unsigned long long e = (i0 < n) ? k4_composite(scores, mask, i0) : 0ull;
    const bool keep = e != 0ull && (e & pmask) >= prefix;
    const unsigned bm = __ballot_sync(full, keep);
    if (bm == 0u) continue;
    int base = 0;
    if (lane == 0) base = atomicAdd(&st.m, __popc(bm));
    base = __shfl_sync(full, base, 0);
    const long long pos = (long long)base + __popc(bm & ((1u << lane) - 1u));
    if (keep && pos < cap) buf[pos] = e;
  }
}

// descending bitonic sort of each kPoolChunk-element chunk of buf[0..m) (zero-padded) in shared memory
__global__ void __launch_bounds__(1024)
k4m_chunk_sort_kernel(const PoolState* __restrict__ stp, unsigned long long* __restrict__ buf) {
  extern __shared__ __align__(16) unsigned char k4_smem[];
  unsigned long long* s = reinterpret_cast<unsigned long long*>(k4_smem);
  const long long m = stp->m;
  const long long c0 = (long long)blockIdx.x * kPoolChunk;
  if (c0 >= m) return;
  for (int i = threadIdx.x; i < kPoolChunk; i += 1024) s[i] = (c0 + i < m) ? buf[c0 + i] : 0ull;
  __syncthreads();
  block_bitonic_desc<1024>(s, kPoolChunk);
  for (int i = threadIdx.x; i < kPoolChunk; i += 1024) buf[c0 + i] = s[i];   // pads (zeros) sort to the end of the chunk
}

// one merge level: descending runs of length run (elements) in src -> runs of 2*run in dst (merge path)
__global__ void __launch_bounds__(256)
k4m_merge_kernel(const PoolState* __restrict__ stp, const unsigned long long* __restrict__ src,
                 unsigned long long* __restrict__ dst, long long run) {
  const long long m = ((long long)stp->m + kPoolChunk - 1) / kPoolChunk * kPoolChunk;   // chunks are zero-padded
  const long long o0 = ((long long)blockIdx.x * 256 + threadIdx.x) * kPoolMergeVT;
  if (o0 >= m) return;
  const long long base = o0 / (2 * run) * (2 * run);
  const long long la = min(run, m - base), lb = max(0ll, min(run, m - base - run));
  const unsigned long long* A = src + base;
  const unsigned long long* Bq = src + base + run;
  const long long diag = o0 - base;
  // how many of the first `diag` merged elements come from A (descending; keys are distinct, zeros only as pads; A first on a tie)
  long long lo = max(0ll, diag - lb), hi = min(diag, la);
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (__ldcg(A + mid) >= __ldcg(Bq + (diag - 1 - mid))) lo = mid + 1; else hi = mid;
  }
  long long i = lo, j = diag - lo;
#pragma unroll
  for (int t = 0; t < kPoolMergeVT; ++t) {
    if (o0 + t >= base + la + lb) break;
    const unsigned long long a = (i < la) ? __ldcg(A + i) : 0ull, b = (j < lb) ? __ldcg(Bq + j) : 0ull;
    const bool ta = (i < la) && (j >= lb || a >= b);
    dst[o0 + t] = ta ? a : b;
    if (ta) ++i; else ++j;
  }
}

__global__ void k4m_output_kernel(const PoolState* __restrict__ stp, const unsigned long long* __restrict__ buf, int k,
                                  long long* __restrict__ idx_out, int* __restrict__ n_sel) {
  const int take = min(k, stp->m);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < take; i += gridDim.x * blockDim.x)
    idx_out[i] = (long long)(buf[i] & 0xffffffffull);
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_sel = take;
}

}  // namespace mehhua
