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

}  // namespace mehhua
