// K2: Dirichlet-sampled epistemic uncertainty per (box, object) pair.
//
// Reference semantics (mmdet/models/dense_heads/Lambda_L2.py:513-525):
//   lambda' = mean(lambda_pairs) / (lambda + 1e-7) * 25;  alpha = score_row * lambda'
//   x_t ~ Dirichlet(alpha), t = 1..T (T = 500)
//   avg = mean_t x_t;  total = -sum_c avg ln avg;  ale = mean_t(-sum_c x ln x);  epi = total - ale
// The sampler behind torch.distributions.Dirichlet is ATen's _sample_dirichlet (third party):
// Marsaglia-Tsang gamma draws with the alpha<1 boost, normalised and clamped to
// [FLT_MIN, 1 - 2^-24].  This kernel is written from the published algorithm (Marsaglia & Tsang
// 2000), works in log space so tiny alphas do not underflow, draws its randomness from a
// counter-based Philox4x32-10 keyed by (seed; image id, row, object, class, sample, attempt) and
// never writes a sample to global memory: one warp owns a pair, lane = sample, the per-sample
// log-gammas live in shared memory, class means are reduced with warp shuffles.
#pragma once
#include "common.cuh"

namespace mehhua {

constexpr int kK2Threads = 256;
constexpr int kK2Warps = kK2Threads / 32;
constexpr float kFltMin = 1.17549435e-38f;
constexpr float kTopClamp = 0.99999994f;          // 1 - 2^-24
constexpr float kLnFltMin = -87.33654475f;        // ln(FLT_MIN)
constexpr float kLnTopClamp = -5.9604646e-08f;    // ln(1 - 2^-24)

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// uniform in (0,1) with 24-bit resolution, never 0 or 1
__device__ __forceinline__ float u24(unsigned w) { return ((float)(w >> 8) + 0.5f) * 5.9604644775390625e-08f; }

// per-warp shared-memory carve-up (floats): lbuf[C*32] | alpha[C] | dd[C] | cc[C] | lnd[C] | inva[C] | avg[C]
__host__ __device__ inline size_t k2_smem_bytes(int C) { return (size_t)kK2Warps * (C * 32 + 6 * C) * sizeof(float) + 1040 * sizeof(int); }

__global__ void __launch_bounds__(kK2Threads)
k2_dirichlet_kernel(const __grid_constant__ Plan p, const float* __restrict__ score_rows,
                    const float* __restrict__ lam_rows, const float* __restrict__ lam_mean,
                    const int* __restrict__ pair_row, const int* __restrict__ pair_obj,
                    const int* __restrict__ pair_off, const long long* __restrict__ image_ids,
                    const float* __restrict__ inj, const long long* __restrict__ inj_off,
                    float* __restrict__ pair_unc, unsigned* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char k2_smem[];
  const int C = p.C, T = p.n_samples;
  int* img_pref = reinterpret_cast<int*>(k2_smem);            // [B+1] exclusive prefix of pair counts
  float* wbase = reinterpret_cast<float*>(img_pref + 1040) + (size_t)(threadIdx.x >> 5) * (C * 32 + 6 * C);
  float* lbuf = wbase;
  float* s_alpha = lbuf + C * 32;
  float* s_dd = s_alpha + C;
  float* s_cc = s_dd + C;
  float* s_lnd = s_cc + C;
  float* s_inva = s_lnd + C;
  float* s_avg = s_inva + C;

  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < p.B; ++b) { img_pref[b] = acc; acc += pair_off[b * (p.S + 1) + p.S]; }
    img_pref[p.B] = acc;
  }
  __syncthreads();
  const int total = img_pref[p.B];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kK2Warps + (threadIdx.x >> 5);
  const int nw = gridDim.x * kK2Warps;
  const uint2 key = make_uint2((unsigned)(p.seed & 0xffffffffull), (unsigned)(p.seed >> 32));

  for (int g = gw; g < total; g += nw) {
    // locate (image, pair) by binary search in the prefix
    int lo = 0, hi = p.B;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (img_pref[mid] <= g) lo = mid; else hi = mid; }
    const int b = lo, q = g - img_pref[b];
    const size_t pq = (size_t)b * p.pair_cap + q;
    const int row = pair_row[pq], obj = pair_obj[pq];
    const int s = level_of_row(p, row);
    float lamp = 1.f;
    if (p.use_lambda) {
      const float lam = lam_rows[(size_t)b * p.K + row];
      lamp = __fmul_rn(__fdiv_rn(lam_mean[b * p.S + s], __fadd_rn(lam, p.lambda_eps)), p.lambda_scale);
    }
    const float* srow = score_rows + ((size_t)b * p.K + row) * C;
    __syncwarp();
    for (int c = lane; c < C; c += 32) {
      float a = __fmul_rn(srow[c], lamp);
      if (!(a > 0.f) || !(a < 3.0e38f)) { a = 0.f; atomicOr(status, MEHHUA_ST_BAD_ALPHA); }
      const float ash = a < 1.f ? a + 1.f : a;          // Marsaglia-Tsang shape (boosted when alpha < 1)
      const float d = ash - (1.f / 3.f);
      s_alpha[c] = a;
      s_dd[c] = d;
      s_cc[c] = rsqrtf(9.f * d);
      s_lnd[c] = logf(d);
      s_inva[c] = a > 0.f ? __fdiv_rn(1.f, a) : 0.f;
      s_avg[c] = 0.f;
    }
    __syncwarp();

    float ent_acc = 0.f;   // sum over this lane's samples of sum_c x ln x
    const long long ioff = (inj != nullptr && inj_off != nullptr) ? inj_off[b * p.S + s] : -1;
    if (ioff >= 0) {
      // ---- injection mode: consume the oracle's drawn samples [T, P_bs, C] ----
      const int pb = pair_off[b * (p.S + 1) + s];
      const int P = pair_off[b * (p.S + 1) + s + 1] - pb;
      const int ql = q - pb;
      for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        const bool active = t < T;
        const float* xs = inj + ioff + ((size_t)(active ? t : 0) * P + ql) * C;
        for (int c = 0; c < C; ++c) {
          float x = active ? __ldg(xs + c) : 0.f;
          if (active) ent_acc = __fmaf_rn(x, logf(x), ent_acc);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(full, x, o);
          if (lane == 0) s_avg[c] += x;
        }
      }
    } else {
      // ---- free-running sampler ----
      const unsigned gid = image_ids ? (unsigned)image_ids[b] : (unsigned)b;
      const unsigned pid = (unsigned)row | ((unsigned)obj << 20);
      for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        const bool active = t < T;
        int c = active ? 0 : C;
        unsigned att = 0;
        // flattened rejection loop: every iteration each unfinished lane makes one attempt for
        // its current class, so lanes never idle while a neighbour retries
        while (__any_sync(full, c < C)) {
          if (c < C) {
            const float a = s_alpha[c];
            if (a <= 0.f) {
              lbuf[c * 32 + lane] = -INFINITY;
              ++c;
            } else {
              const uint4 w = philox4x32_10(make_uint4((unsigned)t, (unsigned)c | (att << 16), pid, gid), key);
              const float r = sqrtf(-2.f * __logf(u24(w.x)));
              const float x = r * __cosf((float)w.y * 1.4629180792671596e-09f);   // 2*pi / 2^32
              const float v1 = fmaf(s_cc[c], x, 1.f);
              bool ok = false;
              float lv = 0.f;
              if (v1 > 0.f) {
                lv = 3.f * __logf(v1);
                const float v = v1 * v1 * v1;
                ok = __logf(u24(w.z)) < fmaf(s_dd[c], 1.f - v + lv, 0.5f * x * x);
              }
              if (ok) {
                float l = s_lnd[c] + lv;
                if (a < 1.f) l = fmaf(__logf(u24(w.w)), s_inva[c], l);
                lbuf[c * 32 + lane] = l;
                ++c;
                att = 0;
              } else {
                ++att;
              }
            }
          }
        }
        __syncwarp();
        // normalise in log space: z = ln x = l - (max + ln sum exp(l - max))
        float m = -INFINITY;
        for (int cc = 0; cc < C; ++cc) m = fmaxf(m, lbuf[cc * 32 + lane]);
        if (!(m > -INFINITY)) m = 0.f;
        float A = 0.f;
        for (int cc = 0; cc < C; ++cc) A += ex2_approx((lbuf[cc * 32 + lane] - m) * kLog2e);
        const float zoff = m + __logf(A);
        for (int cc = 0; cc < C; ++cc) {
          const float z = lbuf[cc * 32 + lane] - zoff;
          float x = ex2_approx(z * kLog2e);
          x = fminf(fmaxf(x, kFltMin), kTopClamp);
          const float zc = fminf(fmaxf(z, kLnFltMin), kLnTopClamp);
          if (!active) x = 0.f;
          else ent_acc = fmaf(x, zc, ent_acc);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(full, x, o);
          if (lane == 0) s_avg[cc] += x;
        }
        __syncwarp();
      }
    }
    __syncwarp();
    // epilogue: total = -sum_c avg ln avg, ale = -(1/T) sum_t sum_c x ln x
    float tot = 0.f;
    const float fT = (float)T;
    for (int c = lane; c < C; c += 32) {
      const float avg = __fdiv_rn(s_avg[c], fT);
      tot = __fmaf_rn(-avg, logf(avg), tot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tot += __shfl_xor_sync(full, tot, o);
      ent_acc += __shfl_xor_sync(full, ent_acc, o);
    }
    if (lane == 0) {
      const float ale = -__fdiv_rn(ent_acc, fT);
      float* o3 = pair_unc + pq * 3;
      o3[0] = tot; o3[1] = ale; o3[2] = tot - ale;
    }
  }
}

// debug / known-answer entry: one Philox4x32-10 block
__global__ void philox_kat_kernel(uint4 ctr, uint2 key, unsigned* out) {
  const uint4 r = philox4x32_10(ctr, key);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

}  // namespace mehhua
