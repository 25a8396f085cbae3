// K2: Dirichlet-sampled epistemic uncertainty per (box, object) pair.
//
// Reference semantics (mmdet/models/dense_heads/Lambda_L2.py:513-525):
//   lambda' = mean(lambda_pairs) / (lambda + 1e-7) * 25;  alpha = score_row * lambda'
//   x_t ~ Dirichlet(alpha), t = 1..T (T = 500)
//   avg = mean_t x_t;  total = -sum_c avg ln avg;  ale = mean_t(-sum_c x ln x);  epi = total - ale
// The sampler behind torch.distributions.Dirichlet is ATen's _sample_dirichlet (third party): gamma
// draws normalised by their sum and clamped to [FLT_MIN, 1 - 2^-24].  This kernel is written from
// the published algorithms and is exact in distribution, not bit-compatible with torch's stream:
//   alpha >= 1 : Marsaglia & Tsang (2000) squeeze-free form, normal by Box-Muller
//   alpha <  1 : Ahrens & Dieter (1974) algorithm GS - two uniforms, no normal, acceptance -> 1 as
//                alpha -> 0, which is where almost every class of a softmax row lives
// Everything is kept in LOG space (ln g), so alpha ~ 1e-5 draws (g ~ e^-100000) neither underflow
// nor need special casing; the reference's FLT_MIN clamp only changes terms below 1e-36.
// Randomness: counter-based Philox4x32-10 keyed by the seed, counter = (sample, call index, pair
// identity (row, object), global image id) - independent of batch composition and world size.
// Layout: one warp owns a pair; lane = sample (32 at a time).  The scaled draws e_c = g_c / 2^m of
// the warp's 32 samples sit in shared memory as lbuf[class][lane] in bfloat16 (round-to-nearest;
// row stride 34 halves = 17 words: conflict-free both for the lane-private stores of the draw loop
// and for the transposed class-sum pass).  Only the class means are formed from the bf16 copies
// (relative rounding error 2^-9 per term, unbiased, averaged over T samples - two orders below the
// Monte-Carlo error of the estimator); the per-sample sum A and the entropy terms stay in fp32
// registers.  Half-width rows are what lets three blocks (24 warps) share an SM.  Samples never touch
// global memory.
#pragma once
#include "common.cuh"

namespace mehhua {

constexpr int kK2Threads = 256;
constexpr int kK2Warps = kK2Threads / 32;
constexpr int kLStride = 34;                          // bf16 elements per class row (17 words)
constexpr float kFltMin = 1.17549435e-38f;
constexpr float kInvE = 0.36787944117144233f;
// The T samples of a pair are accumulated in kK2Sub fixed sub-ranges (whole 32-sample rounds) whose
// partial sums are combined in sub-range order.  The arithmetic is the same whether one warp walks
// all sub-ranges or - when a launch has too few pairs to fill the GPU (the reference's own batch
// sizes of 2 and 8 images) - kK2Sub warps take one each and the last to finish combines them, so a
// score does not depend on the batch it was computed in.
constexpr int kK2Sub = 4;
constexpr int kK2SplitPairs = 8192;                  // launches with at most this many pairs are split ...
constexpr int kK2SplitMaxBatch = 64;                 // ... when the batch is small enough for that to be likely
__host__ __device__ inline size_t k2_part_floats(int C) { return (size_t)kK2SplitPairs * kK2Sub * ((size_t)C + 1); }

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

// shared-memory accessors on 32-bit shared-window addresses (keeps address arithmetic in one IMAD)
__device__ __forceinline__ float4 lds_v4(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_bf16(unsigned a, float v) {      // round-to-nearest-even fp32 -> bf16
  asm volatile("{ .reg .b16 h; cvt.rn.bf16.f32 h, %1; st.shared.b16 [%0], h; }" :: "r"(a), "f"(v));
}
__device__ __forceinline__ unsigned lds_u8(unsigned a) { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

// per-warp shared memory, in floats:
//   cst4[C+4] (float4: b, 1/alpha, alpha-1, -b; in small-list order) | lbuf[C*17 words] | alpha[C] | avg[C] | part[C] |
//   lists 3 x (C+4) bytes.  The small list and its constants carry 4 benign pad entries, so a cursor that has
//   run past the end (by at most 3) still reads valid memory and needs no clamp.
__host__ __device__ inline size_t k2_warp_floats(int C) {
  const size_t f = 4 * ((size_t)C + 4) + (size_t)C * (kLStride / 2) + 3 * (size_t)C + (3 * ((size_t)C + 4) + 3) / 4;
  return (f + 3) & ~(size_t)3;
}
// + the per-image pair-count prefix [B+1], padded to 16 ints
__host__ __device__ inline int k2_pref_ints(int B) { return (B + 1 + 15) & ~15; }
__host__ __device__ inline size_t k2_smem_bytes(int C, int B) {
  return kK2Warps * k2_warp_floats(C) * sizeof(float) + k2_pref_ints(B) * sizeof(int);
}

// One Ahrens-Dieter GS attempt for the class at position i of the small-alpha list, as straight-line
// code (no branches, so the attempts of a lane's cursors interleave in the pipelines).  Works in
// log2 units; returns whether the draw is accepted, its class and l2 = log2(gamma draw).
//   p = b*U1;  p <= 1: x = p^(1/alpha), accept iff U2 <= exp(-x)
//              p >  1: x = -ln((b-p)/alpha) >= 1, accept iff U2 <= x^(alpha-1)
// both tests are done as log2(U2) <= rhs.
__device__ __forceinline__ bool gs_attempt(const int i, const int nsmall, const float f0, const float f1,
                                           const unsigned s_small, const unsigned s_cst4, unsigned& c, float& l2) {
  // f0, f1 in [1,2): U1 = f0 - 1 = (k + 1/2) / 2^16, U2 = f1 - 1 + 2^-17 in (0,1) (16-bit each)
  const unsigned ii = (unsigned)i;                      // <= nsmall + 3: pad entries
  c = lds_u8(s_small + ii);
  const float4 k = lds_v4(s_cst4 + ii * 16u);           // b, 1/alpha, alpha-1, -b (list order)
  const float pp = fmaf(f0, k.x, k.w);                  // b * U1
  const bool lo = pp <= 1.f;
  const float q = lo ? pp : (k.x - pp) * k.y;
  const float lq = lg2_approx(q);
  const float l2a = lq * k.y;                           // log2 x, first branch
  const float rhsa = -kLog2e * ex2_approx(l2a);         // log2 exp(-x)
  const float l2b = lg2_approx(-kLn2 * lq);             // log2 x, second branch
  const float rhsb = k.z * l2b;                         // log2 x^(alpha-1)
  l2 = lo ? l2a : l2b;
  const float rhs = lo ? rhsa : rhsb;
  return (i < nsmall) & (lg2_approx(f1 - 0.99999237060546875f) <= rhs);   // f1 - (1 - 2^-17)
}

// One Philox4x32 word -> one (U1, U2) pair, 16 bits each, as floats in [1,2) built straight from
// the bits (no int->float conversion on the SFU pipe): the 16 bits go to the top of the mantissa,
// U1 gets a half-step offset so that b*U1 is never 0.
__device__ __forceinline__ void split_word(const unsigned w, float& f0, float& f1) {
  f0 = __uint_as_float(0x3f800040u | ((w >> 9) & 0x7fff80u));     // w[31:16]
  f1 = __uint_as_float(0x3f800000u | ((w << 7) & 0x7fff80u));     // w[15:0]
}

// SPLIT = false: one work item per pair (many pairs, or injected samples); SPLIT = true: one item per
// (pair, sub-range).  Both instantiations are launched; the one whose regime does not apply returns
// at once (the pair count is only known on the device).
template <bool SPLIT>
__global__ void __launch_bounds__(kK2Threads, 3)
k2_dirichlet_kernel(const __grid_constant__ Plan p, const float* __restrict__ score_rows,
                    const float* __restrict__ lam_rows, const float* __restrict__ lam_mean,
                    const int* __restrict__ pair_row, const int* __restrict__ pair_obj,
                    const int* __restrict__ pair_off, const long long* __restrict__ image_ids,
                    const float* __restrict__ inj, const long long* __restrict__ inj_off,
                    float* __restrict__ pair_unc, int* __restrict__ work_counter,
                    float* __restrict__ part, int* __restrict__ done, const int allow_split,
                    unsigned* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char k2_smem[];
  const int C = p.C, T = p.n_samples;
  int* img_pref = reinterpret_cast<int*>(k2_smem);            // [B+1] exclusive prefix of pair counts
  float* wbase = reinterpret_cast<float*>(img_pref + k2_pref_ints(p.B)) + (size_t)(threadIdx.x >> 5) * k2_warp_floats(C);
  float4* cst4 = reinterpret_cast<float4*>(wbase);            // [C]
  unsigned* lbuf = reinterpret_cast<unsigned*>(wbase + 4 * (C + 4));   // [C][17 words] = [C][34] bf16
  float* s_alpha = wbase + 4 * (C + 4) + C * (kLStride / 2);  // [C]
  float* s_avg = s_alpha + C;                                 // [C]
  float* s_part = s_avg + C;                                  // [C] class sums of the current sub-range
  unsigned char* s_small = reinterpret_cast<unsigned char*>(s_part + C);
  unsigned char* s_big = s_small + C + 4;
  unsigned char* s_bad = s_big + C + 4;
  const int lane = threadIdx.x & 31;
  const unsigned a_cst4 = (unsigned)__cvta_generic_to_shared(cst4);
  const unsigned a_small = (unsigned)__cvta_generic_to_shared(s_small);
  const unsigned a_lrow = (unsigned)__cvta_generic_to_shared(lbuf) + 2u * lane;   // my column of row 0

  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < p.B; ++b) { img_pref[b] = acc; acc += pair_off[b * (p.S + 1) + p.S]; }
    img_pref[p.B] = acc;
  }
  __syncthreads();
  const int total = img_pref[p.B];
  const unsigned full = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  const uint2 key = make_uint2((unsigned)(p.seed & 0xffffffffull), (unsigned)(p.seed >> 32));
  const float fT = (float)T;
  // sub-ranges of the sample index: whole rounds of 32, kK2Sub of them
  const int sub_len = (((T + 31) / 32 + kK2Sub - 1) / kK2Sub) * 32;
  // few pairs and no injected samples: one work item per (pair, sub-range) instead of per pair
  const bool split_regime = (allow_split != 0 && inj == nullptr && total <= kK2SplitPairs);
  if (split_regime != SPLIT) return;
  constexpr int nsplit = SPLIT ? kK2Sub : 1;
  const int items = total * nsplit;

  for (;;) {
    int g = 0;
    if (lane == 0) g = atomicAdd(work_counter, 1);
    g = __shfl_sync(full, g, 0);
    if (g >= items) break;
    const int gp = SPLIT ? g % total : g;                    // pair (index over the whole launch)
    const int sub0 = SPLIT ? g / total : 0;                  // the sub-range of this item (split form)
    g = gp;
    // locate (image, pair) by binary search in the prefix
    int lo = 0, hi = p.B;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (img_pref[mid] <= g) lo = mid; else hi = mid; }
    const int b = lo, q = g - img_pref[b];
    const size_t pq = (size_t)b * p.pair_cap + q;
    const int row = pair_row[pq], obj = pair_obj[pq];
    const int s = level_of_pair(pair_off + b * (p.S + 1), p.S, q);
    float lamp = 1.f;
    if (p.use_lambda) {
      const float lam = lam_rows[(size_t)b * p.row_stride + row];
      lamp = __fmul_rn(__fdiv_rn(lam_mean[b * p.S + s], __fadd_rn(lam, p.lambda_eps)), p.lambda_scale);
    }
    const float* srow = score_rows + ((size_t)b * p.row_stride + row) * C;
    __syncwarp();
    // class lists: big (alpha >= 1), small (0 < alpha < 1), bad (alpha <= 0, denormal-tiny or non-finite)
    int nsmall = 0, nbig = 0, nbad = 0;
    float amax = 0.f;
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int c = c0 + lane;
      float a = 0.f;
      if (c < C) a = __fmul_rn(srow[c], lamp);
      const bool valid = c < C;
      const bool bad = valid && (!(a > 1e-30f) || !(a < 3.0e38f));
      const bool big = valid && !bad && a >= 1.f;
      const bool small = valid && !bad && a < 1.f;
      const unsigned mb = __ballot_sync(full, big), ms = __ballot_sync(full, small), mx = __ballot_sync(full, bad);
      if (big) { s_big[nbig + __popc(mb & lt_mask)] = (unsigned char)c; amax = fmaxf(amax, a); }
      if (small) {      // list entry = class id; the GS constants sit at the same list position
        const int pos = nsmall + __popc(ms & lt_mask);
        const float bb = fmaf(a, kInvE, 1.f);
        s_small[pos] = (unsigned char)c;
        cst4[pos] = make_float4(bb, __fdiv_rn(1.f, a), a - 1.f, -bb);
      }
      if (bad) s_bad[nbad + __popc(mx & lt_mask)] = (unsigned char)c;
      nbig += __popc(mb); nsmall += __popc(ms); nbad += __popc(mx);
      if (valid) {
        s_alpha[c] = bad ? 0.f : a;
        s_avg[c] = 0.f;
      }
    }
    if (lane < 4) {       // pad entries read by finished cursors (their attempts are discarded)
      s_small[nsmall + lane] = 0;
      cst4[nsmall + lane] = make_float4(1.f, 1.f, 0.f, -1.f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(full, amax, o));
    const float m = nbig > 0 ? lg2_approx(amax) : 0.f;     // reference exponent of the scaled draws
    if (nbad > 0 && lane == 0) atomicOr(status, MEHHUA_ST_BAD_ALPHA);
    __syncwarp();

    float ent_acc = 0.f;   // sum over this lane's samples of (-sum_c x ln x)
    const long long ioff = (inj != nullptr && inj_off != nullptr) ? inj_off[b * p.S + s] : -1;
    if (ioff >= 0) {
      // ---- injection mode: consume the oracle's drawn samples [T, P_bs, C]; reference formulas ----
      const int pb = pair_off[b * (p.S + 1) + s];
      const int P = pair_off[b * (p.S + 1) + s + 1] - pb;
      const int ql = q - pb;
      for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        const bool active = t < T;
        const float* xs = inj + ioff + ((size_t)(active ? t : 0) * P + ql) * C;
        for (int c = 0; c < C; ++c) {
          float x = active ? __ldg(xs + c) : 0.f;
          if (active) ent_acc = __fmaf_rn(-x, logf(x), ent_acc);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(full, x, o);
          if (lane == 0) s_avg[c] += x;
        }
      }
    } else {
      // ---- free-running sampler (log2 units throughout) ----
      const unsigned gid = image_ids ? (unsigned)image_ids[b] : (unsigned)b;
      const unsigned pid = (unsigned)row | ((unsigned)obj << 20);
      // one flat loop over the rounds of this item; the per-pair form folds the running sub-range into
      // the totals whenever a sub-range boundary (multiple of sub_len) is crossed
      for (int c = lane; c < C; c += 32) s_part[c] = 0.f;
      float ent_sub = 0.f;
      __syncwarp();
      const int t_end = SPLIT ? min(T, (sub0 + 1) * sub_len) : T;
      int fold_at = sub_len;
      for (int t0 = SPLIT ? sub0 * sub_len : 0; t0 < t_end; t0 += 32) {
        if (!SPLIT && t0 == fold_at) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) ent_sub += __shfl_xor_sync(full, ent_sub, o);
          ent_acc += ent_sub;
          ent_sub = 0.f;
          for (int c = lane; c < C; c += 32) { s_avg[c] += s_part[c]; s_part[c] = 0.f; }
          fold_at += sub_len;
          __syncwarp();
        }
        const int t = t0 + lane;
        const bool active = t < T;
        // Normalisation is a log-sum-exp around a reference exponent m fixed BEFORE the draws:
        // log2 of the largest alpha when some alpha >= 1 (its draw owns the sample and is within a
        // few octaves of alpha), 0 otherwise.
        //   e_c = 2^(l_c - m),  A = sum_c e_c,  x_c = e_c / A,
        //   -sum_c x ln x = ln A - ln2 * (sum_c e_c d_c) / A   with d_c = l_c - m (log2 units)
        // Any m gives the same value; fixing it early lets every accepted draw be folded into A and
        // the entropy sum on the spot, so the draws are stored as e_c and read back only once.
        float asum = 0.f, bs = 0.f;
        if (active) {
          for (int i = 0; i < nbad; ++i) sts_bf16(a_lrow + s_bad[i] * (kLStride * 2u), 0.f);
          // Marsaglia-Tsang for the few classes with alpha >= 1
          for (int i = 0; i < nbig; ++i) {
            const int c = s_big[i];
            const float d = s_alpha[c] - (1.f / 3.f);
            const float cc = rsqrtf(9.f * d);
            float l2;
            for (unsigned att = 0;; ++att) {
              const uint4 w = philox4x32_10(make_uint4((unsigned)t, 0x80000000u | ((unsigned)c << 8) | (att & 255u), pid, gid), key);
              const float r = sqrtf(-2.f * kLn2 * lg2_approx(u24(w.x)));
              const float x = r * __cosf((float)w.y * 1.4629180792671596e-09f);   // 2*pi / 2^32
              const float v1 = fmaf(cc, x, 1.f);
              if (v1 > 0.f) {
                const float lv2 = 3.f * lg2_approx(v1);
                const float v = v1 * v1 * v1;
                if (kLn2 * lg2_approx(u24(w.z)) < fmaf(d, fmaf(kLn2, lv2, 1.f - v), 0.5f * x * x)) { l2 = lg2_approx(d) + lv2; break; }
              }
            }
            const float dl = l2 - m;
            const float e = ex2_approx(dl);
            asum += e;
            bs = fmaf(e, dl, bs);
            sts_bf16(a_lrow + c * (kLStride * 2u), e);
          }
        } else {
          unsigned addr = a_lrow;
          for (int c = 0; c < C; ++c, addr += kLStride * 2u) sts_bf16(addr, 0.f);
        }
        // Ahrens-Dieter GS for alpha < 1.  Flattened rejection loop with four cursors (list positions
        // mod 4): each iteration one Philox block (4 words = 4 x (16 + 16) bits) feeds one attempt per
        // cursor, so a lane never idles while a neighbour retries and the attempts overlap in the pipes.
        int ia = active ? 0 : nsmall, ib = active ? 1 : nsmall, ic = active ? 2 : nsmall, id = active ? 3 : nsmall;
        unsigned kcall = 0;
        if (nsmall > 0) {
          while (__any_sync(full, (ia < nsmall) | (ib < nsmall) | (ic < nsmall) | (id < nsmall))) {
            const uint4 w = philox4x32_10(make_uint4((unsigned)t, kcall++, pid, gid), key);
            float f0[4], f1[4];
            split_word(w.x, f0[0], f1[0]);
            split_word(w.y, f0[1], f1[1]);
            split_word(w.z, f0[2], f1[2]);
            split_word(w.w, f0[3], f1[3]);
            unsigned ca, cb, cc, cd;
            float la, lb, lc, ld;
            const bool oka = gs_attempt(ia, nsmall, f0[0], f1[0], a_small, a_cst4, ca, la);
            const bool okb = gs_attempt(ib, nsmall, f0[1], f1[1], a_small, a_cst4, cb, lb);
            const bool okc = gs_attempt(ic, nsmall, f0[2], f1[2], a_small, a_cst4, cc, lc);
            const bool okd = gs_attempt(id, nsmall, f0[3], f1[3], a_small, a_cst4, cd, ld);
            // l is finite (U1 > 0, alpha >= 1e-30), so d = l - m needs no clamp: e = 0 below 2^-149 and 0 * d = 0
            const float da = la - m, db = lb - m, dc = lc - m, dd = ld - m;
            const float ea = ex2_approx(da), eb = ex2_approx(db), ec = ex2_approx(dc), ed = ex2_approx(dd);
            if (oka) { sts_bf16(a_lrow + ca * (kLStride * 2u), ea); asum += ea; bs = fmaf(ea, da, bs); ia += 4; }
            if (okb) { sts_bf16(a_lrow + cb * (kLStride * 2u), eb); asum += eb; bs = fmaf(eb, db, bs); ib += 4; }
            if (okc) { sts_bf16(a_lrow + cc * (kLStride * 2u), ec); asum += ec; bs = fmaf(ec, dc, bs); ic += 4; }
            if (okd) { sts_bf16(a_lrow + cd * (kLStride * 2u), ed); asum += ed; bs = fmaf(ed, dd, bs); id += 4; }
          }
        }
        float inv_a = 0.f;
        if (active && asum > 0.f) {
          inv_a = __fdividef(1.f, asum);
          ent_sub += logf(asum) - kLn2 * bs * inv_a;
        }
        __syncwarp();
        // class sums over the 32 samples, transposed: lane = class, walk the samples two at a time
        // (one word = the bf16 copies of samples 2w and 2w+1; bf16 -> fp32 is a shift / a mask)
        for (int c0 = 0; c0 < C; c0 += 128) {
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          const int c = c0 + lane;
          const unsigned* r0p = lbuf + (size_t)min(c, C - 1) * (kLStride / 2);
          const unsigned* r1p = lbuf + (size_t)min(c + 32, C - 1) * (kLStride / 2);
          const unsigned* r2p = lbuf + (size_t)min(c + 64, C - 1) * (kLStride / 2);
          const unsigned* r3p = lbuf + (size_t)min(c + 96, C - 1) * (kLStride / 2);
          const int rem = C - c0;
#pragma unroll 4
          for (int w = 0; w < 16; ++w) {
            const float iv0 = __shfl_sync(full, inv_a, 2 * w), iv1 = __shfl_sync(full, inv_a, 2 * w + 1);
            unsigned u = r0p[w];
            a0 = fmaf(__uint_as_float(u << 16), iv0, a0);
            a0 = fmaf(__uint_as_float(u & 0xffff0000u), iv1, a0);
            if (rem > 32) { u = r1p[w]; a1 = fmaf(__uint_as_float(u << 16), iv0, a1); a1 = fmaf(__uint_as_float(u & 0xffff0000u), iv1, a1); }
            if (rem > 64) { u = r2p[w]; a2 = fmaf(__uint_as_float(u << 16), iv0, a2); a2 = fmaf(__uint_as_float(u & 0xffff0000u), iv1, a2); }
            if (rem > 96) { u = r3p[w]; a3 = fmaf(__uint_as_float(u << 16), iv0, a3); a3 = fmaf(__uint_as_float(u & 0xffff0000u), iv1, a3); }
          }
          if (c < C) s_part[c] += a0;
          if (c + 32 < C) s_part[c + 32] += a1;
          if (c + 64 < C) s_part[c + 64] += a2;
          if (c + 96 < C) s_part[c + 96] += a3;
        }
        __syncwarp();
      }
      // close the (last) sub-range: its class sums and entropy sum, combined in sub-range order
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ent_sub += __shfl_xor_sync(full, ent_sub, o);
      if constexpr (!SPLIT) {
        for (int c = lane; c < C; c += 32) s_avg[c] += s_part[c];
        ent_acc += ent_sub;
      } else {
        float* dst = part + ((size_t)gp * kK2Sub + sub0) * (C + 1);
        for (int c = lane; c < C; c += 32) __stcg(dst + c, s_part[c]);
        if (lane == 0) __stcg(dst + C, ent_sub);
      }
      __syncwarp();
      if constexpr (SPLIT) {
        // the last of the pair's kK2Sub items to arrive combines the partial sums (in sub-range order)
        __threadfence();
        int old = 0;
        if (lane == 0) old = atomicAdd(done + gp, 1);
        old = __shfl_sync(full, old, 0);
        if (old != nsplit - 1) continue;
        __threadfence();
        const float* src = part + (size_t)gp * kK2Sub * (C + 1);
        for (int c = lane; c < C; c += 32) {
          float v = 0.f;
#pragma unroll
          for (int u = 0; u < kK2Sub; ++u) v += __ldcg(src + (size_t)u * (C + 1) + c);
          s_avg[c] = v;
        }
        ent_acc = 0.f;
#pragma unroll
        for (int u = 0; u < kK2Sub; ++u) ent_acc += __ldcg(src + (size_t)u * (C + 1) + C);
      }
    }
    __syncwarp();
    // epilogue: total = -sum_c avg ln avg, ale = (1/T) sum_t (-sum_c x ln x)
    float tot = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float avg = fmaxf(__fdiv_rn(s_avg[c], fT), kFltMin);
      tot = __fmaf_rn(-avg, logf(avg), tot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tot += __shfl_xor_sync(full, tot, o);
      if (ioff >= 0) ent_acc += __shfl_xor_sync(full, ent_acc, o);   // injection: per-lane sums; else already reduced
    }
    if (lane == 0) {
      const float ale = __fdiv_rn(ent_acc, fT);
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
