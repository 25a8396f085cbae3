// K2: Dirichlet-sampled epistemic uncertainty per (box, object) pair.
//
// Reference semantics (mmdet/models/dense_heads/Lambda_L2.py:513-525):
//   lambda' = mean(lambda_pairs) / (lambda + 1e-7) * 25;  alpha = score_row * lambda'
//   x_t ~ Dirichlet(alpha), t = 1..T (T = 500)
//   avg = mean_t x_t;  total = -sum_c avg ln avg;  ale = mean_t(-sum_c x ln x);  epi = total - ale
// The sampler behind torch.distributions.Dirichlet is ATen's _sample_dirichlet (third party): gamma
// draws normalised by their sum and clamped to [FLT_MIN, 1 - 2^-24].  This kernel is written from
// the published algorithms; it is not bit-compatible with torch's stream.  Per class:
//   alpha >= 1 : Marsaglia & Tsang (2000), normals by Box-Muller (both of a pair are used), 23-bit
//                uniforms.
//   alpha <  1 : Ahrens & Dieter (1974) algorithm GS, written with a split point s (the original has
//                s = 1): envelope x^(alpha-1) on [0, s], s^(alpha-1) e^-x beyond.  With
//                b = 1 + alpha e^-s / s and p = b U1:
//                  p <= 1 : x = s p^(1/alpha),                     accept iff U2 <= e^-x
//                  p >  1 : x = s - ln V (V uniform),             accept iff U2 <= (x/s)^(alpha-1)
//     kGsTiny <= alpha < 1 ("GS list"): s = 1, both branches evaluated straight-line.  U1 is on a 2^-23
//                grid (midpoints), U2 on a 2^-19 grid (midpoints); enumerating those grids gives
//                |E[g]/alpha - 1| <= 2e-5 for alpha >= 1e-3 (tools/gs_grid_bias.py,
//                profiles/r2_gs_grid_bias.txt).
//     alpha < kGsTiny ("tiny list"): s = 2 (the p > 1 branch is then taken by 0.068 alpha of the
//                attempts) and U1 = (w + f) / 2^32 resolved to 2^-64 near the branch point.  A p <= 1
//                draw can only be non-zero in fp32 when log2 x - m >= -127, i.e. when the top word w
//                of U1 reaches a per-class threshold: one integer compare decides almost every draw
//                of a tiny-alpha class (the skipped draws are exactly the ones whose scaled value e is
//                flushed to 0 below, and whose acceptance probability e^-x is 1 on any grid), the few
//                others take the full formula with log1p arithmetic.  Exact for every alpha the path
//                can produce.
// Everything is kept in LOG space (log2 g), so tiny draws neither underflow nor need special casing;
// the reference's FLT_MIN clamp only changes terms below 1e-36.
// Normalisation: e_c = g_c / 2^m with a reference exponent m fixed per pair BEFORE the draws (log2 of
// the largest alpha when some alpha >= 1, else 0).  Rows whose alphas are all < 1 and sum to < 1
// (lambda' << 1) can have every draw of a sample far below 2^-126; they take a two-pass form: the
// counter-based generator replays the sample, pass 1 finds its largest log2 draw, pass 2 normalises
// around it.
// Randomness: counter-based Philox4x32 keyed by the seed, counter = (sample, call index, pair
// identity (row, object), global image id) - independent of batch composition and world size.
// MEHHUA_PHILOX_ROUNDS = 7 rounds by default: Philox4x32-7 is the Crush-resistant member of the family
// (Salmon et al., SC'11, table 2; Random123 ships it with known-answer vectors, checked in the tests
// next to the 10-round ones); -DMEHHUA_PHILOX_ROUNDS=10 builds the 10-round variant.
// Layout: one warp owns a pair; lane = sample (32 at a time).  The scaled draws e_c of the warp's 32
// samples sit in shared memory as lbuf[class][lane] in fp32 (row stride 33 words: conflict-free both
// for the lane-private stores of the draw loop and for the transposed class-sum pass); the per-sample
// sum A and the entropy terms stay in registers.  Samples never touch global memory.
// (-DMEHHUA_K2_STAGE_BF16 builds a variant that stages the draws as bfloat16 - half the shared memory,
// three blocks per SM; it is not the shipped build.)
// n_samples == 0 selects the ANALYTIC form (T -> infinity): total = H(alpha/alpha0),
// ale = psi(alpha0 + 1) - sum_c (alpha_c/alpha0) psi(alpha_c + 1), in double precision - a
// deterministic mode for pool-level set-identity tests against the oracle's closed forms.
#pragma once
#include "common.cuh"

#ifndef MEHHUA_PHILOX_ROUNDS
#define MEHHUA_PHILOX_ROUNDS 7
#endif

namespace mehhua {

#ifndef MEHHUA_K2_THREADS
#define MEHHUA_K2_THREADS 256
#endif
constexpr int kK2Threads = MEHHUA_K2_THREADS;
constexpr int kK2Warps = kK2Threads / 32;
#ifdef MEHHUA_K2_STAGE_BF16
constexpr int kLWords = 17;                           // bf16 staging: 34 halves per class row
constexpr int kK2MinBlocks = 3;
#else
constexpr int kLWords = 33;                           // fp32 staging: 32 samples + 1 pad word per class row
constexpr int kK2MinBlocks = 512 / MEHHUA_K2_THREADS;
#endif
constexpr float kFltMin = 1.17549435e-38f;
constexpr float kGsTiny = 1e-3f;                      // below: the tiny list (threshold compare + full-resolution draws)
constexpr float kEm2Half = 0.067667641618306351f;     // e^-2 / 2 (tiny list: s = 2)
constexpr float kEm1 = 0.36787944117144233f;          // e^-1 / 1 (GS list: s = 1)
constexpr float kTinySpan = 130.f;                    // a p<=1 draw is non-zero only when log2 x - m >= -(kTinySpan - log2 s): 126 + margin
#ifndef MEHHUA_K2_UNROLL
#define MEHHUA_K2_UNROLL 2
#endif
constexpr int kK2Unroll = MEHHUA_K2_UNROLL;
constexpr int kQCap = 8;                              // tiny rows a lane can remember having written in one round
// The T samples of a pair are accumulated in kK2Sub fixed sub-ranges (whole 32-sample rounds) whose
// partial sums are combined in sub-range order.  The arithmetic is the same whether one warp walks
// all sub-ranges or - when a launch has too few pairs to fill the GPU (the reference's own batch
// sizes of 2 and 8 images) - kK2Sub warps take one each and the last to finish combines them, so a
// score does not depend on the batch it was computed in.
constexpr int kK2Sub = 4;
constexpr int kK2SplitPairs = 8192;                  // launches with at most this many pairs are split ...
constexpr int kK2SplitMaxBatch = 64;                 // ... when the batch is small enough for that to be likely
__host__ __device__ inline size_t k2_part_floats(int C) { return (size_t)kK2SplitPairs * kK2Sub * ((size_t)C + 1); }

template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ uint4 k2_philox(uint4 c, uint2 k) { return philox4x32<MEHHUA_PHILOX_ROUNDS>(c, k); }

// uniform in (0,1) with 24-bit resolution, never 0 or 1
__device__ __forceinline__ float u24(unsigned w) { return ((float)(w >> 8) + 0.5f) * 5.9604644775390625e-08f; }

// shared-memory accessors on 32-bit shared-window addresses (keeps address arithmetic out of the hot loop)
__device__ __forceinline__ float4 lds_v4(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 lds_u4(unsigned a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned lds_u8(unsigned a) { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u8(unsigned a, unsigned v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v)); }
__device__ __forceinline__ unsigned lds_u16(unsigned a) { unsigned v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u16(unsigned a, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "r"(v)); }
// one staged draw: lbuf[class][lane]
__device__ __forceinline__ void stage_store(unsigned a, float v) {
#ifdef MEHHUA_K2_STAGE_BF16
  asm volatile("{ .reg .b16 h; cvt.rn.bf16.f32 h, %1; st.shared.b16 [%0], h; }" :: "r"(a), "f"(v));   // round-to-nearest-even
#else
  asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v));
#endif
}

// per-warp shared memory, in floats:
//   cst4[C+8] float4, two lists back to back, each followed by pad entries that can never be accepted
//   (so cursors that run past the end need no clamp):
//     GS list  (b, 1/alpha, byte offset of the class row in lbuf, alpha - 1) + 4 pads
//     big list (d = alpha - 1/3, 1/sqrt(9 d), byte offset of the class row, log2 d - m) + 2 pads
//   lbuf[C*kLWords] | alpha[C] | avg[C] | part[C] | thr[C+4 rounded to 4] (tiny-list order) |
//   dirty[32 lanes x kQCap halves] | tiny list (C+4) bytes.
__host__ __device__ inline size_t k2_thr_words(int C) { return ((size_t)C + 4 + 3) & ~(size_t)3; }
__host__ __device__ inline size_t k2_warp_floats(int C) {
  const size_t f = 4 * ((size_t)C + 8) + (size_t)C * kLWords + 3 * (size_t)C + k2_thr_words(C) + 32 * kQCap / 2 +
                   ((size_t)C + 4 + 3) / 4;
  return (f + 3) & ~(size_t)3;
}
// + the per-image pair-count prefix [B+1], padded to 16 ints
__host__ __device__ inline int k2_pref_ints(int B) { return (B + 1 + 15) & ~15; }
__host__ __device__ inline size_t k2_smem_bytes(int C, int B) {
  return kK2Warps * k2_warp_floats(C) * sizeof(float) + k2_pref_ints(B) * sizeof(int);
}

// What a warp knows about the pair it is working on.
struct K2Pair {
  unsigned a_cst4, a_big4, a_lrow, a_thr, a_park;   // shared-window addresses: GS / big constants, my lbuf column, tiny thresholds, my slots for dirty tiny rows
  const unsigned char* s_tiny;
  const float* s_alpha;
  int ngs, nbig, ntiny;
  bool tiny_always;                       // two-pass form: every tiny-class draw is evaluated
  unsigned pid, gid;
  unsigned one;                           // 0x3f800000, opaque to the compiler so that (bits & mask) | one is a single LOP3
  uint2 key;
};

// The draws of ONE sample (this lane's sample t) for every class of the pair.
//   PASS 0: normalise around the pair's fixed exponent m: stage e_c = 2^(l_c - m), accumulate
//           A = sum e and bs = sum e * (l_c - m).
//   PASS 1: only the largest log2 draw of the sample (mx).
//   PASS 2: as PASS 0 around a per-lane exponent m (the pair's exponent is 0 in the two-pass form).
// The three forms consume identical Philox counters and take identical decisions, so PASS 2 replays
// PASS 1 draw for draw.
template <int PASS>
__device__ __forceinline__ void k2_draw_sample(const K2Pair& W, const unsigned t, const bool active, const float m,
                                               float& asum, float& bs, float& mx, unsigned& dirty) {
  const unsigned full = 0xffffffffu;
  // a draw known by its log2 (l2) that belongs to the class row at byte offset `row` of lbuf
  auto fold_log = [&](const unsigned row, const float l2) {
    if (PASS == 1) { mx = fmaxf(mx, l2); return; }
    const float d = l2 - m;
    const float e = ex2_approx(d);          // exactly 0 below 2^-126 (ftz); l2 is finite, so 0 * d = 0
    stage_store(W.a_lrow + row, e);
    asum += e;
    bs = fmaf(e, d, bs);
  };
  // ---- big list (alpha >= 1): Marsaglia-Tsang.  Flattened rejection loop with two cursors fed from a
  // shared work list: one Philox block per iteration gives a Box-Muller pair of normals (one per
  // cursor) and the two acceptance uniforms; a cursor that accepts takes the next undrawn class.
  //   d = alpha - 1/3, c = 1/sqrt(9 d);  v = (1 + c x)^3;  accept iff v > 0 and ln U < x^2/2 + d - d v + d ln v;  g = d v
  if (W.nbig > 0) {
    const unsigned done_at = W.a_big4 + (unsigned)(W.nbig + 2) * 16u;
    const unsigned pad0 = W.a_big4 + (unsigned)W.nbig * 16u;
    unsigned ca = active ? W.a_big4 : pad0, cb = active ? W.a_big4 + 16u : pad0;
    unsigned nxt = active ? W.a_big4 + 32u : done_at;
    unsigned kcall = 0x80000000u;
    auto attempt = [&](unsigned& cur, float4& k, const float x, const unsigned wu) {
      // k = the class's constants: d, c, row offset, log2 d - m
      const float v1 = fmaf(k.y, x, 1.f);
      const float lv = lg2_approx(v1);                      // NaN for v1 < 0 (and for a pad: c = NaN): never accepted
      const float v = v1 * v1 * v1;
      const float lnu = kLn2 * lg2_approx(__uint_as_float(0x3f800000u | (wu >> 9)) - 0.99999994039535522f);
      float rhs = fmaf(0.5f * x, x, k.x);
      rhs = fmaf(-k.x, v, rhs);
      rhs = fmaf(3.f * kLn2 * k.x, lv, rhs);
      const float l2 = fmaf(3.f, lv, k.w);                  // log2 (d v) [- m]
      if (lnu < rhs) {
        if (PASS == 1) {
          mx = fmaxf(mx, l2);
        } else {
          const float d = (PASS == 0) ? l2 : l2 - m;
          const float e = ex2_approx(d);
          stage_store(W.a_lrow + __float_as_uint(k.z), e);
          asum += e;
          bs = fmaf(e, d, bs);
        }
        cur = nxt;
        nxt += 16u;
        k = lds_v4(cur);                                    // the next class's constants, fetched as soon as the cursor moves
      }
    };
    float4 ka = lds_v4(ca), kb = lds_v4(cb);
    while (__any_sync(full, nxt < done_at)) {
      const uint4 w = k2_philox(make_uint4(t, kcall++, W.pid, W.gid), W.key);
      const float r2 = (-2.f * kLn2) * lg2_approx(__uint_as_float(0x3f800000u | (w.x >> 9)) - 0.99999994039535522f);
      const float r = r2 * rsqrtf(r2);                      // sqrt(-2 ln U), U = (k + 1/2) / 2^23 < 1: r2 > 0
      float sn, cs;
      __sincosf((float)w.y * 1.4629180792671596e-09f, &sn, &cs);   // 2 pi / 2^32
      attempt(ca, ka, r * cs, w.z);
      attempt(cb, kb, r * sn, w.w);
    }
  }
  // ---- GS list (kGsTiny <= alpha < 1), split point s = 1.  Flattened rejection loop with three cursors
  // fed from a shared work list: each iteration one Philox block (4 words) feeds one attempt per
  // cursor - word j gives U1 (its top 23 bits) and the top of U2 (its low 9 bits), the fourth word
  // the rest of the three U2 (10 bits each) - and a cursor that accepts takes the next undrawn class, so
  // a lane never idles while a neighbour retries and the lanes of a warp finish within a few attempts
  // of each other.  Both branches are evaluated straight-line (no divergence):
  //   p = b U1 <= 1: log2 x = log2(p) / alpha,                         accept iff log2 U2 <= -log2(e) x
  //   p > 1        : x = -ln q, q = (b - p) / alpha (= V / e),         accept iff log2 U2 <= (alpha - 1) log2 x
  if (W.ngs > 0) {
    // cursors are shared-window addresses of cst4 entries (16 B each)
    const unsigned done_at = W.a_cst4 + (unsigned)(W.ngs + 3) * 16u;
    const unsigned pad0 = W.a_cst4 + (unsigned)W.ngs * 16u;
    unsigned ca = active ? W.a_cst4 : pad0, cb = active ? W.a_cst4 + 16u : pad0, cc = active ? W.a_cst4 + 32u : pad0;
    unsigned nxt = active ? W.a_cst4 + 48u : done_at;
    unsigned kcall = 0;
    const float kneg = (PASS == 0) ? -kLog2e * ex2_approx(m) : -kLog2e;   // PASS 0: log2 e^-x = kneg * e, x = e * 2^m
    const float sc = (PASS == 0) ? ex2_approx(-m) : 1.f;
    auto attempt = [&](unsigned& cur, float4& k, const unsigned wj, const unsigned u2bits) {
      // k = the class's constants: b, 1/alpha, row offset, alpha - 1
      const float pp = (__uint_as_float(0x3f800000u | (wj >> 9)) - 0.99999994039535522f) * k.x;   // b * U1, U1 = (k + 1/2) / 2^23
      const bool hi = pp > 1.f;                             // never for a pad (b = 1)
      const float q = hi ? (k.x - pp) * k.y : pp;
      const float lq = lg2_approx(q);
      const float lu2 = lg2_approx(__uint_as_float(u2bits) - 0.99999904632568359f);   // U2 = (j + 1/2) / 2^19
      const float xh = -kLn2 * lq;                          // p > 1: x
      const float l2h = lg2_approx(xh);                     //        log2 x
      float d, e, rhs;
      if (PASS == 0) {
        d = fmaf(lq, k.y, -m);                              // p <= 1: log2 x - m
        e = ex2_approx(d);
        rhs = e * kneg;
        if (hi) { d = l2h - m; e = xh * sc; rhs = k.w * l2h; }
      } else {
        d = hi ? l2h : lq * k.y;                            // log2 x
        rhs = hi ? k.w * l2h : kneg * ex2_approx(d);
        e = 0.f;
      }
      if (lu2 <= rhs) {                                     // false for a pad (1/alpha = NaN)
        if (PASS == 1) {
          mx = fmaxf(mx, d);
        } else {
          if (PASS == 2) { d -= m; e = ex2_approx(d); }
          stage_store(W.a_lrow + __float_as_uint(k.z), e);
          asum += e;
          bs = fmaf(e, d, bs);
        }
        cur = nxt;
        nxt += 16u;
        k = lds_v4(cur);                                    // the next class's constants, fetched as soon as the cursor moves
      }
    };
    float4 ka = lds_v4(ca), kb = lds_v4(cb), kc = lds_v4(cc);
    while (__any_sync(full, nxt < done_at)) {
      // the "anyone left?" vote is taken once per kK2Unroll Philox blocks
#pragma unroll
      for (int rep = 0; rep < kK2Unroll; ++rep) {
        const uint4 w = k2_philox(make_uint4(t, kcall++, W.pid, W.gid), W.key);
        // U2 mantissa bits 22..4 = [9 low bits of the word | 10 bits of w.w]
        attempt(ca, ka, w.x, (__funnelshift_l(w.w, w.x, 14) & 0x007ffff0u) | W.one);
        attempt(cb, kb, w.y, (__funnelshift_l(w.w << 10, w.y, 14) & 0x007ffff0u) | W.one);
        attempt(cc, kc, w.z, (__funnelshift_l(w.w << 20, w.z, 14) & 0x007ffff0u) | W.one);
      }
    }
  }
  // ---- tiny list: four classes per Philox block, one compare each; a hit takes the full formula
  for (int i0 = 0; i0 < W.ntiny; i0 += 4) {
    const uint4 w = k2_philox(make_uint4(t, 0x40000000u | (unsigned)(i0 >> 2), W.pid, W.gid), W.key);
    const uint4 th = lds_u4(W.a_thr + (unsigned)i0 * 4u);
    unsigned hm = (unsigned)(w.x >= th.x) | ((unsigned)(w.y >= th.y) << 1) | ((unsigned)(w.z >= th.z) << 2) |
                  ((unsigned)(w.w >= th.w) << 3);
    if (!active) hm = 0u;
#pragma unroll 1
    while (hm != 0u) {            // rare: a lane's hits, one at a time
      const int j = __ffs(hm) - 1;
      hm &= hm - 1u;
      if (i0 + j >= W.ntiny) break;                                         // pad entries of the last chunk
      unsigned wt = j == 0 ? w.x : (j == 1 ? w.y : (j == 2 ? w.z : w.w));
      const unsigned thr = j == 0 ? th.x : (j == 1 ? th.y : (j == 2 ? th.z : th.w));
      const unsigned c = W.s_tiny[i0 + j];
      const float a = W.s_alpha[c];
      const float ia = __fdiv_rn(1.f, a);
      const float eb = a * kEm2Half;                                        // b - 1 (s = 2: alpha < kGsTiny < kGsSplitAlpha)
      const float lnb = log1pf(eb);
      const float hb = __fdiv_rn(eb, 1.f + eb);                             // P(p > 1)
      for (unsigned att = 0;; ++att) {
        const uint4 r = k2_philox(make_uint4(t, 0x20000000u | (c << 8) | (att & 255u), W.pid, W.gid), W.key);
        // 1 - U1 = (2^32 - wt - f) / 2^32 with f = 1 - (r.x + 1/2) / 2^32: 24 significant bits at any magnitude
        float delta = fmaf((float)r.x + 0.5f, 2.3283064365386963e-10f, (float)(~wt)) * 2.3283064365386963e-10f;
        delta = fminf(delta, 0.99999994f);
        const float lu2 = lg2_approx(u24(r.y));
        float l2;
        bool ok;
        if (delta < hb) {                                                   // p > 1: x = s - ln V, V = delta / hb
          const float x = fmaf(-kLn2, lg2_approx(__fdiv_rn(delta, hb)), 2.f);
          l2 = lg2_approx(x);
          ok = lu2 <= (a - 1.f) * (l2 - 1.f);
        } else {                                                            // ln p = ln b + ln(1 - delta)
          l2 = fmaf((lnb + log1pf(-delta)) * kLog2e, ia, 1.f);
          ok = lu2 <= -kLog2e * ex2_approx(l2);
        }
        if (ok) {
          fold_log(c * (kLWords * 4u), l2);
          if (PASS != 1) {                                                  // remember the row: it goes back to zero after the class sums
            if ((dirty & 0xffu) < (unsigned)kQCap) { sts_u16(W.a_park + 2u * (dirty & 0xffu), (unsigned)(i0 + j)); ++dirty; }
            else dirty |= 0x100u;
          }
          break;
        }
        // rejected: a fresh attempt, whose top word says whether it can matter at all; if not it is a
        // p <= 1 draw with x < 2^-126 * 2^m: accepted (e^-x = 1 on the U2 grid) and staged as the 0 already there
        wt = r.z;
        if (wt < thr) break;
      }
    }
  }
}

// digamma for x >= 1 (double): recurrence up to x >= 10, then the asymptotic series
__device__ __forceinline__ double digamma_ge1(double x) {
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0)))));
  return r + log(x) - 0.5 / x + t;
}

// SPLIT = false: one work item per pair (many pairs, or injected samples); SPLIT = true: one item per
// (pair, sub-range).  Both instantiations are launched; the one whose regime does not apply returns
// at once (the pair count is only known on the device).
template <bool SPLIT>
__global__ void __launch_bounds__(kK2Threads, kK2MinBlocks)
k2_dirichlet_kernel(const __grid_constant__ Plan p, const float* __restrict__ score_rows,
                    const float* __restrict__ lam_rows, const float* __restrict__ lam_mean,
                    const int* __restrict__ pair_row, const int* __restrict__ pair_obj,
                    const int* __restrict__ pair_off, const long long* __restrict__ image_ids,
                    const float* __restrict__ inj, const long long* __restrict__ inj_off,
                    float* __restrict__ pair_unc, float* __restrict__ pair_avg, int* __restrict__ work_counter,
                    float* __restrict__ part, int* __restrict__ done, const int allow_split,
                    unsigned* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char k2_smem[];
  const int C = p.C, T = p.n_samples;
  int* img_pref = reinterpret_cast<int*>(k2_smem);            // [B+1] exclusive prefix of pair counts
  float* wbase = reinterpret_cast<float*>(img_pref + k2_pref_ints(p.B)) + (size_t)(threadIdx.x >> 5) * k2_warp_floats(C);
  float4* cst4 = reinterpret_cast<float4*>(wbase);            // [C+8]
  unsigned* lbuf = reinterpret_cast<unsigned*>(wbase + 4 * (C + 8));   // [C][kLWords]
  float* s_alpha = wbase + 4 * (C + 8) + C * kLWords;         // [C]
  float* s_avg = s_alpha + C;                                 // [C]
  float* s_part = s_avg + C;                                  // [C] class sums of the current sub-range
  unsigned* s_thr = reinterpret_cast<unsigned*>(s_part + C);  // [C+4 rounded to 4]
  unsigned char* s_dirty = reinterpret_cast<unsigned char*>(s_thr + k2_thr_words(C));   // [32][kQCap] halves
  unsigned char* s_tiny = s_dirty + 32 * kQCap * 2;
  const int lane = threadIdx.x & 31;
  K2Pair W;
  W.a_cst4 = (unsigned)__cvta_generic_to_shared(cst4);
  W.a_thr = (unsigned)__cvta_generic_to_shared(s_thr);
  W.a_park = (unsigned)__cvta_generic_to_shared(s_dirty) + (unsigned)(lane * kQCap * 2);
  W.one = 0x3f800000u | ((unsigned)p.S >> 31);
#ifdef MEHHUA_K2_STAGE_BF16
  W.a_lrow = (unsigned)__cvta_generic_to_shared(lbuf) + 2u * lane;   // my column of row 0
#else
  W.a_lrow = (unsigned)__cvta_generic_to_shared(lbuf) + 4u * lane;
#endif
  W.s_tiny = s_tiny; W.s_alpha = s_alpha;
  W.key = make_uint2((unsigned)(p.seed & 0xffffffffull), (unsigned)(p.seed >> 32));

  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < p.B; ++b) { img_pref[b] = acc; acc += pair_off[b * (p.S + 1) + p.S]; }
    img_pref[p.B] = acc;
  }
  // staged draws of lanes without a sample are multiplied by 0 in the class sums: they must be finite
  for (int i = lane; i < C * kLWords; i += 32) lbuf[i] = 0u;
  __syncthreads();
  const int total = img_pref[p.B];
  const unsigned full = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  const float fT = (float)T;
  // sub-ranges of the sample index: whole rounds of 32, kK2Sub of them
  const int sub_len = (((T + 31) / 32 + kK2Sub - 1) / kK2Sub) * 32;
  // few pairs and no injected samples: one work item per (pair, sub-range) instead of per pair
  const bool split_regime = (allow_split != 0 && inj == nullptr && T > 0 && total <= kK2SplitPairs);
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
    // alpha row; bad = alpha <= 0, denormal-tiny or non-finite (treated as always clamped)
    float amax = 0.f, a0 = 0.f;
    int nbad = 0;
    for (int c = lane; c < C; c += 32) {
      float a = __fmul_rn(srow[c], lamp);
      const bool bad = !(a > 1e-30f) || !(a < 3.0e38f);
      if (bad) { a = 0.f; ++nbad; }
      s_alpha[c] = a;
      s_avg[c] = 0.f;
      amax = fmaxf(amax, a);
      a0 += a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      amax = fmaxf(amax, __shfl_xor_sync(full, amax, o));
      a0 += __shfl_xor_sync(full, a0, o);
      nbad += __shfl_xor_sync(full, nbad, o);
    }
    if (nbad > 0 && lane == 0 && p.act != MEHHUA_ACT_RELU) atomicOr(status, MEHHUA_ST_BAD_ALPHA);   // relu rows have zeros by construction
    __syncwarp();

    const long long ioff = (inj != nullptr && inj_off != nullptr) ? inj_off[b * p.S + s] : -1;
    float ent_acc = 0.f;   // sum over the samples of (-sum_c x ln x)
    if (T == 0) {
      // ---- analytic form (T -> infinity), double precision ----
      const double da0 = (double)a0;
      double h = 0.0, ps = 0.0;
      for (int c = lane; c < C; c += 32) {
        const double a = (double)s_alpha[c];
        if (a > 0.0) {
          const double mc = a / da0;
          h -= mc * log(mc);
          ps += mc * digamma_ge1(a + 1.0);
        }
        if (pair_avg) pair_avg[pq * C + c] = (a > 0.0) ? (float)(a / da0) : 0.f;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        h += __shfl_xor_sync(full, h, o);
        ps += __shfl_xor_sync(full, ps, o);
      }
      if (lane == 0) {
        const double ale = (da0 > 0.0) ? digamma_ge1(da0 + 1.0) - ps : 0.0;
        float* o3 = pair_unc + pq * 3;
        o3[0] = (float)h; o3[1] = (float)ale; o3[2] = (float)(h - ale);
      }
      continue;
    }
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
      // reference exponent of the scaled draws: log2 of the largest alpha when some alpha >= 1 (its
      // draw owns the sample and is within a few octaves of alpha), 0 otherwise
      const float m = amax >= 1.f ? lg2_approx(amax) : 0.f;
      const bool two_pass = !(amax >= 1.f) && a0 < 1.f;
      // a p<=1 draw of a tiny class is non-zero only when log2 p >= -alpha * span
      const float span = fmaxf(kTinySpan + 1.f - m, 0.f);
      // class lists: big (alpha >= 1, Marsaglia-Tsang), GS (kGsTiny <= alpha < 1), tiny (alpha < kGsTiny)
      int ngs = 0, nbig = 0, ntiny = 0;
      int bigpos[8];        // C <= 256: at most 8 classes per lane
#pragma unroll
      for (int i = 0; i < 8; ++i) bigpos[i] = -1;
#pragma unroll
      for (int c0 = 0; c0 < 256; c0 += 32) {
        if (c0 >= C) break;
        const int c = c0 + lane;
        const float a = (c < C) ? s_alpha[c] : 0.f;
        const bool valid = c < C && a > 0.f;
        const bool big = valid && a >= 1.f;
        const bool tiny = valid && a < kGsTiny;
        const bool gs = valid && !big && !tiny;
        const unsigned mb = __ballot_sync(full, big), mg = __ballot_sync(full, gs), mt = __ballot_sync(full, tiny);
        // constants of the class in list order; the big list follows the GS list and its pads, so it
        // is written after the counts are known: park the position in a register
        if (gs) {
          const int pos = ngs + __popc(mg & lt_mask);
          cst4[pos] = make_float4(fmaf(a, kEm1, 1.f), __fdiv_rn(1.f, a), __uint_as_float((unsigned)c * (kLWords * 4u)), a - 1.f);
        }
        if (big) bigpos[c0 >> 5] = nbig + __popc(mb & lt_mask);
        if (tiny) {
          // hit <=> U1 >= 2^(-alpha span) / b, conservatively: P(hit) = 1 - 2^(-alpha span) / b, rounded up
          const int pos = ntiny + __popc(mt & lt_mask);
          const float ph = -expm1f(-(a * span * kLn2 + log1pf(a * kEm2Half)));
          const float phs = fmaf(ph * 4294967296.f, 1.00001f, 2.f);
          s_tiny[pos] = (unsigned char)c;
          s_thr[pos] = (two_pass || !(phs < 4294967040.f)) ? 0u : 0u - (unsigned)phs;
        }
        if (c < C && !gs && !big) {           // rows that are written rarely (tiny) or never (bad) start at zero
          unsigned* rowp = lbuf + (size_t)c * kLWords;
          for (int i = 0; i < kLWords; ++i) rowp[i] = 0u;
        }
        nbig += __popc(mb); ngs += __popc(mg); ntiny += __popc(mt);
      }
      float4* big4 = cst4 + ngs + 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (bigpos[i] >= 0) {
          const int c = i * 32 + lane;
          const float d = s_alpha[c] - (1.f / 3.f);
          big4[bigpos[i]] = make_float4(d, rsqrtf(9.f * d), __uint_as_float((unsigned)c * (kLWords * 4u)), lg2_approx(d) - m);
        }
      }
      if (lane < 4) {       // pad entries read by finished cursors (never accepted) / the last threshold chunk
        cst4[ngs + lane] = make_float4(1.f, __int_as_float(0x7fc00000), 0.f, 0.f);
        if (lane < 2) big4[nbig + lane] = make_float4(1.f, __int_as_float(0x7fc00000), 0.f, 0.f);
        s_thr[ntiny + lane] = 0xffffffffu;
      }
      W.a_big4 = W.a_cst4 + (unsigned)(ngs + 4) * 16u;
      W.ngs = ngs; W.nbig = nbig; W.ntiny = ntiny;
      W.tiny_always = two_pass;
      W.gid = image_ids ? (unsigned)image_ids[b] : (unsigned)b;
      W.pid = (unsigned)row | ((unsigned)obj << 20);
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
        // Normalisation is a log-sum-exp around the reference exponent:
        //   e_c = 2^(l_c - m),  A = sum_c e_c,  x_c = e_c / A,
        //   -sum_c x ln x = ln A - ln2 * (sum_c e_c d_c) / A   with d_c = l_c - m (log2 units)
        // Any m gives the same value; fixing it early lets every accepted draw be folded into A and
        // the entropy sum on the spot, so the draws are stored as e_c and read back only once.
        float asum = 0.f, bs = 0.f, mx = -INFINITY;
        unsigned dirty = 0;    // tiny rows this lane wrote (count, bit 8 = more than its parking slots hold)
        if (!two_pass) {
          k2_draw_sample<0>(W, (unsigned)t, active, m, asum, bs, mx, dirty);
        } else {
          k2_draw_sample<1>(W, (unsigned)t, active, 0.f, asum, bs, mx, dirty);
          const float ml = (mx > -INFINITY) ? mx : 0.f;
          k2_draw_sample<2>(W, (unsigned)t, active, ml, asum, bs, mx, dirty);
        }
        float inv_a = 0.f;
        if (active && asum > 0.f) {
          inv_a = __fdividef(1.f, asum);
          ent_sub += logf(asum) - kLn2 * bs * inv_a;
        }
        __syncwarp();
        // class sums over the 32 samples, transposed: lane = class
#ifndef MEHHUA_K2_STAGE_BF16
        for (int c0 = 0; c0 < C; c0 += 96) {      // three class groups share the broadcasts of 1/A
          const int c = c0 + lane, rem = C - c0;
          const float* r0p = reinterpret_cast<const float*>(lbuf) + (size_t)min(c, C - 1) * kLWords;
          const float* r1p = reinterpret_cast<const float*>(lbuf) + (size_t)min(c + 32, C - 1) * kLWords;
          const float* r2p = reinterpret_cast<const float*>(lbuf) + (size_t)min(c + 64, C - 1) * kLWords;
          float a0s = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 8
          for (int w = 0; w < 32; ++w) {
            const float iv = __shfl_sync(full, inv_a, w);
            a0s = fmaf(r0p[w], iv, a0s);
            if (rem > 32) a1 = fmaf(r1p[w], iv, a1);
            if (rem > 64) a2 = fmaf(r2p[w], iv, a2);
          }
          if (c < C) s_part[c] += a0s;
          if (c + 32 < C) s_part[c + 32] += a1;
          if (c + 64 < C) s_part[c + 64] += a2;
        }
#else
        // walk the samples two at a time (one word = the bf16 copies of samples 2w and 2w+1; bf16 -> fp32 is a shift / a mask)
        for (int c0 = 0; c0 < C; c0 += 128) {
          float a0s = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          const int c = c0 + lane;
          const unsigned* r0p = lbuf + (size_t)min(c, C - 1) * kLWords;
          const unsigned* r1p = lbuf + (size_t)min(c + 32, C - 1) * kLWords;
          const unsigned* r2p = lbuf + (size_t)min(c + 64, C - 1) * kLWords;
          const unsigned* r3p = lbuf + (size_t)min(c + 96, C - 1) * kLWords;
          const int rem = C - c0;
#pragma unroll 4
          for (int w = 0; w < 16; ++w) {
            const float iv0 = __shfl_sync(full, inv_a, 2 * w), iv1 = __shfl_sync(full, inv_a, 2 * w + 1);
            unsigned u = r0p[w];
            a0s = fmaf(__uint_as_float(u << 16), iv0, a0s);
            a0s = fmaf(__uint_as_float(u & 0xffff0000u), iv1, a0s);
            if (rem > 32) { u = r1p[w]; a1 = fmaf(__uint_as_float(u << 16), iv0, a1); a1 = fmaf(__uint_as_float(u & 0xffff0000u), iv1, a1); }
            if (rem > 64) { u = r2p[w]; a2 = fmaf(__uint_as_float(u << 16), iv0, a2); a2 = fmaf(__uint_as_float(u & 0xffff0000u), iv1, a2); }
            if (rem > 96) { u = r3p[w]; a3 = fmaf(__uint_as_float(u << 16), iv0, a3); a3 = fmaf(__uint_as_float(u & 0xffff0000u), iv1, a3); }
          }
          if (c < C) s_part[c] += a0s;
          if (c + 32 < C) s_part[c + 32] += a1;
          if (c + 64 < C) s_part[c + 64] += a2;
          if (c + 96 < C) s_part[c + 96] += a3;
        }
#endif
        __syncwarp();
        // a tiny-class slot that took a draw this round goes back to zero (the other rounds skip it).  In
        // the two-pass form every tiny slot of an active lane is rewritten every round instead.
        if (!two_pass && __any_sync(full, dirty != 0u)) {
          if (__any_sync(full, (dirty & 0x100u) != 0u)) {
            for (int i = lane; i < ntiny * kLWords; i += 32) lbuf[(size_t)s_tiny[i / kLWords] * kLWords + (i % kLWords)] = 0u;
          } else {
            for (unsigned i = 0; i < (dirty & 0xffu); ++i)
              stage_store(W.a_lrow + (unsigned)s_tiny[lds_u16(W.a_park + 2u * i)] * (kLWords * 4u), 0.f);
          }
          __syncwarp();
        }
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
      const float mean = __fdiv_rn(s_avg[c], fT);
      if (pair_avg) pair_avg[pq * C + c] = mean;
      const float avg = fmaxf(mean, kFltMin);
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

// debug / known-answer entry: one Philox4x32 block with 10 rounds (out[0..3]) and with 7 (out[4..7])
__global__ void philox_kat_kernel(uint4 ctr, uint2 key, unsigned* out) {
  const uint4 r = philox4x32<10>(ctr, key);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
  const uint4 q = philox4x32<7>(ctr, key);
  out[4] = q.x; out[5] = q.y; out[6] = q.z; out[7] = q.w;
}

}  // namespace mehhua
