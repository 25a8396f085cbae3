// K2: Dirichlet-sampled epistemic uncertainty per (box, object) pair.
//
// Reference semantics (mmdet/models/dense_heads/Lambda_L2.py:513-525):
//   lambda' = mean(lambda_pairs) / (lambda + 1e-7) * 25;  alpha = score_row * lambda'
//   x_t ~ Dirichlet(alpha), t = 1..T (T = 500)
//   avg = mean_t x_t;  total = -sum_c avg ln avg;  ale = mean_t(-sum_c x ln x);  epi = total - ale
// The sampler behind torch.distributions.Dirichlet is ATen's _sample_dirichlet (third party): gamma
// draws normalised by their sum and clamped to [FLT_MIN, 1 - 2^-24].  This kernel is written from
// the published algorithms; it is not bit-compatible with torch's stream.  Per class:
//   alpha >= 1        : Marsaglia & Tsang (2000), normal by Box-Muller (24-bit uniforms)
//   alpha_t <= alpha<1: Ahrens & Dieter (1974) algorithm GS, straight-line; U1 on a 2^-23 grid
//                       (midpoints), U2 on a 2^-19 grid (midpoints).  Enumerating those grids gives
//                       |E[g]/alpha - 1| <= 1e-5 for alpha >= 1e-3 (tools/gs_grid_bias.py,
//                       profiles/r2_gs_grid_bias.txt); the acceptance probability is quantised at 2^-20.
//   alpha <  alpha_t  : the boost identity g = G' * exp(-E / alpha), G' ~ Gamma(1 + alpha) (Marsaglia-
//                       Tsang), E = -ln(1 - V) with V a 64-bit uniform.  A draw can only matter when
//                       E < alpha * ln2 * (span of representable exponents), i.e. when the top word
//                       of V is below a per-class threshold: one integer compare decides almost every
//                       draw of a tiny-alpha class ("surely flushed to zero by the fp32 arithmetic
//                       below" - the skipped draws are exactly the ones whose scaled value would have
//                       been 0), the few others take the full formula.  Resolution of V near 0 is
//                       2^-64, so the form is exact for every alpha the path can produce.
//   alpha_t is where the skip probability reaches 3/4: alpha_t = 0.25 / (ln2 * (135 - m)) ~ 2.7e-3.
// Everything is kept in LOG space (log2 g), so tiny draws neither underflow nor need special casing;
// the reference's FLT_MIN clamp only changes terms below 1e-36.
// Normalisation: e_c = g_c / 2^m with a reference exponent m fixed per pair BEFORE the draws (log2 of
// the largest alpha when some alpha >= 1, else 0).  Rows whose alphas are all < 1 and sum to < 1
// (lambda' << 1) can have every draw of a sample far below 2^-126; they take a two-pass form: the
// counter-based generator replays the sample, pass 1 finds its largest log2 draw, pass 2 normalises
// around it.
// Randomness: counter-based Philox4x32-10 keyed by the seed, counter = (sample, call index, pair
// identity (row, object), global image id) - independent of batch composition and world size.
// Layout: one warp owns a pair; lane = sample (32 at a time).  The scaled draws e_c of the warp's 32
// samples sit in shared memory as lbuf[class][lane] in bfloat16 (round-to-nearest; row stride 17 words:
// conflict-free both for the lane-private stores of the draw loop and for the transposed class-sum
// pass).  Only the class means are formed from the bf16 copies (relative rounding error 2^-9 per term,
// unbiased, averaged over T samples); the per-sample sum A and the entropy terms stay in fp32
// registers.  -DMEHHUA_K2_STAGE_FP32 builds the same kernel with fp32 staging (an A/B build for the
// tests: twice the shared memory, two blocks per SM instead of three).  Samples never touch global memory.
// n_samples == 0 selects the ANALYTIC form (T -> infinity): total = H(alpha/alpha0),
// ale = psi(alpha0 + 1) - sum_c (alpha_c/alpha0) psi(alpha_c + 1), in double precision - a
// deterministic mode for pool-level set-identity tests against the oracle's closed forms.
#pragma once
#include "common.cuh"

#ifndef MEHHUA_PHILOX_ROUNDS
#define MEHHUA_PHILOX_ROUNDS 10
#endif

namespace mehhua {

constexpr int kK2Threads = 256;
constexpr int kK2Warps = kK2Threads / 32;
#ifdef MEHHUA_K2_STAGE_FP32
constexpr int kLWords = 33;                           // fp32 staging: 32 samples + 1 pad word per class row
constexpr int kK2MinBlocks = 2;
#else
constexpr int kLWords = 17;                           // bf16 staging: 34 halves per class row
constexpr int kK2MinBlocks = 3;
#endif
constexpr float kFltMin = 1.17549435e-38f;
constexpr float kInvE = 0.36787944117144233f;
constexpr float kTinySpan = 135.f;                   // exponents between the largest possible boost factor (2^7) and the flush point (2^-126), + margin
constexpr float kTinyEps = 0.25f;                    // a class is "tiny" when alpha * ln2 * span <= this (skip probability >= e^-0.25)
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
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) { return philox4x32<10>(c, k); }
__device__ __forceinline__ uint4 k2_philox(uint4 c, uint2 k) { return philox4x32<MEHHUA_PHILOX_ROUNDS>(c, k); }

// uniform in (0,1) with 24-bit resolution, never 0 or 1
__device__ __forceinline__ float u24(unsigned w) { return ((float)(w >> 8) + 0.5f) * 5.9604644775390625e-08f; }

// shared-memory accessors on 32-bit shared-window addresses (keeps address arithmetic in one IMAD)
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
// one staged draw: lbuf[class][lane]
__device__ __forceinline__ void stage_store(unsigned a, float v) {
#ifdef MEHHUA_K2_STAGE_FP32
  asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v));
#else
  asm volatile("{ .reg .b16 h; cvt.rn.bf16.f32 h, %1; st.shared.b16 [%0], h; }" :: "r"(a), "f"(v));   // round-to-nearest-even
#endif
}

// per-warp shared memory, in floats:
//   cst4[C+4] (float4: b, 1/alpha, alpha-1, b/alpha; in GS-list order) | lbuf[C*kLWords] | alpha[C] | avg[C] | part[C] |
//   thr[C+4 rounded to 4] (tiny-list order) | lists 3 x (C+4) bytes (GS, big, tiny).
// The GS list and its constants carry 4 pad entries that can never be accepted, the threshold list is
// padded with zeros (never hit), so cursors / chunks that run past the end need no clamp.
__host__ __device__ inline size_t k2_thr_words(int C) { return ((size_t)C + 4 + 3) & ~(size_t)3; }
__host__ __device__ inline size_t k2_warp_floats(int C) {
  const size_t f = 4 * ((size_t)C + 4) + (size_t)C * kLWords + 3 * (size_t)C + k2_thr_words(C) + (3 * ((size_t)C + 4) + 3) / 4;
  return (f + 3) & ~(size_t)3;
}
// + the per-image pair-count prefix [B+1], padded to 16 ints
__host__ __device__ inline int k2_pref_ints(int B) { return (B + 1 + 15) & ~15; }
__host__ __device__ inline size_t k2_smem_bytes(int C, int B) {
  return kK2Warps * k2_warp_floats(C) * sizeof(float) + k2_pref_ints(B) * sizeof(int);
}

// What a warp knows about the pair it is working on.
struct K2Pair {
  unsigned a_cst4, a_gs, a_lrow, a_thr;   // shared-window addresses: GS constants, GS class list, my lbuf column, tiny thresholds
  const unsigned char* s_big;
  const unsigned char* s_tiny;
  const float* s_alpha;
  int ngs, nbig, ntiny;
  bool tiny_always;                       // two-pass form: every tiny-class draw is evaluated
  unsigned pid, gid;
  uint2 key;
};

// log2 of a Gamma(shape) draw, shape >= 1 (Marsaglia-Tsang; normal by Box-Muller).  One Philox block
// per attempt, counter word 1 = tag | attempt.  `spare` receives the unused fourth word of the first
// block (the boost form takes the low half of its 64-bit uniform from it).
__device__ __forceinline__ float mt_log2_gamma(const float shape, const unsigned t, const unsigned tag, const K2Pair& W,
                                               unsigned& spare) {
  const float d = shape - (1.f / 3.f);
  const float cc = rsqrtf(9.f * d);
  for (unsigned att = 0;; ++att) {
    const uint4 w = k2_philox(make_uint4(t, tag | (att & 255u), W.pid, W.gid), W.key);
    if (att == 0) spare = w.w;
    const float r = sqrtf(-2.f * kLn2 * lg2_approx(u24(w.x)));
    const float x = r * __cosf((float)w.y * 1.4629180792671596e-09f);   // 2*pi / 2^32
    const float v1 = fmaf(cc, x, 1.f);
    if (v1 > 0.f) {
      const float lv2 = 3.f * lg2_approx(v1);
      const float v = v1 * v1 * v1;
      if (kLn2 * lg2_approx(u24(w.z)) < fmaf(d, fmaf(kLn2, lv2, 1.f - v), 0.5f * x * x)) return lg2_approx(d) + lv2;
    }
  }
}

// One Ahrens-Dieter GS attempt for the class at position i of the GS list, as straight-line code (no
// branches, so the attempts of a lane's cursors interleave in the pipelines).  Returns whether the
// draw is accepted, its class, l2 = log2(draw) and x = the draw itself (0 when it is below 2^-126).
//   p = b*U1;  p <= 1: x = p^(1/alpha), accept iff U2 <= exp(-x)
//              p >  1: x = -ln((b-p)/alpha) >= 1, accept iff U2 <= x^(alpha-1)
// both tests are done as log2(U2) <= rhs.  f0, f1 in [1,2): U1 = f0 - 1 + 2^-24 (23-bit grid midpoints),
// U2 = f1 - 1 + 2^-20 (19-bit grid midpoints); both subtractions are exact.  Pad entries carry
// 1/alpha = NaN: every comparison with their rhs is false.
__device__ __forceinline__ bool gs_attempt(const int i, const float f0, const float f1, const K2Pair& W,
                                           unsigned& c, float& l2, float& x) {
  const unsigned ii = (unsigned)i;
  c = lds_u8(W.a_gs + ii);
  const float4 k = lds_v4(W.a_cst4 + ii * 16u);         // b, 1/alpha, alpha-1, b/alpha (list order)
  const float pp = (f0 - 0.99999994039535522f) * k.x;   // b * U1
  const bool lo = pp <= 1.f;
  const float q = lo ? pp : fmaf(-pp, k.y, k.w);        // second branch: (b - p) / alpha
  const float lq = lg2_approx(q);
  const float l2a = lq * k.y;                           // log2 x, first branch
  const float xa = ex2_approx(l2a);                     // x
  const float xb = -kLn2 * lq;                          // x, second branch
  const float l2b = lg2_approx(xb);
  x = lo ? xa : xb;
  l2 = lo ? l2a : l2b;
  const float rhs = lo ? -kLog2e * xa : k.z * l2b;      // log2 exp(-x)  /  log2 x^(alpha-1)
  return lg2_approx(f1 - 0.99999904632568359f) <= rhs;
}

// The draws of ONE sample (this lane's sample t) for every class of the pair.
//   PASS 0: normalise around the pair's fixed exponent m (sc = 2^-m): stage e_c, accumulate A = sum e and
//           bs = sum e * log2 e.
//   PASS 1: only the largest log2 draw of the sample (mx).
//   PASS 2: as PASS 0 around a per-lane exponent m (no 2^-m factor: it may overflow).
// The three forms consume identical Philox counters, so PASS 2 replays PASS 1 draw for draw.
template <int PASS>
__device__ __forceinline__ void k2_draw_sample(const K2Pair& W, const unsigned t, const bool active, const float m,
                                               const float sc, float& asum, float& bs, float& mx, bool& tiny_hit) {
  const unsigned full = 0xffffffffu;
  // a draw known by its log2 only (Marsaglia-Tsang classes, boost form)
  auto fold_log = [&](const unsigned c, const float l2) {
    if (PASS == 1) { mx = fmaxf(mx, l2); return; }
    const float d = l2 - m;
    const float e = ex2_approx(d);          // exactly 0 below 2^-126 (ftz)
    stage_store(W.a_lrow + c * (kLWords * 4u), e);
    asum += e;
    bs = fmaf(e, d, bs);
  };
  if (active) {
    for (int i = 0; i < W.nbig; ++i) {
      const unsigned c = W.s_big[i];
      unsigned spare;
      fold_log(c, mt_log2_gamma(W.s_alpha[c], t, 0x80000000u | (c << 8), W, spare));
    }
  }
  // Ahrens-Dieter GS classes.  Flattened rejection loop with three cursors fed from a shared work
  // list: each iteration one Philox block (4 words) feeds one attempt per cursor - word j gives U1
  // (23 bits) and the top of U2 (its 9 spare bits), a third of the fourth word the rest of U2 - and a
  // cursor that accepts takes the next undrawn class, so a lane never idles while a neighbour
  // retries and the lanes of a warp finish within a few attempts of each other.
  if (W.ngs > 0) {
    const int done_at = W.ngs + 3;          // a lane is done when it has taken ngs + 3 list positions
    int ia = active ? 0 : W.ngs, ib = active ? 1 : W.ngs, ic = active ? 2 : W.ngs;
    int nxt = active ? 3 : done_at;
    unsigned kcall = 0;
    while (__any_sync(full, nxt < done_at)) {
      const uint4 w = k2_philox(make_uint4(t, kcall++, W.pid, W.gid), W.key);
      const float f0a = __uint_as_float(0x3f800000u | (w.x >> 9));
      const float f0b = __uint_as_float(0x3f800000u | (w.y >> 9));
      const float f0c = __uint_as_float(0x3f800000u | (w.z >> 9));
      // U2 mantissa = [9 low bits of the word | 10 or 11 bits of w.w], the rest of the mantissa zero
      const float f1a = __uint_as_float(0x3f800000u | (__funnelshift_l(w.w, w.x, 14) & 0x007ffff0u));
      const float f1b = __uint_as_float(0x3f800000u | (__funnelshift_l(w.w << 10, w.y, 14) & 0x007ffff8u));
      const float f1c = __uint_as_float(0x3f800000u | (__funnelshift_l(w.w << 21, w.z, 14) & 0x007ffff8u));
      unsigned ca, cb, cc;
      float la, lb, lc, xa, xb, xc;
      const bool oka = gs_attempt(ia, f0a, f1a, W, ca, la, xa);
      const bool okb = gs_attempt(ib, f0b, f1b, W, cb, lb, xb);
      const bool okc = gs_attempt(ic, f0c, f1c, W, cc, lc, xc);
      if (PASS == 1) {
        if (oka) { mx = fmaxf(mx, la); ia = nxt; ++nxt; }
        if (okb) { mx = fmaxf(mx, lb); ib = nxt; ++nxt; }
        if (okc) { mx = fmaxf(mx, lc); ic = nxt; ++nxt; }
      } else {
        // l is finite (U1 > 0, alpha >= alpha_t), so d = l - m needs no clamp: e = 0 below 2^-126 and 0 * d = 0
        const float da = la - m, db = lb - m, dc = lc - m;
        const float ea = (PASS == 0) ? xa * sc : ex2_approx(da);
        const float eb = (PASS == 0) ? xb * sc : ex2_approx(db);
        const float ec = (PASS == 0) ? xc * sc : ex2_approx(dc);
        if (oka) { stage_store(W.a_lrow + ca * (kLWords * 4u), ea); asum += ea; bs = fmaf(ea, da, bs); ia = nxt; ++nxt; }
        if (okb) { stage_store(W.a_lrow + cb * (kLWords * 4u), eb); asum += eb; bs = fmaf(eb, db, bs); ib = nxt; ++nxt; }
        if (okc) { stage_store(W.a_lrow + cc * (kLWords * 4u), ec); asum += ec; bs = fmaf(ec, dc, bs); ic = nxt; ++nxt; }
      }
    }
  }
  // tiny-alpha classes: four per Philox block, one compare each; a hit takes the boost formula
  for (int i0 = 0; i0 < W.ntiny; i0 += 4) {
    const uint4 w = k2_philox(make_uint4(t, 0x40000000u | (unsigned)(i0 >> 2), W.pid, W.gid), W.key);
    const uint4 th = lds_u4(W.a_thr + (unsigned)i0 * 4u);
    bool h0 = w.x < th.x, h1 = w.y < th.y, h2 = w.z < th.z, h3 = w.w < th.w;
    if (W.tiny_always) { h0 = true; h1 = i0 + 1 < W.ntiny; h2 = i0 + 2 < W.ntiny; h3 = i0 + 3 < W.ntiny; }
    unsigned hm = active ? ((unsigned)h0 | ((unsigned)h1 << 1) | ((unsigned)h2 << 2) | ((unsigned)h3 << 3)) : 0u;
#pragma unroll 1
    while (hm != 0u) {            // rare: a lane's hits, one at a time
      const int j = __ffs(hm) - 1;
      hm &= hm - 1u;
      const unsigned w0 = j == 0 ? w.x : (j == 1 ? w.y : (j == 2 ? w.z : w.w));
      const unsigned c = W.s_tiny[i0 + j];
      const float a = W.s_alpha[c];
      unsigned spare;
      const float lg = mt_log2_gamma(1.f + a, t, 0x20000000u | (c << 8), W, spare);
      // V = (w0 + (spare + 1/2) / 2^32) / 2^32: 24 significant bits at any magnitude; E = -ln(1 - V)
      float v = fmaf((float)spare + 0.5f, 2.3283064365386963e-10f, (float)w0) * 2.3283064365386963e-10f;
      v = fminf(v, 0.99999994f);
      const float e1 = -log1pf(-v);
      fold_log(c, lg - __fdiv_rn(e1 * kLog2e, a));
      tiny_hit = true;
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
  float4* cst4 = reinterpret_cast<float4*>(wbase);            // [C+4]
  unsigned* lbuf = reinterpret_cast<unsigned*>(wbase + 4 * (C + 4));   // [C][kLWords]
  float* s_alpha = wbase + 4 * (C + 4) + C * kLWords;         // [C]
  float* s_avg = s_alpha + C;                                 // [C]
  float* s_part = s_avg + C;                                  // [C] class sums of the current sub-range
  unsigned* s_thr = reinterpret_cast<unsigned*>(s_part + C);  // [C+4 rounded to 4]
  unsigned char* s_gs = reinterpret_cast<unsigned char*>(s_thr + k2_thr_words(C));
  unsigned char* s_big = s_gs + C + 4;
  unsigned char* s_tiny = s_big + C + 4;
  const int lane = threadIdx.x & 31;
  K2Pair W;
  W.a_cst4 = (unsigned)__cvta_generic_to_shared(cst4);
  W.a_gs = (unsigned)__cvta_generic_to_shared(s_gs);
  W.a_thr = (unsigned)__cvta_generic_to_shared(s_thr);
#ifdef MEHHUA_K2_STAGE_FP32
  W.a_lrow = (unsigned)__cvta_generic_to_shared(lbuf) + 4u * lane;   // my column of row 0
#else
  W.a_lrow = (unsigned)__cvta_generic_to_shared(lbuf) + 2u * lane;
#endif
  W.s_big = s_big; W.s_tiny = s_tiny; W.s_alpha = s_alpha;
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
    if (nbad > 0 && lane == 0) atomicOr(status, MEHHUA_ST_BAD_ALPHA);
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
      const float sc = ex2_approx(-m);
      const bool two_pass = !(amax >= 1.f) && a0 < 1.f;
      const float span = fmaxf(kTinySpan - m, 0.f);
      // class lists: big (alpha >= 1, Marsaglia-Tsang), GS (alpha_t <= alpha < 1), tiny (alpha < alpha_t, boost form)
      int ngs = 0, nbig = 0, ntiny = 0;
      for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        const float a = (c < C) ? s_alpha[c] : 0.f;
        const bool valid = c < C && a > 0.f;
        const float eps = a * kLn2 * span;
        const bool big = valid && a >= 1.f;
        const bool tiny = valid && !big && eps <= kTinyEps;
        const bool gs = valid && !big && !tiny;
        const unsigned mb = __ballot_sync(full, big), mg = __ballot_sync(full, gs), mt = __ballot_sync(full, tiny);
        if (big) s_big[nbig + __popc(mb & lt_mask)] = (unsigned char)c;
        if (gs) {         // list entry = class id; the GS constants sit at the same list position
          const int pos = ngs + __popc(mg & lt_mask);
          const float bb = fmaf(a, kInvE, 1.f);
          const float ia = __fdiv_rn(1.f, a);
          s_gs[pos] = (unsigned char)c;
          cst4[pos] = make_float4(bb, ia, a - 1.f, bb * ia);
        }
        if (tiny) {       // skip unless the top word of V is below P(E < eps) = 1 - exp(-eps), rounded up
          const int pos = ntiny + __popc(mt & lt_mask);
          s_tiny[pos] = (unsigned char)c;
          s_thr[pos] = (unsigned)fminf(ceilf(-expm1f(-eps) * 4294967296.f * 1.000001f), 4294967040.f);
        }
        if (c < C && !gs && !big) {           // rows that are written rarely (tiny) or never (bad) start at zero
          unsigned* rowp = lbuf + (size_t)c * kLWords;
          for (int i = 0; i < kLWords; ++i) rowp[i] = 0u;
        }
        nbig += __popc(mb); ngs += __popc(mg); ntiny += __popc(mt);
      }
      if (lane < 4) {       // pad entries read by finished cursors / the last threshold chunk
        s_gs[ngs + lane] = 0;
        cst4[ngs + lane] = make_float4(1.f, __int_as_float(0x7fc00000), 0.f, 1.f);
        s_thr[ntiny + lane] = 0u;
      }
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
        bool tiny_hit = false;
        if (!two_pass) {
          k2_draw_sample<0>(W, (unsigned)t, active, m, sc, asum, bs, mx, tiny_hit);
        } else {
          k2_draw_sample<1>(W, (unsigned)t, active, 0.f, 1.f, asum, bs, mx, tiny_hit);
          const float ml = (mx > -INFINITY) ? mx : 0.f;
          k2_draw_sample<2>(W, (unsigned)t, active, ml, 1.f, asum, bs, mx, tiny_hit);
        }
        float inv_a = 0.f;
        if (active && asum > 0.f) {
          inv_a = __fdividef(1.f, asum);
          ent_sub += logf(asum) - kLn2 * bs * inv_a;
        }
        __syncwarp();
        // class sums over the 32 samples, transposed: lane = class
#ifdef MEHHUA_K2_STAGE_FP32
        for (int c0 = 0; c0 < C; c0 += 32) {
          const int c = c0 + lane;
          const float* rp = reinterpret_cast<const float*>(lbuf) + (size_t)min(c, C - 1) * kLWords;
          float a = 0.f;
#pragma unroll 8
          for (int w = 0; w < 32; ++w) a = fmaf(rp[w], __shfl_sync(full, inv_a, w), a);
          if (c < C) s_part[c] += a;
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
        // a tiny-class row that took a draw this round goes back to zero (its other rounds skip it)
        if (!two_pass && __any_sync(full, tiny_hit)) {
          for (int i = lane; i < ntiny * kLWords; i += 32) lbuf[(size_t)s_tiny[i / kLWords] * kLWords + (i % kLWords)] = 0u;
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

// debug / known-answer entry: one Philox4x32-10 block
__global__ void philox_kat_kernel(uint4 ctr, uint2 key, unsigned* out) {
  const uint4 r = philox4x32_10(ctr, key);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

}  // namespace mehhua
