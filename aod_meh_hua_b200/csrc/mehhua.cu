// C ABI of the mehhua scoring path (see include/mehhua.h).  Host side: argument validation, the
// batch Plan, workspace carving and kernel launches.  No torch types, no hidden allocation except
// inside the explicit host-buffer context.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"
#include "k1_alpha_topk.cuh"
#include "k2_dirichlet.cuh"
#include "k3_hua.cuh"
#include "k4_pool_topk.cuh"
#include "ka_entropy_all.cuh"
#include "km_mutual_info.cuh"

using namespace mehhua;

// a level's kept rows come from the coalesced rescan instead of the gather when RATIO * k >= n
#ifndef MEHHUA_RESCAN_RATIO
#define MEHHUA_RESCAN_RATIO 2
#endif
// ... and, for levels that are not captured, when RESCAN_WIDE * k >= n
#ifndef MEHHUA_RESCAN_WIDE
#define MEHHUA_RESCAN_WIDE 2
#endif

namespace {

thread_local char g_err[256] = "";
std::atomic<unsigned long long> g_launches{0};

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return MEHHUA_E_CUDA;
}
int arg_fail(const char* what) {
  snprintf(g_err, sizeof(g_err), "invalid argument: %s", what);
  return MEHHUA_E_ARG;
}
#define CU(call)                                            \
  do {                                                      \
    cudaError_t e_ = (call);                                \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);     \
  } while (0)
#define LAUNCHED(name)                                      \
  do {                                                      \
    g_launches.fetch_add(1, std::memory_order_relaxed);     \
    cudaError_t e_ = cudaGetLastError();                    \
    if (e_ != cudaSuccess) return cuda_fail(e_, name);      \
  } while (0)

size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// Optional per-stage CUDA-event timing of mehhua_score_batch (bench.py's live roofline numbers).
constexpr int kStages = 7;   // K1a, K1b, K1c, K3a, K3b, K2, K3c
struct StageTimer {
  bool on = false;
  std::vector<cudaEvent_t> ev;   // (kStages + 1) events per recorded call
  size_t calls = 0, max_calls = 0;
} g_timer;

void timer_mark(cudaStream_t st, int slot) {
  if (!g_timer.on || g_timer.calls >= g_timer.max_calls) return;
  cudaEventRecord(g_timer.ev[g_timer.calls * (kStages + 1) + slot], st);
}

bool k1_has_typed(int head, int c_out) {      // the <C, HEAD> instantiations of launch_k1
  return head == MEHHUA_HEAD_RETINA ? (c_out == 20 || c_out == 80) : (c_out == 21 || c_out == 81);
}
bool no_capture() {
  static const bool v = [] { const char* e = getenv("MEHHUA_NO_CAPTURE"); return e && e[0] == '1'; }();
  return v;
}

bool no_parked_kernel() {      // MEHHUA_NO_PARKED_KERNEL=1: parked rows are read by the thread-per-row gather kernel (A/B, tests)
  static const bool v = [] { const char* e = getenv("MEHHUA_NO_PARKED_KERNEL"); return e && e[0] == '1'; }();
  return v;
}

bool parked_bulk() {           // MEHHUA_PARKED_BULK=1: the parked rows are staged in shared memory by bulk-async copies (A/B, tests)
  static const bool v = [] { const char* e = getenv("MEHHUA_PARKED_BULK"); return e && e[0] == '1'; }();
  return v;
}

int build_plan(const mehhua_config_t* cfg, const mehhua_level_t* lv, int B, bool need_ptrs, Plan* out) {
  if (!cfg || !lv) return arg_fail("null config / levels");
  if (cfg->num_levels < 1 || cfg->num_levels > kMaxLevels) return arg_fail("num_levels must be 1..8");
  if (cfg->head != MEHHUA_HEAD_RETINA && cfg->head != MEHHUA_HEAD_SSD) return arg_fail("head");
  if (cfg->c_out < 2 || cfg->c_out > 1024) return arg_fail("c_out must be 2..1024");
  if (cfg->max_per_img < 1 || cfg->max_per_img > MEHHUA_MAX_DETS) return arg_fail("max_per_img must be 1..256");
  if (cfg->nms_pre > MEHHUA_MAX_NMS_PRE) return arg_fail("nms_pre must be <= 4096");
  if (B < 1 || B > 1024) return arg_fail("batch must be 1..1024");
  if (cfg->pair_cap < 1) return arg_fail("pair_cap must be positive");
  if (cfg->n_samples < 0) return arg_fail("n_samples must be >= 0 (0 = analytic form)");
  if (cfg->mode != MEHHUA_MODE_NMS && cfg->mode != MEHHUA_MODE_ALL) return arg_fail("mode");
  for (int a : {cfg->agg_object, cfg->agg_scale})
    if (a < MEHHUA_AGG_SUM || a > MEHHUA_AGG_MAX) return arg_fail("aggregation op");
  if (cfg->agg_class < MEHHUA_AGG_SUM || cfg->agg_class > MEHHUA_AGG_POOL ||
      (cfg->agg_class == MEHHUA_AGG_POOL && cfg->mode != MEHHUA_MODE_ALL))
    return arg_fail("class aggregation op (MEHHUA_AGG_POOL needs MEHHUA_MODE_ALL)");
  if (cfg->activation != MEHHUA_ACT_SOFTMAX &&
      !(cfg->activation == MEHHUA_ACT_RELU_PLUS_ONE && cfg->mode == MEHHUA_MODE_NMS && cfg->head == MEHHUA_HEAD_RETINA) &&
      !(cfg->activation == MEHHUA_ACT_RELU && cfg->mode == MEHHUA_MODE_ALL && cfg->head == MEHHUA_HEAD_RETINA))
    return arg_fail("activation: relu_plus_one needs MODE_NMS + Retina head, relu needs MODE_ALL + Retina head");
  Plan p;
  memset(&p, 0, sizeof(p));
  p.S = cfg->num_levels; p.B = B; p.C = cfg->c_out; p.head = cfg->head;
  p.num_fg = cfg->head == MEHHUA_HEAD_SSD ? cfg->c_out - 1 : cfg->c_out;
  long long n_off = 0, k_off = 0, tile0 = 0, rtile0 = 0;
  for (int s = 0; s < p.S; ++s) {
    LevelDev& L = p.lv[s];
    if (lv[s].H < 1 || lv[s].W < 1 || lv[s].A < 1) return arg_fail("level geometry");
    if (need_ptrs && (!lv[s].logits || !lv[s].lambda)) return arg_fail("null level pointer");
    if (need_ptrs && cfg->mode == MEHHUA_MODE_NMS && (!lv[s].deltas || !lv[s].anchors))
      return arg_fail("null level pointer");
    L.logits = lv[s].logits; L.deltas = lv[s].deltas; L.lam = lv[s].lambda; L.anchors = lv[s].anchors;
    L.H = lv[s].H; L.W = lv[s].W; L.A = lv[s].A; L.HW = L.H * L.W;
    const long long n = (long long)L.HW * L.A;
    if (n > (1ll << 30)) return arg_fail("level too large");
    L.n = (int)n;
    L.topk = (cfg->nms_pre > 0 && cfg->nms_pre < L.n) ? 1 : 0;
    L.k = L.topk ? cfg->nms_pre : L.n;
    L.n_off = (int)n_off; L.k_off = (int)k_off;
    const int tile = cfg->mode == MEHHUA_MODE_ALL ? kKaThreads : kK1aThreads;      // the streaming kernel of the mode
    L.tpp = (L.HW + tile - 1) / tile;
    L.tile0 = (int)tile0;
    L.rescan = 0;
    L.cap = -1;
    n_off += n; k_off += L.k; tile0 += (long long)L.tpp * L.A;
  }
  // Where the kept rows of a level come from (k1_alpha_topk.cuh), decided per level:
  //   rescan  - no top-k, or at least half of the priors are kept: a second coalesced pass;
  //   capture - sparse top-k levels of the class counts with a register-resident instantiation park their rows while
  //             K1a streams them, when few of the level's priors are kept (n >= kCapMinRatio * k: few warps pay for
  //             parking) or when the level is a small part of the image (16 n <= N: whatever parking costs there is
  //             small next to a second pass over it).  MEHHUA_NO_CAPTURE=1 in the environment turns capture off;
  //   rescan  - what is left, when at least 1 / MEHHUA_RESCAN_WIDE of the priors are kept: the strided gather touches
  //             one 32-byte sector (64-byte DRAM burst) per 4-byte logit, i.e. as many bytes as the whole level;
  //   gather  - the rest.
  const bool may_capture = k1_has_typed(cfg->head, cfg->c_out) && cfg->mode == MEHHUA_MODE_NMS && !no_capture();
  for (int s = 0; s < p.S; ++s) {
    LevelDev& L = p.lv[s];
    const long long n = L.n;
    if (!L.topk || (long long)MEHHUA_RESCAN_RATIO * L.k >= n) L.rescan = 1;
    else if (may_capture && (n >= (long long)kCapMinRatio * L.k || 16 * n <= n_off)) L.cap = p.n_cap_levels++;
    else if ((long long)MEHHUA_RESCAN_WIDE * L.k >= n) L.rescan = 1;
    L.rtile0 = (int)rtile0;
    if (L.rescan) rtile0 += (n + kRescanThreads - 1) / kRescanThreads;      // tiles over the flattened (anchor, position) axis
  }
  if (n_off > (1ll << 30) || k_off > (1 << 20) || tile0 * B > 0x7fffffffll) return arg_fail("geometry too large");
  p.N = (int)n_off; p.K = (int)k_off; p.tiles_per_image = (int)tile0; p.rtiles_per_image = (int)rtile0;
  p.row_stride = cfg->mode == MEHHUA_MODE_ALL ? cfg->pair_cap : p.K;
  if (cfg->mode == MEHHUA_MODE_ALL)
    for (int s = 0; s < p.S; ++s)
      if (p.lv[s].n >= (1 << 28)) return arg_fail("level too large for Entropy_ALL");
  p.nms_pre = cfg->nms_pre; p.max_per_img = cfg->max_per_img; p.pair_cap = cfg->pair_cap;
  p.n_samples = cfg->n_samples; p.use_lambda = cfg->use_lambda;
  p.agg_object = cfg->agg_object; p.agg_scale = cfg->agg_scale; p.agg_class = cfg->agg_class;
  p.cls_w = cfg->cls_w; p.rescale = cfg->rescale; p.act = cfg->activation; p.mode = cfg->mode;
  p.score_thr = cfg->score_thr; p.nms_iou = cfg->nms_iou; p.fg_thr = cfg->fg_thr;
  p.obj_thr = cfg->obj_thr; p.cluster_iou = cfg->cluster_iou;
  p.lambda_scale = cfg->lambda_scale; p.lambda_eps = cfg->lambda_eps;
  for (int i = 0; i < 4; ++i) { p.means[i] = cfg->means[i]; p.stds[i] = cfg->stds[i]; }
  if (!(cfg->wh_ratio_clip > 0.f)) return arg_fail("wh_ratio_clip");
  p.max_ratio = (float)std::fabs(std::log((double)cfg->wh_ratio_clip));
  p.seed = cfg->seed;
  *out = p;
  return 0;
}

size_t carve(const Plan& p, void* base, Workspace* ws) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  // the status word and the work counter come first: their addresses do not depend on the batch size,
  // so a workspace used with different B (host-buffer chunks, partial last batches) keeps one status
  const size_t o_status = take(sizeof(unsigned));
  const size_t o_work = take(4 * sizeof(int));
  const size_t o_k2done = take((size_t)kK2SplitPairs * sizeof(int));
  const size_t o_k2part = take(k2_part_floats(p.C) * sizeof(float));
  const size_t o_keys = take((size_t)p.B * p.N * sizeof(float));
  const size_t o_cand = take((size_t)p.B * p.K * p.num_fg * sizeof(unsigned long long));
  const size_t o_cnt = take((size_t)p.B * sizeof(int));
  const size_t o_maxc = take((size_t)p.B * sizeof(unsigned));
  const size_t o_inv = take((size_t)p.B * p.N * sizeof(int));
  const bool all_mode = p.mode == MEHHUA_MODE_ALL;      // Entropy_ALL family: foreground bit masks, tile counts and prefixes
  const size_t o_fgm = take(all_mode ? (size_t)p.B * p.tiles_per_image * (kKaThreads / 32) * sizeof(unsigned) : 0);
  const size_t o_tcnt = take(all_mode ? (size_t)p.B * p.tiles_per_image * sizeof(int) : 0);
  const size_t o_tpre = take(all_mode ? (size_t)p.B * p.tiles_per_image * sizeof(int) : 0);
  const size_t o_lamp = take((size_t)p.B * p.tiles_per_image * sizeof(float));
  const size_t o_tau = take((size_t)p.B * p.S * sizeof(float));
  const size_t o_capc = take((size_t)p.B * p.S * sizeof(int));
  const size_t o_slot = take((size_t)p.B * p.K * sizeof(int));
  const size_t o_ccomp = take((size_t)p.B * p.n_cap_levels * kCapRows * sizeof(unsigned long long));
  const size_t o_cscore = take((size_t)p.B * p.n_cap_levels * kCapRows * cap_row_floats(p.C) * sizeof(float));
  if (ws) {
    char* b = static_cast<char*>(base);
    ws->keys = reinterpret_cast<float*>(b + o_keys);
    ws->cand = reinterpret_cast<unsigned long long*>(b + o_cand);
    ws->cand_cnt = reinterpret_cast<int*>(b + o_cnt);
    ws->cand_maxc = reinterpret_cast<unsigned*>(b + o_maxc);
    ws->status = reinterpret_cast<unsigned*>(b + o_status);
    ws->work_counter = reinterpret_cast<int*>(b + o_work);
    ws->k2_done = reinterpret_cast<int*>(b + o_k2done);
    ws->k2_part = reinterpret_cast<float*>(b + o_k2part);
    ws->inv_map = reinterpret_cast<int*>(b + o_inv);
    ws->fg_mask = reinterpret_cast<unsigned*>(b + o_fgm);
    ws->tile_cnt = reinterpret_cast<int*>(b + o_tcnt);
    ws->tile_pref = reinterpret_cast<int*>(b + o_tpre);
    ws->lam_part = reinterpret_cast<float*>(b + o_lamp);
    ws->tau = reinterpret_cast<float*>(b + o_tau);
    ws->cap_cnt = reinterpret_cast<int*>(b + o_capc);
    ws->row_slot = reinterpret_cast<int*>(b + o_slot);
    ws->cap_comp = reinterpret_cast<unsigned long long*>(b + o_ccomp);
    ws->cap_scores = reinterpret_cast<float*>(b + o_cscore);
    ws->bytes = off;
  }
  return off;
}

int check_device() {
  static std::atomic<unsigned long long> ok_mask{0};   // devices (by ordinal < 64) already verified
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { cuda_fail(e, "cudaGetDevice"); return MEHHUA_E_NODEVICE; }
  if (dev < 64 && (ok_mask.load(std::memory_order_relaxed) >> dev) & 1ull) return 0;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { cuda_fail(e, "cudaGetDeviceProperties"); return MEHHUA_E_NODEVICE; }
  if (prop.major != 10) {
    snprintf(g_err, sizeof(g_err), "mehhua kernels are built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
    return MEHHUA_E_NODEVICE;
  }
  if (dev < 64) ok_mask.fetch_or(1ull << dev, std::memory_order_relaxed);
  return 0;
}

int sm_count() {
  static std::atomic<int> cached[64];
  int dev = 0;
  cudaGetDevice(&dev);
  int n = (dev >= 0 && dev < 64) ? cached[dev].load(std::memory_order_relaxed) : 0;
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    if (dev >= 0 && dev < 64) cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// Opt-in dynamic shared memory is a per-device function attribute: remember, per (device, kernel),
// the largest size already granted.  Thread-safe; one map lookup per launch.
template <typename F>
int ensure_dyn_smem(F* kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> granted;
  int dev = 0;
  CU(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = granted[{dev, reinterpret_cast<const void*>(kernel)}];
  if (bytes > cur) {
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cur = bytes;
  }
  return 0;
}

struct Prepared {
  Plan plan;
  Workspace ws;
  cudaStream_t stream;
};

int prepare(const mehhua_config_t* cfg, const mehhua_level_t* levels, int B, bool need_ptrs, void* workspace,
            size_t workspace_bytes, void* stream, Prepared* out) {
  int rc = check_device();
  if (rc) return rc;
  rc = build_plan(cfg, levels, B, need_ptrs, &out->plan);
  if (rc) return rc;
  if (!workspace) return arg_fail("null workspace");
  if (carve(out->plan, workspace, &out->ws) > workspace_bytes) {
    snprintf(g_err, sizeof(g_err), "workspace too small: need %zu bytes", out->ws.bytes);
    return MEHHUA_E_WORKSPACE;
  }
  out->stream = static_cast<cudaStream_t>(stream);
  return 0;
}

template <int C, int HEAD>
int launch_k1_typed(const Plan& p, const Workspace& ws, const float* img_shapes, const float* scale_factors,
                    const mehhua_buffers_t* o, cudaStream_t st) {
  timer_mark(st, 0);
  if (p.n_cap_levels > 0) {      // capture thresholds from a 1/32 sample of the sparse top-k levels
    CU(cudaMemsetAsync(ws.cap_cnt, 0, (size_t)p.B * p.S * sizeof(int), st));
    if (int rc = ensure_dyn_smem(k1t_threshold_kernel<C, HEAD>, kThrSmem)) return rc;
    k1t_threshold_kernel<C, HEAD><<<dim3(p.S, p.B), kThrThreads, kThrSmem, st>>>(p, ws.tau, ws.status);
    LAUNCHED("k1t_threshold_kernel");
  }
  k1a_keys_kernel<C, HEAD><<<p.B * p.tiles_per_image, kK1aThreads, 0, st>>>(
      p, ws.keys, o->level_fg, reinterpret_cast<unsigned*>(o->level_maxconf),
      p.n_cap_levels > 0 ? ws.tau : nullptr, ws.cap_cnt, ws.cap_comp, ws.cap_scores);
  LAUNCHED("k1a_keys_kernel");
  timer_mark(st, 1);
  bool any_topk = false;
  for (int s = 0; s < p.S; ++s) any_topk |= p.lv[s].topk != 0;
  if (any_topk) {
    if (int rc = ensure_dyn_smem(k1b_select_kernel, kSelSmem)) return rc;
    k1b_select_kernel<<<dim3(p.S, p.B), kSelThreads, kSelSmem, st>>>(p, ws.keys, o->topk_idx, ws.inv_map, ws.cap_cnt,
                                                                     ws.cap_comp, ws.row_slot, ws.status);
    LAUNCHED("k1b_select_kernel");
  }
  timer_mark(st, 2);
  bool any_gather = false;
  for (int s = 0; s < p.S; ++s) any_gather |= p.lv[s].rescan == 0;
  // kept rows of the capture levels: parked records, staged by bulk-async copies (typed class counts only)
  const bool parked = C > 0 && p.n_cap_levels > 0 && !no_parked_kernel();
  if (parked) {
    const dim3 grid((p.K + 32 * kParkWarps - 1) / (32 * kParkWarps), p.B);
    if (parked_bulk()) {
      if (int rc = ensure_dyn_smem(k1c_parked_kernel<HEAD, true>, k1c_parked_smem(p.C, true))) return rc;
      k1c_parked_kernel<HEAD, true><<<grid, 32 * kParkWarps, k1c_parked_smem(p.C, true), st>>>(
          p, img_shapes, scale_factors, o->topk_idx, o->score_rows, o->lam_rows, o->boxes, o->row_max, o->row_argmax,
          ws.cand, ws.cand_cnt, ws.cand_maxc, ws.row_slot, ws.cap_scores);
    } else {
      k1c_parked_kernel<HEAD, false><<<grid, 32 * kParkWarps, k1c_parked_smem(p.C, false), st>>>(
          p, img_shapes, scale_factors, o->topk_idx, o->score_rows, o->lam_rows, o->boxes, o->row_max, o->row_argmax,
          ws.cand, ws.cand_cnt, ws.cand_maxc, ws.row_slot, ws.cap_scores);
    }
    LAUNCHED("k1c_parked_kernel");
  }
  if (any_gather) {
    k1c_gather_kernel<C, HEAD><<<dim3((p.K + kGatherThreads - 1) / kGatherThreads, p.B), kGatherThreads, 0, st>>>(
        p, img_shapes, scale_factors, o->topk_idx, o->score_rows, o->lam_rows, o->boxes, o->row_max,
        o->row_argmax, ws.cand, ws.cand_cnt, ws.cand_maxc, ws.row_slot, ws.cap_scores, parked ? 1 : 0);
    LAUNCHED("k1c_gather_kernel");
  }
  if (p.rtiles_per_image > 0) {
    k1c_rescan_kernel<C, HEAD><<<p.B * p.rtiles_per_image, kRescanThreads, 0, st>>>(
        p, img_shapes, scale_factors, ws.inv_map, o->topk_idx, o->score_rows, o->lam_rows, o->boxes, o->row_max,
        o->row_argmax, ws.cand, ws.cand_cnt, ws.cand_maxc);
    LAUNCHED("k1c_rescan_kernel");
  }
  return 0;
}

int launch_k1(const Plan& p, const Workspace& ws, const float* img_shapes, const float* scale_factors,
              const mehhua_buffers_t* o, cudaStream_t st) {
  if (!img_shapes || !scale_factors) return arg_fail("null img_shapes / scale_factors");
  if (!o || !o->score_rows || !o->lam_rows || !o->boxes || !o->topk_idx || !o->row_max || !o->row_argmax ||
      !o->level_fg)
    return arg_fail("null K1 output buffer");
  CU(cudaMemsetAsync(o->level_fg, 0, (size_t)p.B * p.S * sizeof(int), st));
  if (o->level_maxconf) CU(cudaMemsetAsync(o->level_maxconf, 0, (size_t)p.B * p.S * sizeof(float), st));
  CU(cudaMemsetAsync(ws.cand_cnt, 0, (size_t)p.B * sizeof(int), st));
  CU(cudaMemsetAsync(ws.cand_maxc, 0, (size_t)p.B * sizeof(unsigned), st));
  if (p.head == MEHHUA_HEAD_RETINA && p.act == MEHHUA_ACT_RELU_PLUS_ONE) {
    switch (p.C) {
      case 20: return launch_k1_typed<20, kHeadRpo>(p, ws, img_shapes, scale_factors, o, st);
      case 80: return launch_k1_typed<80, kHeadRpo>(p, ws, img_shapes, scale_factors, o, st);
      default: return launch_k1_typed<0, kHeadRpo>(p, ws, img_shapes, scale_factors, o, st);
    }
  }
  if (p.head == MEHHUA_HEAD_RETINA) {
    switch (p.C) {
      case 20: return launch_k1_typed<20, MEHHUA_HEAD_RETINA>(p, ws, img_shapes, scale_factors, o, st);
      case 80: return launch_k1_typed<80, MEHHUA_HEAD_RETINA>(p, ws, img_shapes, scale_factors, o, st);
      default: return launch_k1_typed<0, MEHHUA_HEAD_RETINA>(p, ws, img_shapes, scale_factors, o, st);
    }
  }
  switch (p.C) {
    case 21: return launch_k1_typed<21, MEHHUA_HEAD_SSD>(p, ws, img_shapes, scale_factors, o, st);
    case 81: return launch_k1_typed<81, MEHHUA_HEAD_SSD>(p, ws, img_shapes, scale_factors, o, st);
    default: return launch_k1_typed<0, MEHHUA_HEAD_SSD>(p, ws, img_shapes, scale_factors, o, st);
  }
}

template <int C, int HEAD, int ACT>
int launch_all_typed(const Plan& p, const Workspace& ws, const mehhua_buffers_t* o, cudaStream_t st) {
  ka_fg_kernel<C, HEAD, ACT><<<p.B * p.tiles_per_image, kKaThreads, 0, st>>>(
      p, ws.fg_mask, ws.tile_cnt, ws.lam_part, reinterpret_cast<unsigned*>(o->level_maxconf));
  LAUNCHED("ka_fg_kernel");
  ka_scan_kernel<<<p.B, kAllScanThreads, 0, st>>>(p, ws.tile_cnt, ws.tile_pref, ws.lam_part, o->pair_off, o->level_fg,
                                                  o->lam_mean, o->n_obj, o->n_det, ws.status);
  LAUNCHED("ka_scan_kernel");
  ka_rows_kernel<C, HEAD, ACT><<<p.B * p.tiles_per_image, kKaThreads, 0, st>>>(
      p, ws.fg_mask, ws.tile_cnt, ws.tile_pref, o->pair_off, o->score_rows, o->lam_rows, o->topk_idx, o->row_max,
      o->row_argmax, o->pair_row, o->pair_obj, o->pair_cls);
  LAUNCHED("ka_rows_kernel");
  return 0;
}

int launch_all(const Plan& p, const Workspace& ws, const mehhua_buffers_t* o, cudaStream_t st) {
  if (!o || !o->score_rows || !o->lam_rows || !o->topk_idx || !o->row_max || !o->row_argmax || !o->level_fg ||
      !o->pair_row || !o->pair_obj || !o->pair_cls || !o->pair_off || !o->lam_mean || !o->n_obj || !o->n_det)
    return arg_fail("null Entropy_ALL buffer");
  if (o->level_maxconf) CU(cudaMemsetAsync(o->level_maxconf, 0, (size_t)p.B * p.S * sizeof(float), st));
  if (p.head == MEHHUA_HEAD_RETINA) {
    if (p.act == MEHHUA_ACT_RELU) {
      switch (p.C) {
        case 20: return launch_all_typed<20, MEHHUA_HEAD_RETINA, MEHHUA_ACT_RELU>(p, ws, o, st);
        case 80: return launch_all_typed<80, MEHHUA_HEAD_RETINA, MEHHUA_ACT_RELU>(p, ws, o, st);
        default: return launch_all_typed<0, MEHHUA_HEAD_RETINA, MEHHUA_ACT_RELU>(p, ws, o, st);
      }
    }
    switch (p.C) {
      case 20: return launch_all_typed<20, MEHHUA_HEAD_RETINA, MEHHUA_ACT_SOFTMAX>(p, ws, o, st);
      case 80: return launch_all_typed<80, MEHHUA_HEAD_RETINA, MEHHUA_ACT_SOFTMAX>(p, ws, o, st);
      default: return launch_all_typed<0, MEHHUA_HEAD_RETINA, MEHHUA_ACT_SOFTMAX>(p, ws, o, st);
    }
  }
  switch (p.C) {
    case 21: return launch_all_typed<21, MEHHUA_HEAD_SSD, MEHHUA_ACT_SOFTMAX>(p, ws, o, st);
    case 81: return launch_all_typed<81, MEHHUA_HEAD_SSD, MEHHUA_ACT_SOFTMAX>(p, ws, o, st);
    default: return launch_all_typed<0, MEHHUA_HEAD_SSD, MEHHUA_ACT_SOFTMAX>(p, ws, o, st);
  }
}

int launch_all_reduce(const Plan& p, const mehhua_buffers_t* o, cudaStream_t st) {
  if (!o || !o->pair_cls || !o->pair_off || !o->pair_unc || !o->image_scores) return arg_fail("null Entropy_ALL buffer");
  const size_t smem = ka_reduce_smem_bytes(p.S, p.C);
  if (smem > 227 * 1024) return arg_fail("c_out too large for the Entropy_ALL reduce kernel");
  if (int rc = ensure_dyn_smem(ka_reduce_kernel, smem)) return rc;
  ka_reduce_kernel<<<p.B, kAllAggThreads, smem, st>>>(p, o->pair_cls, o->pair_off, o->pair_unc, o->image_scores, o->group_unc);
  LAUNCHED("ka_reduce_kernel");
  return 0;
}

int launch_nms(const Plan& p, const Workspace& ws, const mehhua_buffers_t* o, cudaStream_t st) {
  if (!o || !o->boxes || !o->dets || !o->det_labels || !o->det_flat || !o->n_det || !o->n_obj)
    return arg_fail("null NMS buffer");
  if (int rc = ensure_dyn_smem(k3a_nms_kernel, kNmsSmem)) return rc;
  k3a_nms_kernel<<<p.B, kNmsThreads, kNmsSmem, st>>>(p, ws.cand, ws.cand_cnt, ws.cand_maxc, o->boxes, o->dets,
                                                    o->det_labels, o->det_flat, o->n_det, o->n_obj, ws.status);
  LAUNCHED("k3a_nms_kernel");
  return 0;
}

int launch_pairs(const Plan& p, const Workspace& ws, const mehhua_buffers_t* o, cudaStream_t st) {
  if (!o || !o->boxes || !o->row_max || !o->row_argmax || !o->lam_rows || !o->level_fg || !o->dets || !o->n_obj ||
      !o->pair_row || !o->pair_obj || !o->pair_cls || !o->pair_off || !o->lam_mean)
    return arg_fail("null pair buffer");
  if (int rc = ensure_dyn_smem(k3b_pairs_kernel, kPairSmem)) return rc;
  k3b_pairs_kernel<<<p.B, kPairThreads, kPairSmem, st>>>(p, o->boxes, o->row_max, o->row_argmax, o->lam_rows, o->level_fg,
                                                 o->dets, o->n_obj, o->pair_row, o->pair_obj, o->pair_cls,
                                                 o->pair_off, o->lam_mean, ws.status);
  LAUNCHED("k3b_pairs_kernel");
  return 0;
}

int launch_k2(const Plan& p, const Workspace& ws, const int64_t* image_ids, const float* inj,
              const int64_t* inj_off, const mehhua_buffers_t* o, cudaStream_t st) {
  if (!o || !o->score_rows || !o->lam_rows || !o->lam_mean || !o->pair_row || !o->pair_obj || !o->pair_off ||
      !o->pair_unc)
    return arg_fail("null K2 buffer");
  if (p.C > 256) return arg_fail("c_out must be <= 256 for the K2 class lists");
  const size_t smem = k2_smem_bytes(p.C, p.B);
  if (smem > 227 * 1024) return arg_fail("c_out too large for the K2 shared-memory layout");
  if (int rc = ensure_dyn_smem(k2_dirichlet_kernel<false>, smem)) return rc;
  if (int rc = ensure_dyn_smem(k2_dirichlet_kernel<true>, smem)) return rc;
  int blocks_per_sm = 1;
  {   // resident blocks per SM for this (device, shared-memory size): the persistent grid's width
    static std::mutex mu;
    static std::map<std::pair<int, size_t>, int> cache;
    int dev = 0;
    CU(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    int& v = cache[{dev, smem}];
    if (v == 0) {
      int v0 = 0, v1 = 0;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v0, k2_dirichlet_kernel<false>, kK2Threads, smem));
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v1, k2_dirichlet_kernel<true>, kK2Threads, smem));
      v = std::max(1, std::min(v0, v1));
    }
    blocks_per_sm = v;
  }
  // work_counter[0] / [1]: the queues of the per-pair and the per-(pair, sub-range) instantiation
  CU(cudaMemsetAsync(ws.work_counter, 0, 2 * sizeof(int), st));
  const long long* ids = reinterpret_cast<const long long*>(image_ids);
  const long long* ioff = reinterpret_cast<const long long*>(inj_off);
  const int grid = sm_count() * blocks_per_sm;
  // the pair count lives on the device: for small batches both regimes are launched and the one that
  // does not apply returns at once.  Large batches and injected samples (tests) take the per-pair form.
  const int allow_split = (inj == nullptr && p.n_samples > 0 && p.B <= kK2SplitMaxBatch) ? 1 : 0;
  k2_dirichlet_kernel<false><<<grid, kK2Threads, smem, st>>>(
      p, o->score_rows, o->lam_rows, o->lam_mean, o->pair_row, o->pair_obj, o->pair_off, ids, inj, ioff,
      o->pair_unc, o->pair_avg, ws.work_counter, ws.k2_part, ws.k2_done, allow_split, ws.status);
  LAUNCHED("k2_dirichlet_kernel");
  if (allow_split) {
    CU(cudaMemsetAsync(ws.k2_done, 0, (size_t)kK2SplitPairs * sizeof(int), st));
    k2_dirichlet_kernel<true><<<grid, kK2Threads, smem, st>>>(
        p, o->score_rows, o->lam_rows, o->lam_mean, o->pair_row, o->pair_obj, o->pair_off, ids, inj, ioff,
        o->pair_unc, o->pair_avg, ws.work_counter + 1, ws.k2_part, ws.k2_done, allow_split, ws.status);
    LAUNCHED("k2_dirichlet_kernel(split)");
  }
  return 0;
}

int launch_hua(const Plan& p, const Workspace& ws, const mehhua_buffers_t* o, cudaStream_t st) {
  if (p.mode == MEHHUA_MODE_ALL) return launch_all_reduce(p, o, st);     // Entropy_ALL family: (level, class) means, any count
  if (!o || !o->pair_row || !o->pair_obj || !o->pair_cls || !o->pair_off || !o->pair_unc || !o->n_obj ||
      !o->image_scores)
    return arg_fail("null HUA buffer");
  const size_t smem = k3c_smem_bytes(kHuaCap, p.C), smem_small = k3c_smem_bytes(kHuaCapSmall, p.C);
  if (smem > 227 * 1024) return arg_fail("c_out too large for the K3c shared-memory layout");
  if (int rc = ensure_dyn_smem(k3c_hua_kernel<kHuaCap>, smem)) return rc;
  if (int rc = ensure_dyn_smem(k3c_hua_kernel<kHuaCapSmall>, smem_small)) return rc;
  // both instantiations: images with few pairs are served by the small one (all resident at once), the others by the large one
  k3c_hua_kernel<kHuaCapSmall><<<p.B, kHuaThreads, smem_small, st>>>(p, o->pair_row, o->pair_obj, o->pair_cls, o->pair_off,
                                                                   o->pair_unc, o->n_obj, o->image_scores, ws.status);
  LAUNCHED("k3c_hua_kernel");
  k3c_hua_kernel<kHuaCap><<<p.B, kHuaThreads, smem, st>>>(p, o->pair_row, o->pair_obj, o->pair_cls, o->pair_off,
                                                          o->pair_unc, o->n_obj, o->image_scores, ws.status);
  LAUNCHED("k3c_hua_kernel");
  return 0;
}

}  // namespace

extern "C" {

int mehhua_abi_version(void) { return MEHHUA_ABI_VERSION; }
const char* mehhua_last_cuda_error(void) { return g_err; }
uint64_t mehhua_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int64_t mehhua_rows_per_image(const mehhua_config_t* cfg, const mehhua_level_t* levels) {
  Plan p;
  mehhua_config_t c = *cfg;
  if (c.pair_cap < 1) c.pair_cap = 1;
  const int rc = build_plan(&c, levels, 1, false, &p);
  return rc ? rc : p.K;
}

size_t mehhua_workspace_bytes(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B) {
  Plan p;
  if (build_plan(cfg, levels, B, false, &p)) return 0;
  return carve(p, nullptr, nullptr);
}

int mehhua_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
  if (!workspace) return arg_fail("null workspace");
  CU(cudaMemsetAsync(workspace, 0, workspace_bytes, static_cast<cudaStream_t>(stream)));
  return 0;
}

int mehhua_read_status(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B, void* workspace,
                       void* stream, uint32_t* status_out) {
  if (!workspace || !status_out) return arg_fail("null status pointer");
  Plan p;
  int rc = build_plan(cfg, levels, B, false, &p);
  if (rc) return rc;
  Workspace ws;
  carve(p, workspace, &ws);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned v = 0;
  CU(cudaMemcpyAsync(&v, ws.status, sizeof(v), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (v) CU(cudaMemsetAsync(ws.status, 0, sizeof(v), st));
  *status_out = v;
  return 0;
}

int mehhua_k1_alpha_topk(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                         const float* img_shapes, const float* scale_factors, const mehhua_buffers_t* out,
                         void* workspace, size_t workspace_bytes, void* stream) {
  if (cfg && cfg->mode != MEHHUA_MODE_NMS) return arg_fail("cfg->mode must be MEHHUA_MODE_NMS for this entry point");
  Prepared pr;
  int rc = prepare(cfg, levels, B, true, workspace, workspace_bytes, stream, &pr);
  if (rc) return rc;
  return launch_k1(pr.plan, pr.ws, img_shapes, scale_factors, out, pr.stream);
}

int mehhua_nms_objects(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                       const mehhua_buffers_t* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (cfg && cfg->mode != MEHHUA_MODE_NMS) return arg_fail("cfg->mode must be MEHHUA_MODE_NMS for this entry point");
  Prepared pr;
  int rc = prepare(cfg, levels, B, false, workspace, workspace_bytes, stream, &pr);
  if (rc) return rc;
  return launch_nms(pr.plan, pr.ws, out, pr.stream);
}

int mehhua_iou_pairs(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                     const mehhua_buffers_t* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (cfg && cfg->mode != MEHHUA_MODE_NMS) return arg_fail("cfg->mode must be MEHHUA_MODE_NMS for this entry point");
  Prepared pr;
  int rc = prepare(cfg, levels, B, false, workspace, workspace_bytes, stream, &pr);
  if (rc) return rc;
  return launch_pairs(pr.plan, pr.ws, out, pr.stream);
}

int mehhua_k2_dirichlet_epi(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                            const int64_t* image_ids, const float* inj_samples, const int64_t* inj_off,
                            const mehhua_buffers_t* out, void* workspace, size_t workspace_bytes, void* stream) {
  Prepared pr;
  int rc = prepare(cfg, levels, B, false, workspace, workspace_bytes, stream, &pr);
  if (rc) return rc;
  return launch_k2(pr.plan, pr.ws, image_ids, inj_samples, inj_off, out, pr.stream);
}

int mehhua_k3_hua(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                  const mehhua_buffers_t* out, void* workspace, size_t workspace_bytes, void* stream) {
  Prepared pr;
  int rc = prepare(cfg, levels, B, false, workspace, workspace_bytes, stream, &pr);
  if (rc) return rc;
  return launch_hua(pr.plan, pr.ws, out, pr.stream);
}

int mehhua_score_batch(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                       const float* img_shapes, const float* scale_factors, const int64_t* image_ids,
                       const mehhua_buffers_t* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (cfg && cfg->mode != MEHHUA_MODE_NMS) return arg_fail("cfg->mode must be MEHHUA_MODE_NMS for this entry point");
  Prepared pr;
  int rc = prepare(cfg, levels, B, true, workspace, workspace_bytes, stream, &pr);
  if (rc) return rc;
  if ((rc = launch_k1(pr.plan, pr.ws, img_shapes, scale_factors, out, pr.stream))) return rc;
  timer_mark(pr.stream, 3);
  if ((rc = launch_nms(pr.plan, pr.ws, out, pr.stream))) return rc;
  timer_mark(pr.stream, 4);
  if ((rc = launch_pairs(pr.plan, pr.ws, out, pr.stream))) return rc;
  timer_mark(pr.stream, 5);
  if ((rc = launch_k2(pr.plan, pr.ws, image_ids, nullptr, nullptr, out, pr.stream))) return rc;
  timer_mark(pr.stream, 6);
  rc = launch_hua(pr.plan, pr.ws, out, pr.stream);
  timer_mark(pr.stream, 7);
  if (g_timer.on && g_timer.calls < g_timer.max_calls) ++g_timer.calls;
  return rc;
}

int mehhua_all_fg_rows(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                       const mehhua_buffers_t* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (cfg && cfg->mode != MEHHUA_MODE_ALL) return arg_fail("cfg->mode must be MEHHUA_MODE_ALL");
  Prepared pr;
  int rc = prepare(cfg, levels, B, true, workspace, workspace_bytes, stream, &pr);
  if (rc) return rc;
  return launch_all(pr.plan, pr.ws, out, pr.stream);
}

int mehhua_score_batch_all(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                           const int64_t* image_ids, const mehhua_buffers_t* out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (cfg && cfg->mode != MEHHUA_MODE_ALL) return arg_fail("cfg->mode must be MEHHUA_MODE_ALL");
  Prepared pr;
  int rc = prepare(cfg, levels, B, true, workspace, workspace_bytes, stream, &pr);
  if (rc) return rc;
  if ((rc = launch_all(pr.plan, pr.ws, out, pr.stream))) return rc;
  if ((rc = launch_k2(pr.plan, pr.ws, image_ids, nullptr, nullptr, out, pr.stream))) return rc;
  return launch_hua(pr.plan, pr.ws, out, pr.stream);
}

int mehhua_stage_timing_begin(int32_t max_calls) {
  if (max_calls < 1 || max_calls > 100000) return arg_fail("max_calls");
  for (cudaEvent_t e : g_timer.ev) cudaEventDestroy(e);
  g_timer.ev.assign((size_t)max_calls * (kStages + 1), nullptr);
  for (auto& e : g_timer.ev) CU(cudaEventCreate(&e));
  g_timer.calls = 0;
  g_timer.max_calls = (size_t)max_calls;
  g_timer.on = true;
  return 0;
}

int mehhua_stage_timing_end(double* ms_sum, int32_t* calls_out) {
  if (!ms_sum || !calls_out) return arg_fail("null timing output");
  g_timer.on = false;
  for (int i = 0; i < kStages; ++i) ms_sum[i] = 0.0;
  for (size_t c = 0; c < g_timer.calls; ++c) {
    cudaEvent_t* e = &g_timer.ev[c * (kStages + 1)];
    CU(cudaEventSynchronize(e[kStages]));
    for (int i = 0; i < kStages; ++i) {
      float ms = 0.f;
      CU(cudaEventElapsedTime(&ms, e[i], e[i + 1]));
      ms_sum[i] += ms;
    }
  }
  *calls_out = (int32_t)g_timer.calls;
  for (cudaEvent_t e : g_timer.ev) cudaEventDestroy(e);
  g_timer.ev.clear();
  g_timer.calls = g_timer.max_calls = 0;
  return 0;
}

// any k <= n is served: small pools by one block (256 bytes of status), large ones by the grid-wide form
size_t mehhua_pool_topk_workspace_bytes(int64_t n) {
  if (n < kPoolMultiMin) return 256;
  return k4m_workspace_bytes((long long)n, (int)std::min<int64_t>(n, 0x7fffffffll));
}

size_t mehhua_pool_topk_workspace_bytes_k(int64_t n, int32_t k) {
  if (n < kPoolMultiMin || k <= 0) return 256;
  return k4m_workspace_bytes((long long)n, (int)std::min<int64_t>(n, (int64_t)k));
}

int mehhua_k4_pool_topk(const float* scores, const uint8_t* mask, int64_t n, int32_t k, int64_t* idx_out,
                        int32_t* n_selected_out, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_device();
  if (rc) return rc;
  if (!scores || !idx_out || !n_selected_out || !workspace) return arg_fail("null K4 pointer");
  if (workspace_bytes < 256) return MEHHUA_E_WORKSPACE;
  if (n < 0 || n > 0x7fffffffll || k < 0) return arg_fail("pool size / k");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k > n) k = (int32_t)n;
  // large pools: grid-wide radix select + chunk sort + merges, when the workspace holds its buffers
  if (n >= kPoolMultiMin && k > 0 && workspace_bytes >= k4m_workspace_bytes((long long)n, k)) {
    unsigned char* w = static_cast<unsigned char*>(workspace);
    PoolState* ps = reinterpret_cast<PoolState*>(w);
    int* ghist = reinterpret_cast<int*>(w + 256);
    const size_t cap = k4m_buf_elems((long long)n, k);
    unsigned long long* buf0 = reinterpret_cast<unsigned long long*>(w + k4m_state_bytes());
    unsigned long long* buf1 = buf0 + cap;
    CU(cudaMemsetAsync(w, 0, k4m_state_bytes(), st));
    int dev = 0, sms = 148;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long per_block = (long long)kPoolHistThreads * kSelUnroll;
    const int grid = (int)std::max<long long>(1, std::min<long long>(2ll * sms, (n + per_block - 1) / per_block));
    for (int pass = 0; pass < kPoolPasses; ++pass) {
      k4m_hist_kernel<<<grid, kPoolHistThreads, 0, st>>>(scores, mask, (long long)n, k, pass, ps, ghist);
      LAUNCHED("k4m_hist_kernel");
    }
    k4m_compact_kernel<<<grid, kPoolHistThreads, 0, st>>>(scores, mask, (long long)n, ps, buf0, (long long)cap);
    LAUNCHED("k4m_compact_kernel");
    const int chunks = (int)(cap / kPoolChunk);
    if (int rc2 = ensure_dyn_smem(k4m_chunk_sort_kernel, (size_t)kPoolChunk * 8)) return rc2;
    k4m_chunk_sort_kernel<<<chunks, 1024, (size_t)kPoolChunk * 8, st>>>(ps, buf0);
    LAUNCHED("k4m_chunk_sort_kernel");
    unsigned long long *src = buf0, *dst = buf1;
    const int mgrid = (int)((cap / kPoolMergeVT + 255) / 256);
    for (long long run = kPoolChunk; run < (long long)cap; run *= 2) {
      k4m_merge_kernel<<<mgrid, 256, 0, st>>>(ps, src, dst, run);
      LAUNCHED("k4m_merge_kernel");
      std::swap(src, dst);
    }
    k4m_output_kernel<<<std::max(1, std::min(sms, (k + 255) / 256)), 256, 0, st>>>(ps, src, k, reinterpret_cast<long long*>(idx_out),
                                                                                   n_selected_out);
    LAUNCHED("k4m_output_kernel");
    return 0;
  }
  if (int rc = ensure_dyn_smem(k4_pool_topk_kernel, kPoolSmem)) return rc;
  k4_pool_topk_kernel<<<1, kPoolThreads, kPoolSmem, st>>>(scores, mask, (long long)n, k,
                                                         reinterpret_cast<long long*>(idx_out), n_selected_out,
                                                         static_cast<unsigned*>(workspace));
  LAUNCHED("k4_pool_topk_kernel");
  return 0;
}

// debug: rows parked per (image, level) by the last K1 call on this workspace, -1 for levels not in capture mode
int mehhua_debug_capture_counts(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B, void* workspace,
                                void* stream, int32_t* counts_out) {
  if (!workspace || !counts_out) return arg_fail("null pointer");
  Plan p;
  int rc = build_plan(cfg, levels, B, false, &p);
  if (rc) return rc;
  Workspace ws;
  carve(p, workspace, &ws);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(cudaMemcpyAsync(counts_out, ws.cap_cnt, (size_t)B * p.S * sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int b = 0; b < B; ++b)
    for (int s = 0; s < p.S; ++s)
      if (p.lv[s].cap < 0) counts_out[b * p.S + s] = -1;
  return 0;
}

// KM: ensemble / MC-dropout mutual-information baseline
static int mi_plan(const float* const* member_logits, int32_t M, const mehhua_level_t* lv, int32_t S, int32_t n_cls,
                   int32_t B, bool need_ptrs, MiPlan* out) {
  if (!lv || S < 1 || S > kMaxLevels || M < 1 || M > kMiMaxMembers || n_cls < 1 || B < 1) return arg_fail("MI geometry");
  MiPlan& p = *out;
  memset(&p, 0, sizeof(p));
  p.S = S; p.B = B; p.M = M; p.n_cls = n_cls;
  long long tile0 = 0;
  for (int s = 0; s < S; ++s) {
    if (lv[s].H < 1 || lv[s].W < 1 || lv[s].A < 1) return arg_fail("level geometry");
    p.HW[s] = lv[s].H * lv[s].W; p.A[s] = lv[s].A;
    p.tpp[s] = (p.HW[s] + kMiThreads - 1) / kMiThreads;
    p.tile0[s] = (int)tile0;
    tile0 += (long long)p.tpp[s] * p.A[s];
    for (int m = 0; m < M; ++m) {
      if (need_ptrs && (!member_logits || !member_logits[m * S + s])) return arg_fail("null member logits");
      p.logits[m][s] = member_logits ? member_logits[m * S + s] : nullptr;
    }
  }
  if (tile0 * B > 0x7fffffffll) return arg_fail("geometry too large");
  p.tiles_per_image = (int)tile0;
  return 0;
}

size_t mehhua_mi_workspace_bytes(const mehhua_level_t* levels, int32_t num_levels, int32_t B) {
  MiPlan p;
  if (mi_plan(nullptr, 1, levels, num_levels, 1, B, false, &p)) return 0;
  return ((size_t)B * p.tiles_per_image * sizeof(float) + 255) & ~(size_t)255;
}

int mehhua_mi_score_batch(const float* const* member_logits, int32_t n_members, const mehhua_level_t* levels,
                          int32_t num_levels, int32_t n_cls, int32_t B, float* level_mi, float* image_scores,
                          void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_device();
  if (rc) return rc;
  if (!image_scores || !workspace) return arg_fail("null MI pointer");
  MiPlan p;
  rc = mi_plan(member_logits, n_members, levels, num_levels, n_cls, B, true, &p);
  if (rc) return rc;
  if (workspace_bytes < mehhua_mi_workspace_bytes(levels, num_levels, B)) return MEHHUA_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* tile_sum = static_cast<float*>(workspace);
  km_mi_tiles_kernel<<<B * p.tiles_per_image, kMiThreads, 0, st>>>(p, tile_sum);
  LAUNCHED("km_mi_tiles_kernel");
  km_mi_finish_kernel<<<B, 32 * kMaxLevels, 0, st>>>(p, tile_sum, level_mi, image_scores);
  LAUNCHED("km_mi_finish_kernel");
  return 0;
}

// debug / known-answer entry for the Philox block (tests only): out = 8 host uint32 (10 rounds, then 7)
int mehhua_debug_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[8]) {
  int rc = check_device();
  if (rc) return rc;
  unsigned* d = nullptr;
  CU(cudaMalloc(&d, 32));
  philox_kat_kernel<<<1, 1>>>(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]), d);
  LAUNCHED("philox_kat_kernel");
  cudaError_t e = cudaMemcpy(out, d, 32, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// host-buffer context
// ---------------------------------------------------------------------------------------------
constexpr int kHostChunks = 4;      // a host batch is uploaded in up to 4 image chunks; chunk j is scored
                                    // while chunk j+1 is still on the PCIe link
struct mehhua_host_ctx {
  mehhua_config_t cfg;
  mehhua_level_t shapes[MEHHUA_MAX_LEVELS];
  mehhua_level_t dev[MEHHUA_MAX_LEVELS];
  int max_batch;
  int64_t K;
  cudaStream_t stream;       // compute (and result read-back)
  cudaStream_t copy_stream;  // host -> device input chunks, overlapped with the compute of earlier chunks
  cudaEvent_t chunk_ready[kHostChunks];
  void* arena;          // one device allocation holding inputs, outputs, workspace
  size_t arena_bytes;
  void* workspace;
  size_t workspace_bytes;
  unsigned* status;
  float* img_shapes;
  float* scale_factors;
  int64_t* image_ids;
  mehhua_buffers_t bufs;
  bool anchors_loaded;
};

int mehhua_host_ctx_create(const mehhua_config_t* cfg, const mehhua_level_t* level_shapes, int32_t max_batch,
                           mehhua_host_ctx_t** ctx_out) {
  int rc = check_device();
  if (rc) return rc;
  if (!ctx_out) return arg_fail("null ctx_out");
  if (cfg && cfg->mode != MEHHUA_MODE_NMS) return arg_fail("the host-buffer context serves the Entropy_NMS route: cfg->mode must be MEHHUA_MODE_NMS");
  Plan p;
  if ((rc = build_plan(cfg, level_shapes, max_batch, false, &p))) return rc;
  mehhua_host_ctx* c = new (std::nothrow) mehhua_host_ctx();
  if (!c) return arg_fail("out of host memory");
  c->cfg = *cfg;
  c->max_batch = max_batch;
  c->K = p.K;
  c->anchors_loaded = false;
  const int B = max_batch, S = p.S, C = p.C, D = p.max_per_img;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  size_t o_log[MEHHUA_MAX_LEVELS], o_del[MEHHUA_MAX_LEVELS], o_lam[MEHHUA_MAX_LEVELS], o_anc[MEHHUA_MAX_LEVELS];
  for (int s = 0; s < S; ++s) {
    const size_t n = (size_t)p.lv[s].n;
    o_log[s] = take((size_t)B * n * C * 4);
    o_del[s] = take((size_t)B * n * 4 * 4);
    o_lam[s] = take((size_t)B * n * 4);
    o_anc[s] = take(n * 16);
  }
  const size_t K = (size_t)p.K, PC = (size_t)p.pair_cap;
  const size_t o_rows = take(B * K * C * 4), o_lamr = take(B * K * 4), o_box = take(B * K * 16);
  const size_t o_idx = take(B * K * 4), o_rmax = take(B * K * 4), o_rarg = take(B * K * 4);
  const size_t o_lfg = take((size_t)B * S * 4), o_dets = take((size_t)B * D * 20), o_dl = take((size_t)B * D * 4);
  const size_t o_df = take((size_t)B * D * 4), o_nd = take((size_t)B * 4), o_no = take((size_t)B * 4);
  const size_t o_pr = take(B * PC * 4), o_po = take(B * PC * 4), o_pc = take(B * PC * 4);
  const size_t o_poff = take((size_t)B * (S + 1) * 4), o_lm = take((size_t)B * S * 4), o_pu = take(B * PC * 12);
  const size_t o_sc = take((size_t)B * 4), o_shp = take((size_t)B * 8), o_sf = take((size_t)B * 16);
  const size_t o_ids = take((size_t)B * 8);
  c->workspace_bytes = carve(p, nullptr, nullptr);
  const size_t o_ws = take(c->workspace_bytes);
  cudaError_t e = cudaMalloc(&c->arena, off);
  if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaMalloc(host ctx arena)"); }
  c->arena_bytes = off;
  char* a = static_cast<char*>(c->arena);
  for (int s = 0; s < S; ++s) {
    c->shapes[s] = level_shapes[s];
    c->dev[s] = level_shapes[s];
    c->dev[s].logits = reinterpret_cast<float*>(a + o_log[s]);
    c->dev[s].deltas = reinterpret_cast<float*>(a + o_del[s]);
    c->dev[s].lambda = reinterpret_cast<float*>(a + o_lam[s]);
    c->dev[s].anchors = reinterpret_cast<float*>(a + o_anc[s]);
  }
  mehhua_buffers_t& b = c->bufs;
  b.score_rows = reinterpret_cast<float*>(a + o_rows);  b.lam_rows = reinterpret_cast<float*>(a + o_lamr);
  b.boxes = reinterpret_cast<float*>(a + o_box);        b.topk_idx = reinterpret_cast<int32_t*>(a + o_idx);
  b.row_max = reinterpret_cast<float*>(a + o_rmax);     b.row_argmax = reinterpret_cast<int32_t*>(a + o_rarg);
  b.level_fg = reinterpret_cast<int32_t*>(a + o_lfg);   b.dets = reinterpret_cast<float*>(a + o_dets);
  b.det_labels = reinterpret_cast<int32_t*>(a + o_dl);  b.det_flat = reinterpret_cast<int32_t*>(a + o_df);
  b.n_det = reinterpret_cast<int32_t*>(a + o_nd);       b.n_obj = reinterpret_cast<int32_t*>(a + o_no);
  b.pair_row = reinterpret_cast<int32_t*>(a + o_pr);    b.pair_obj = reinterpret_cast<int32_t*>(a + o_po);
  b.pair_cls = reinterpret_cast<int32_t*>(a + o_pc);    b.pair_off = reinterpret_cast<int32_t*>(a + o_poff);
  b.lam_mean = reinterpret_cast<float*>(a + o_lm);      b.pair_unc = reinterpret_cast<float*>(a + o_pu);
  b.image_scores = reinterpret_cast<float*>(a + o_sc);
  b.level_maxconf = nullptr;               // getMaxConf is not part of the host-buffer call
  b.pair_avg = nullptr;
  b.group_unc = nullptr;
  c->img_shapes = reinterpret_cast<float*>(a + o_shp);
  c->scale_factors = reinterpret_cast<float*>(a + o_sf);
  c->image_ids = reinterpret_cast<int64_t*>(a + o_ids);
  c->workspace = a + o_ws;
  Workspace ws;
  carve(p, c->workspace, &ws);
  c->status = ws.status;
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
  for (int j = 0; j < kHostChunks && e == cudaSuccess; ++j)
    e = cudaEventCreateWithFlags(&c->chunk_ready[j], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMemsetAsync(c->workspace, 0, c->workspace_bytes, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) { cudaFree(c->arena); delete c; return cuda_fail(e, "host ctx stream setup"); }
  *ctx_out = c;
  return 0;
}

void mehhua_host_ctx_destroy(mehhua_host_ctx_t* c) {
  if (!c) return;
  cudaStreamSynchronize(c->copy_stream);
  cudaStreamSynchronize(c->stream);
  for (int j = 0; j < kHostChunks; ++j) cudaEventDestroy(c->chunk_ready[j]);
  cudaStreamDestroy(c->copy_stream);
  cudaStreamDestroy(c->stream);
  cudaFree(c->arena);
  delete c;
}

int mehhua_score_batch_host(mehhua_host_ctx_t* c, const mehhua_level_t* lh, int32_t B, const float* img_shapes_host,
                            const float* scale_factors_host, const int64_t* image_ids_host,
                            float* image_scores_host, uint32_t* status_out) {
  if (!c || !lh || !img_shapes_host || !scale_factors_host || !image_scores_host) return arg_fail("null host pointer");
  if (B < 1 || B > c->max_batch) return arg_fail("batch exceeds the context's max_batch");
  cudaStream_t st = c->stream, cp = c->copy_stream;
  const int S = c->cfg.num_levels, C = c->cfg.c_out;
  for (int s = 0; s < S; ++s) {
    if (lh[s].H != c->shapes[s].H || lh[s].W != c->shapes[s].W || lh[s].A != c->shapes[s].A)
      return arg_fail("level geometry differs from the context's");
    if (!lh[s].logits || !lh[s].deltas || !lh[s].lambda) return arg_fail("null host level pointer");
    if (!lh[s].anchors && !c->anchors_loaded) return arg_fail("anchors must be given on the first call");
  }
  // per-batch small inputs and (when given) the anchors go first on the copy stream
  for (int s = 0; s < S; ++s) {
    const size_t n = (size_t)lh[s].H * lh[s].W * lh[s].A;
    if (lh[s].anchors)
      CU(cudaMemcpyAsync(const_cast<float*>(c->dev[s].anchors), lh[s].anchors, n * 16, cudaMemcpyHostToDevice, cp));
  }
  c->anchors_loaded = true;
  CU(cudaMemcpyAsync(c->img_shapes, img_shapes_host, (size_t)B * 8, cudaMemcpyHostToDevice, cp));
  CU(cudaMemcpyAsync(c->scale_factors, scale_factors_host, (size_t)B * 16, cudaMemcpyHostToDevice, cp));
  std::vector<int64_t> iota;
  if (!image_ids_host) {      // default Philox image id = position in the batch (not in the chunk)
    iota.resize(B);
    for (int b = 0; b < B; ++b) iota[b] = b;
    image_ids_host = iota.data();
  }
  CU(cudaMemcpyAsync(c->image_ids, image_ids_host, (size_t)B * 8, cudaMemcpyHostToDevice, cp));
  // image chunks: upload chunk j on the copy stream, score it on the compute stream as soon as it has
  // landed - the path of chunk j overlaps the upload of chunk j+1, only the last chunk's is exposed
  const int nchunks = B >= 2 * kHostChunks ? kHostChunks : 1;
  const int cb = (B + nchunks - 1) / nchunks;
  int used = 0;
  for (int j0 = 0; j0 < B; j0 += cb, ++used) {
    const int nb = std::min(cb, B - j0);
    for (int s = 0; s < S; ++s) {
      const size_t n = (size_t)lh[s].H * lh[s].W * lh[s].A;
      CU(cudaMemcpyAsync(const_cast<float*>(c->dev[s].logits) + (size_t)j0 * n * C, lh[s].logits + (size_t)j0 * n * C,
                         (size_t)nb * n * C * 4, cudaMemcpyHostToDevice, cp));
      CU(cudaMemcpyAsync(const_cast<float*>(c->dev[s].deltas) + (size_t)j0 * n * 4, lh[s].deltas + (size_t)j0 * n * 4,
                         (size_t)nb * n * 16, cudaMemcpyHostToDevice, cp));
      CU(cudaMemcpyAsync(const_cast<float*>(c->dev[s].lambda) + (size_t)j0 * n, lh[s].lambda + (size_t)j0 * n,
                         (size_t)nb * n * 4, cudaMemcpyHostToDevice, cp));
    }
    CU(cudaEventRecord(c->chunk_ready[used], cp));
    CU(cudaStreamWaitEvent(st, c->chunk_ready[used], 0));
    mehhua_level_t lv[MEHHUA_MAX_LEVELS];
    for (int s = 0; s < S; ++s) {
      const size_t n = (size_t)lh[s].H * lh[s].W * lh[s].A;
      lv[s] = c->dev[s];
      lv[s].logits = c->dev[s].logits + (size_t)j0 * n * C;
      lv[s].deltas = c->dev[s].deltas + (size_t)j0 * n * 4;
      lv[s].lambda = c->dev[s].lambda + (size_t)j0 * n;
    }
    mehhua_buffers_t bufs = c->bufs;            // chunks run back to back on one stream: they share the
    bufs.image_scores = c->bufs.image_scores + j0;   // intermediate buffers, only the scores are kept apart
    const int rc = mehhua_score_batch(&c->cfg, lv, nb, c->img_shapes + 2 * j0, c->scale_factors + 4 * j0,
                                      c->image_ids + j0, &bufs, c->workspace, c->workspace_bytes, st);
    if (rc) { cudaStreamSynchronize(cp); cudaStreamSynchronize(st); return rc; }
  }
  CU(cudaMemcpyAsync(image_scores_host, c->bufs.image_scores, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  unsigned status = 0;
  CU(cudaMemcpyAsync(&status, c->status, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (status) CU(cudaMemsetAsync(c->status, 0, 4, st));
  if (status_out) *status_out = status;
  return 0;
}

int mehhua_host_pin(void* ptr, size_t bytes) {
  if (!ptr || bytes == 0) return arg_fail("null host pointer");
  if (int rc = check_device()) return rc;
  CU(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  return 0;
}

int mehhua_host_unpin(void* ptr) {
  if (!ptr) return arg_fail("null host pointer");
  CU(cudaHostUnregister(ptr));
  return 0;
}

int mehhua_host_is_pinned(const void* ptr) {
  if (!ptr) return arg_fail("null host pointer");
  if (int rc = check_device()) return rc;
  cudaPointerAttributes a;
  CU(cudaPointerGetAttributes(&a, ptr));
  return a.type == cudaMemoryTypeHost ? 1 : 0;
}

}  // extern "C"
