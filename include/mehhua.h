/*
 * mehhua.h - C ABI of the B200-native MEH alpha -> Dirichlet uncertainty -> HUA -> pool top-k path.
 *
 * The reference (MoonLab-YH/AOD_MEH_HUA) has no native plugin layer: its operator API for this
 * path is a set of Python methods on the MMDetection dense heads.  Each entry point below names
 * the reference code it replaces (paths relative to /root/reference/mmdet/).  INTEGRATION.md shows
 * the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every buffer is owned by the caller and pre-allocated at the
 *     documented upper bound; the only hidden storage is the opaque workspace whose size is
 *     returned by mehhua_workspace_bytes();
 *   - device pointers unless the name ends in _host; work is enqueued on `stream` (a cudaStream_t
 *     passed as void*) and the call returns without synchronising, except the *_host entry points
 *     and mehhua_read_status();
 *   - return value 0 = enqueued, negative = MEHHUA_E_* (nothing enqueued);
 *   - data-dependent failures (capacity overflow) are reported through the status word, read
 *     with mehhua_read_status(): a dropped batch is never silent (cf. apis/test.py:122-128).
 *
 * Layouts: all tensors are dense row-major fp32 / int32.  Level s of a batch holds
 *   logits [B, A*C_out, H, W]   channel = a*C_out + c      (Lambda_L2.py:266 before permute)
 *   deltas [B, A*4,     H, W]   channel = a*4 + j          (Lambda_L2.py:278)
 *   lambda [B, A,       H, W]   relu output of the MEH branch (Lambda_L2.py:96-103, :267)
 *   anchors[H*W*A, 4]           prior n = (h*W + w)*A + a   (core/anchor/anchor_generator.py:367-378)
 * K_s = min(N_s, nms_pre) rows per level, concatenated level-major into K_tot rows per image.
 */
#ifndef MEHHUA_H_
#define MEHHUA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MEHHUA_ABI_VERSION 6
#define MEHHUA_MAX_LEVELS 8
#define MEHHUA_MAX_DETS 256      /* upper bound on max_per_img */
#define MEHHUA_MAX_NMS_PRE 4096  /* upper bound on nms_pre */

/* error codes (negative return values) */
#define MEHHUA_E_ARG        (-1)  /* invalid argument / unsupported configuration */
#define MEHHUA_E_WORKSPACE  (-2)  /* workspace too small */
#define MEHHUA_E_CUDA       (-3)  /* a CUDA runtime call failed (see mehhua_last_cuda_error) */
#define MEHHUA_E_NODEVICE   (-4)  /* no CUDA device / wrong architecture */

/* status word bits (mehhua_read_status) */
#define MEHHUA_ST_PAIR_OVERFLOW   1u  /* an image produced more (box, object) pairs than pair_cap */
#define MEHHUA_ST_SELECT_SLOWPATH 2u  /* info: a top-k needed extra radix passes (ties / dense bins) */
#define MEHHUA_ST_BAD_ALPHA       4u  /* info: a Dirichlet alpha <= 0 was met (treated as clamped) */
#define MEHHUA_ST_CAPTURE_FALLBACK 8u /* info: a level's capture estimate missed; its rows came from the gather path */

#define MEHHUA_HEAD_RETINA 0   /* Lambda_L2Net: C_out = C, score = p / (sum(p) + 1e-20 + 1e-9) */
#define MEHHUA_HEAD_SSD    1   /* MyLSSDHead:   C_out = C + 1 (background last), score = p     */

#define MEHHUA_MODE_NMS 0      /* uncertainty_pool = 'Entropy_NMS' (Config_RetinaNet.py:14): objects from NMS */
#define MEHHUA_MODE_ALL 1      /* uncertainty_pool = 'Entropy_ALL': every foreground prior, no objects;
                                  rows == pairs, row buffers are [B, pair_cap, ...] */

#define MEHHUA_AGG_SUM 0
#define MEHHUA_AGG_AVG 1
#define MEHHUA_AGG_MAX 2
#define MEHHUA_AGG_POOL 3  /* agg_class only, MEHHUA_MODE_ALL only: no class split - the mean over ALL foreground priors
                              of a level (ComputeAvgUnc / AggregateAvgUnc, Lambda_L2_ReLU.py:446-474, 532-541) */

/* how a row of class values is formed from the logits */
#define MEHHUA_ACT_SOFTMAX       0  /* the scoring heads: p = softmax(logits) (Lambda_L2.py:269-273, My_L_ssd_head.py:331) */
#define MEHHUA_ACT_RELU_PLUS_ONE 1  /* MEHHUA_MODE_NMS, Retina head: alpha = relu(logits) + 1, score = alpha / (sum alpha + 1e-20)
                                       - the base head's evidential form (L_anchor_head.py:401-406); detection route */
#define MEHHUA_ACT_RELU          2  /* MEHHUA_MODE_ALL, Retina head: row = relu(logits), FG = max(row / (sum + 1e-9)) > fg_thr,
                                       alpha = row * lambda' (zeros allowed) - uncertainty_pool 'Entropy_Avg' */

typedef struct mehhua_level {
  const float* logits;
  const float* deltas;
  const float* lambda;
  const float* anchors;
  int32_t H, W, A;
  int32_t reserved;
} mehhua_level_t;

/* Constants of the path; defaults = the reference's hard-coded values. */
typedef struct mehhua_config {
  int32_t head;            /* MEHHUA_HEAD_* */
  int32_t c_out;           /* cls_out_channels */
  int32_t num_levels;
  int32_t nms_pre;         /* test_cfg.nms_pre (1000); <= 0 disables the per-level top-k */
  float   score_thr;       /* test_cfg.score_thr: 0.05 Retina / 0.02 SSD, strict > */
  float   nms_iou;         /* test_cfg.nms.iou_threshold 0.5, strict > */
  int32_t max_per_img;     /* test_cfg.max_per_img 100 / 200 */
  float   fg_thr;          /* 0.3  Lambda_L2.py:500,508 */
  float   obj_thr;         /* 0.3  Lambda_L2.py:349 */
  float   cluster_iou;     /* 0.5  Lambda_L2.py:349 */
  float   lambda_scale;    /* 25   Lambda_L2.py:515 */
  float   lambda_eps;      /* 1e-7 Lambda_L2.py:514 */
  int32_t use_lambda;      /* 1; 0 = Lambda_L2_noL.py:531 */
  int32_t n_samples;       /* 500  Lambda_L2.py:520; 0 = analytic form (T -> infinity: total = H(alpha/alpha0),
                              aleatoric = psi(alpha0+1) - sum (alpha_c/alpha0) psi(alpha_c+1)), a deterministic
                              mode for set-identity tests - not a reference mode */
  int32_t agg_object, agg_scale, agg_class;  /* MEHHUA_AGG_*; 'objectSum_scaleMax_classSum' */
  int32_t cls_w;           /* clsW: multiply by the number of distinct classes (Lambda_L2.py:617) */
  float   means[4];        /* bbox_coder target_means */
  float   stds[4];         /* bbox_coder target_stds  */
  float   wh_ratio_clip;   /* 16/1000 */
  int32_t rescale;         /* divide boxes by scale_factor (Lambda_L2.py:307-308) */
  int32_t pair_cap;        /* capacity of the per-image pair list */
  int32_t mode;            /* MEHHUA_MODE_*: which uncertainty_pool route the buffers are used for */
  int32_t activation;      /* MEHHUA_ACT_* (0 = softmax, the scoring heads' form) */
  uint64_t seed;           /* Philox key of the free-running sampler */
} mehhua_config_t;

/* Caller-owned result buffers.  K_tot = sum_s K_s, S = num_levels, D = max_per_img. */
typedef struct mehhua_buffers {
  float*   score_rows;   /* [B, K_tot, C_out]  scores of the kept priors (Lambda_L2.py:295)       */
  float*   lam_rows;     /* [B, K_tot]         their lambda (Lambda_L2.py:297)                    */
  float*   boxes;        /* [B, K_tot, 4]      decoded, clipped, rescaled (Lambda_L2.py:299-308)  */
  int32_t* topk_idx;     /* [B, K_tot]         prior index inside its level (Lambda_L2.py:290)    */
  float*   row_max;      /* [B, K_tot]         max_c score (incl. background for SSD)             */
  int32_t* row_argmax;   /* [B, K_tot]         argmax_c score (Lambda_L2.py:526)                  */
  int32_t* level_fg;     /* [B, S]             any(max_c softmax > fg_thr) (Lambda_L2.py:496-502) */
  float*   dets;         /* [B, D, 5]          x1,y1,x2,y2,score in descending score              */
  int32_t* det_labels;   /* [B, D]                                                                */
  int32_t* det_flat;     /* [B, D]             row*C + class of each detection                    */
  int32_t* n_det;        /* [B]                                                                   */
  int32_t* n_obj;        /* [B]                detections with score > obj_thr                    */
  int32_t* pair_row;     /* [B, pair_cap]      row (0..K_tot) of each (box, object) pair          */
  int32_t* pair_obj;     /* [B, pair_cap]      object index                                       */
  int32_t* pair_cls;     /* [B, pair_cap]      argmax class of the row                            */
  int32_t* pair_off;     /* [B, S+1]           first pair of each level; [S] = pairs of the image */
  float*   lam_mean;     /* [B, S]             mean lambda over the level's pairs                 */
  float*   pair_unc;     /* [B, pair_cap, 3]   total, aleatoric, epistemic (Lambda_L2.py:521-525) */
  float*   image_scores; /* [B]                AggregateObjScaleUnc output                        */
  float*   level_maxconf;/* [B, S] or NULL     max over ALL priors of max_c softmax: the `output`   *
                          *                    of getMaxConf (utils/functions.py:467-476); fused   *
                          *                    into the logits pass (K1a / KA1), NULL = skipped     */
  float*   group_unc;    /* [B, S, C_out, 3] or NULL, MEHHUA_MODE_ALL: per (level, class) the number of foreground priors, *
                          *                    their mean aleatoric and mean epistemic uncertainty - the `scaleUnc`    *
                          *                    third return item of _get_bboxes (Lambda_L2.py:377-378)                  */
  float*   pair_avg;     /* [B, pair_cap, C_out] or NULL: mean_t x_c of every pair (`avg` of Lambda_L2.py:521),  *
                          *                    a diagnostic output of K2 for the per-class moment tests     */
} mehhua_buffers_t;

int         mehhua_abi_version(void);
const char* mehhua_last_cuda_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches counter) */
uint64_t    mehhua_launch_count(void);

/* rows per image after the per-level top-k (K_tot) for a geometry; levels' pointers are ignored */
int64_t mehhua_rows_per_image(const mehhua_config_t* cfg, const mehhua_level_t* levels);
size_t  mehhua_workspace_bytes(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B);
/* zero the workspace once after allocating it (counters and the status word live there) */
int     mehhua_workspace_init(void* workspace, size_t workspace_bytes, void* stream);
/* blocking: OR of MEHHUA_ST_* bits accumulated in the workspace since the last read; clears them.
 * cfg / levels / B must be the ones the workspace is used with (they fix the status word's offset) */
int     mehhua_read_status(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                           void* workspace, void* stream, uint32_t* status_out);

/* K1: fused logits -> softmax/score -> ranking key -> per-level top-k -> gather rows / lambda /
 * decoded boxes (+ level-FG flags, NMS candidates).  Replaces the per-level block of
 * _get_bboxes (Lambda_L2.py:264-326, My_L_ssd_head.py:325-361), delta2bbox
 * (core/bbox/coder/delta_xywh_bbox_coder.py:205-267) and the second softmax pass of
 * ComputeObjUnc (Lambda_L2.py:496-502).
 * img_shapes [B,2] = (H, W) used for clipping; scale_factors [B,4]. */
int mehhua_k1_alpha_topk(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                         const float* img_shapes, const float* scale_factors,
                         const mehhua_buffers_t* out, void* workspace, size_t workspace_bytes,
                         void* stream);

/* K3a: multi-class NMS -> detections and objects.  Replaces multiclass_nms
 * (core/post_processing/bbox_nms.py:7-93 -> mmcv batched_nms) and the det[:, -1] > 0.3 filter of
 * GetObjectIdx (Lambda_L2.py:344).  Needs K1's outputs in `out` and its workspace. */
int mehhua_nms_objects(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                       const mehhua_buffers_t* out, void* workspace, size_t workspace_bytes,
                       void* stream);

/* K3b: IoU clustering of kept boxes onto objects -> ordered pair list.  Replaces GetObjectIdx /
 * bbox_overlaps (Lambda_L2.py:343-349, iou2d_calculator.py:206-252) and the mask / nonzero /
 * lambda-mean prologue of ComputeObjUnc (Lambda_L2.py:503-515). */
int mehhua_iou_pairs(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                     const mehhua_buffers_t* out, void* workspace, size_t workspace_bytes,
                     void* stream);

/* K2: alpha = score * lambda', T Dirichlet samples per pair kept in registers / shared memory,
 * -> total / aleatoric / epistemic per pair.  Replaces Lambda_L2.py:517-525.
 * image_ids [B] (device, may be NULL = 0..B-1) key the Philox streams by global image id.
 * inj_samples (may be NULL): the oracle's drawn samples, one [T, P_bs, C_out] block per
 * (image, level) at element offset inj_off[b*S + s] (device int64, -1 = no block); when given
 * the kernel consumes them instead of drawing.
 * Free-running results are a pure function of (seed, image id, row, object, alpha row, T): they do
 * not depend on the batch, the launch geometry or the world size (the T samples are accumulated in
 * four fixed sub-ranges combined in order, whether one warp or four work on a pair). */
int mehhua_k2_dirichlet_epi(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                            const int64_t* image_ids, const float* inj_samples,
                            const int64_t* inj_off, const mehhua_buffers_t* out, void* workspace,
                            size_t workspace_bytes, void* stream);

/* K3c: segmented means keyed by (object, level, class) and the bottom-up class -> level ->
 * object aggregation.  Replaces Lambda_L2.py:526-536 and AggregateObjScaleUnc (:597-619). */
int mehhua_k3_hua(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                  const mehhua_buffers_t* out, void* workspace, size_t workspace_bytes,
                  void* stream);

/* The whole per-batch path K1 -> K3a -> K3b -> K2 -> K3c on one stream, no host sync.
 * Replaces the Entropy_NMS route of _get_bboxes (Lambda_L2.py:254-384). */
int mehhua_score_batch(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                       const float* img_shapes, const float* scale_factors,
                       const int64_t* image_ids, const mehhua_buffers_t* out, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Entropy_ALL route (cfg->mode = MEHHUA_MODE_ALL; replaces ComputeScaleUnc + AggregateScaleUnc,
 * Lambda_L2.py:539-569, 636-691 and the nms_pre = -1 branch of _get_bboxes, :281-283).
 * mehhua_all_fg_rows: stream the logits once, collect every prior with max foreground softmax >
 * fg_thr, ordered by (level, prior index); writes for them score_rows (= softmax p) [B, pair_cap, C_out],
 * lam_rows / topk_idx (= prior index) / row_max / row_argmax [B, pair_cap], level_fg, pair_* (rows ==
 * pairs, object 0), pair_off, lam_mean (= mean lambda over ALL priors of the level), n_obj (0/1).
 * mehhua_score_batch_all: that + K2 + K3c with agg_object = SUM, agg_scale / agg_class = the
 * 'scaleX_classY' type; image_scores [B]. */
int mehhua_all_fg_rows(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                       const mehhua_buffers_t* out, void* workspace, size_t workspace_bytes, void* stream);
int mehhua_score_batch_all(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B,
                           const int64_t* image_ids, const mehhua_buffers_t* out, void* workspace,
                           size_t workspace_bytes, void* stream);

/* Optional live timing of mehhua_score_batch's stages with CUDA events recorded on the call's own
 * stream (used by bench.py for the roofline numbers).  _begin arms it for up to max_calls calls;
 * _end blocks, writes the summed milliseconds of the 7 stages
 * {K1a keys, K1b select, K1c gather, K3a nms, K3b pairs, K2 sampler, K3c hua} and the number of
 * recorded calls, and disarms it.  Not thread-safe; one timed stream at a time. */
#define MEHHUA_NUM_STAGES 7
int mehhua_stage_timing_begin(int32_t max_calls);
int mehhua_stage_timing_end(double* ms_sum, int32_t* calls_out);

/* K4: indices of the k largest scores among candidates (mask[i] != 0; mask may be NULL), written
 * in descending score order (ties: larger index first, the order a stable ascending argsort
 * followed by [-k:] would keep).  Replaces the arg[-nonZeroSize:] part of update_X_L
 * (utils/active_datasets.py:106-107, 124).  n_selected_out (device int32) = min(k, #candidates). */
size_t mehhua_pool_topk_workspace_bytes(int64_t n);               /* enough for any k <= n */
size_t mehhua_pool_topk_workspace_bytes_k(int64_t n, int32_t k);  /* enough for this k: pools >= 32 768 take the grid-wide
                                                                    * form (state + histograms + two (k + 4096)-element buffers)
                                                                    * when the workspace is at least this large, one block otherwise */
int    mehhua_k4_pool_topk(const float* scores, const uint8_t* mask, int64_t n, int32_t k,
                           int64_t* idx_out, int32_t* n_selected_out, void* workspace,
                           size_t workspace_bytes, void* stream);

/* KM: mutual-information score of the ensemble / MC-dropout baselines.  Replaces ComputeMI
 * (apis/CalEnsembleUnc.py:166-181, three members) and ComputeMCDropoutMI (apis/CalMCDropoutUnc.py:185-201,
 * n stochastic passes): p_m = sigmoid(logits_m), avg = mean_m p_m, total = -sum_c avg ln avg,
 * ale = mean_m(-sum_c p_m ln p_m), level value = mean over the level's priors of (total - ale),
 * image score = mean over levels.  member_logits: HOST array of n_members * num_levels DEVICE pointers,
 * [m * num_levels + s] -> member m's level-s output [B, A*n_cls, H, W] (channel = a*n_cls + c);
 * n_members <= 32; only H, W, A of `levels` are read.  level_mi: device [B, num_levels] or NULL (the
 * reference's `buffer`); image_scores: device [B]. */
size_t mehhua_mi_workspace_bytes(const mehhua_level_t* levels, int32_t num_levels, int32_t B);
int    mehhua_mi_score_batch(const float* const* member_logits, int32_t n_members, const mehhua_level_t* levels,
                             int32_t num_levels, int32_t n_cls, int32_t B, float* level_mi, float* image_scores,
                             void* workspace, size_t workspace_bytes, void* stream);

/* Diagnostic: rows parked per (image, level) by the last K1 call on this workspace (capture mode of K1, see
 * csrc/k1_alpha_topk.cuh), -1 for levels that are not in capture mode.  counts_out: host int32[B * num_levels]. */
int mehhua_debug_capture_counts(const mehhua_config_t* cfg, const mehhua_level_t* levels, int32_t B, void* workspace,
                                void* stream, int32_t* counts_out);

/* Known-answer hook for the Philox4x32 block of K2 (tests): ctr[4], key[2] -> out[8], host arrays:
 * out[0..3] = the 10-round block, out[4..7] = the 7-round block (the one K2's sampler uses). */
int mehhua_debug_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[8]);

/* Host-buffer entry point: same as mehhua_score_batch but every pointer in `levels`, img_shapes,
 * scale_factors, image_ids and image_scores_host is HOST memory.  The call copies the inputs to
 * the device, scores them and copies the B image scores back; it blocks until image_scores_host is
 * valid.  Handle = resident device buffers for one geometry, a copy stream and a compute stream: the
 * batch is uploaded in up to four image chunks and chunk j is scored while chunk j+1 is on the wire.
 * The copies are issued STRAIGHT FROM THE CALLER'S BUFFERS (no staging copy inside the library), so
 * that overlap - and the full PCIe rate - needs page-locked memory on the caller's side: torch's
 * pin_memory(), cudaHostAlloc, or mehhua_host_pin() below on an existing allocation.  Pageable
 * buffers are accepted and give the same results, but the CUDA driver then stages every copy through
 * its own bounce buffer synchronously (about half the rate, no overlap). */
typedef struct mehhua_host_ctx mehhua_host_ctx_t;
int  mehhua_host_ctx_create(const mehhua_config_t* cfg, const mehhua_level_t* level_shapes,
                            int32_t max_batch, mehhua_host_ctx_t** ctx_out);
void mehhua_host_ctx_destroy(mehhua_host_ctx_t* ctx);
int  mehhua_score_batch_host(mehhua_host_ctx_t* ctx, const mehhua_level_t* levels_host, int32_t B,
                             const float* img_shapes_host, const float* scale_factors_host,
                             const int64_t* image_ids_host, float* image_scores_host,
                             uint32_t* status_out);

/* Page-lock / unlock an existing host allocation (cudaHostRegister / cudaHostUnregister) so that the copies
 * of mehhua_score_batch_host run asynchronously from it - for callers that have no CUDA binding of their own
 * (e.g. numpy arrays through ctypes).  mehhua_host_is_pinned: 1 = page-locked, 0 = pageable, < 0 = error. */
int  mehhua_host_pin(void* ptr, size_t bytes);
int  mehhua_host_unpin(void* ptr);
int  mehhua_host_is_pinned(const void* ptr);

#ifdef __cplusplus
}
#endif
#endif  /* MEHHUA_H_ */
