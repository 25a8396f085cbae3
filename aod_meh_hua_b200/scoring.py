"""Host side of the scoring path: owns the device buffers and drives the C-ABI kernels.

`Scorer` is what the drop-in head methods (dropin.py), bench.py and the tests call.  PyTorch is
used only for device memory, streams and (in pool.py) torch.distributed; every computation on the
path happens inside libmehhua.so.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .specs import DetectorSpec, ScoringParams, parse_agg_spec, parse_scale_agg


def make_config(spec: DetectorSpec, params: ScoringParams, pair_cap: int, rescale: bool = True,
                mode: str = "nms") -> _lib.Config:
    ao, asc, ac = parse_scale_agg(params.agg) if mode == "all" else parse_agg_spec(params.agg)
    cfg = _lib.Config()
    cfg.mode = _lib.MODE_ALL if mode == "all" else _lib.MODE_NMS
    cfg.head = spec.head
    cfg.c_out = spec.c_out
    cfg.num_levels = spec.num_levels
    cfg.nms_pre = spec.nms_pre
    cfg.score_thr = spec.score_thr
    cfg.nms_iou = spec.nms_iou
    cfg.max_per_img = spec.max_per_img
    cfg.fg_thr = params.fg_thr
    cfg.obj_thr = params.obj_thr
    cfg.cluster_iou = params.cluster_iou
    cfg.lambda_scale = params.lambda_scale
    cfg.lambda_eps = params.lambda_eps
    cfg.use_lambda = int(params.use_lambda)
    cfg.n_samples = params.n_samples
    cfg.agg_object, cfg.agg_scale, cfg.agg_class = ao, asc, ac
    cfg.cls_w = int(params.cls_w)
    cfg.means = (C.c_float * 4)(0.0, 0.0, 0.0, 0.0)
    cfg.stds = (C.c_float * 4)(*spec.target_stds)
    cfg.wh_ratio_clip = 16 / 1000
    cfg.rescale = int(rescale)
    cfg.pair_cap = pair_cap
    cfg.activation = params.activation_code
    cfg.seed = params.seed
    return cfg


@dataclasses.dataclass
class BatchResult:
    """Views (first B entries) of the scorer's device buffers after a call."""
    B: int
    score_rows: torch.Tensor
    lam_rows: torch.Tensor
    boxes: torch.Tensor
    topk_idx: torch.Tensor
    row_max: torch.Tensor
    row_argmax: torch.Tensor
    level_fg: torch.Tensor
    dets: torch.Tensor
    det_labels: torch.Tensor
    det_flat: torch.Tensor
    n_det: torch.Tensor
    n_obj: torch.Tensor
    pair_row: torch.Tensor
    pair_obj: torch.Tensor
    pair_cls: torch.Tensor
    pair_off: torch.Tensor
    lam_mean: torch.Tensor
    pair_unc: torch.Tensor
    image_scores: torch.Tensor
    level_maxconf: torch.Tensor
    group_unc: Optional[torch.Tensor] = None      # Entropy_ALL family only


class Scorer:
    """Device-resident buffers + kernel launches for one detector geometry."""

    def __init__(self, spec: DetectorSpec, params: Optional[ScoringParams] = None, max_batch: int = 8,
                 device="cuda:0", pair_cap: Optional[int] = None, rescale: bool = True, mode: str = "nms", lib=None):
        """mode 'nms' = the Entropy_NMS route (objects from NMS, HUA over object/scale/class);
        mode 'all' = the Entropy_ALL route (every foreground prior, params.agg is one of the four
        'scaleX_classY' types, row buffers hold up to pair_cap foreground priors per image).
        lib: an alternative build of the library (experiments)."""
        self.lib = lib or _lib.load()
        if not torch.cuda.is_available():
            raise _lib.MehhuaError("the MEH/HUA scoring path needs a CUDA device (sm_100a); none is visible")
        self.spec = spec
        self.params = params or ScoringParams()
        self.device = torch.device(device)
        self.max_batch = int(max_batch)
        self.mode = mode
        K = spec.k_tot
        if pair_cap is None:
            pair_cap = min(K * spec.max_per_img, 65536) if mode == "nms" else min(spec.num_priors, 32768)
        self.pair_cap = int(pair_cap)
        self.cfg = make_config(spec, self.params, self.pair_cap, rescale, mode)
        self._shape_levels = _lib.LevelArray()
        for s, ((h, w), a) in enumerate(zip(spec.featmaps, spec.num_anchors)):
            self._shape_levels[s].H, self._shape_levels[s].W, self._shape_levels[s].A = h, w, a
        assert self.lib.mehhua_rows_per_image(C.byref(self.cfg), self._shape_levels) == K
        if mode == "all":
            K = self.pair_cap            # rows == pairs
        B, S, Cc, D, P = self.max_batch, spec.num_levels, spec.c_out, spec.max_per_img, self.pair_cap
        f32 = dict(dtype=torch.float32, device=self.device)
        i32 = dict(dtype=torch.int32, device=self.device)
        self.t: Dict[str, torch.Tensor] = dict(
            score_rows=torch.zeros(B, K, Cc, **f32), lam_rows=torch.zeros(B, K, **f32),
            boxes=torch.zeros(B, K, 4, **f32), topk_idx=torch.zeros(B, K, **i32),
            row_max=torch.zeros(B, K, **f32), row_argmax=torch.zeros(B, K, **i32),
            level_fg=torch.zeros(B, S, **i32), dets=torch.zeros(B, D, 5, **f32),
            det_labels=torch.zeros(B, D, **i32), det_flat=torch.zeros(B, D, **i32),
            n_det=torch.zeros(B, **i32), n_obj=torch.zeros(B, **i32),
            pair_row=torch.zeros(B, P, **i32), pair_obj=torch.zeros(B, P, **i32),
            pair_cls=torch.zeros(B, P, **i32), pair_off=torch.zeros(B, S + 1, **i32),
            lam_mean=torch.zeros(B, S, **f32), pair_unc=torch.zeros(B, P, 3, **f32),
            image_scores=torch.zeros(B, **f32), level_maxconf=torch.zeros(B, S, **f32))
        if mode == "all":        # per (level, class): count, mean aleatoric, mean epistemic (the scaleUnc return item)
            self.t["group_unc"] = torch.zeros(B, S, Cc, 3, **f32)
        self.bufs = _lib.Buffers()
        for name in _lib.BUFFER_FIELDS:
            if name in self.t:
                setattr(self.bufs, name, self.t[name].data_ptr())
        self.bufs.level_maxconf = None          # optional output, off unless save_max_conf(True)
        self.bufs.pair_avg = None               # diagnostic output of K2, off unless save_pair_avg(True)
        self.ws_bytes = int(self.lib.mehhua_workspace_bytes(C.byref(self.cfg), self._shape_levels, B))
        if self.ws_bytes == 0:
            _lib.check(_lib.E_ARG, "mehhua_workspace_bytes")
        self.workspace = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self._levels = _lib.LevelArray()
        self._keep: list = []       # keeps the tensors of the current call alive
        self._B = 0
        self._img_shapes = None
        self._scale_factors = None
        self._ids = None

    def save_max_conf(self, on: bool = True) -> None:
        """Ask the logits pass to also produce getMaxConf's per-level maxima (utils/functions.py:467-476;
        the reference computes them only under saveMaxConf).  Off by default."""
        self.bufs.level_maxconf = self.t["level_maxconf"].data_ptr() if on else None

    def save_pair_avg(self, on: bool = True) -> None:
        """Ask K2 to also write mean_t x_c of every pair (`avg` of Lambda_L2.py:521) into
        `self.t['pair_avg']` [B, pair_cap, C_out] - a diagnostic for the per-class moment tests."""
        if on and "pair_avg" not in self.t:
            self.t["pair_avg"] = torch.zeros(self.max_batch, self.pair_cap, self.spec.c_out, dtype=torch.float32,
                                             device=self.device)
        self.bufs.pair_avg = self.t["pair_avg"].data_ptr() if on else None

    # ------------------------------------------------------------------ inputs
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _dev_f32(self, x, what: str) -> torch.Tensor:
        if not torch.is_tensor(x):
            x = torch.as_tensor(np.asarray(x, dtype=np.float32))
        x = x.to(device=self.device, dtype=torch.float32)
        if not x.is_contiguous():
            x = x.contiguous()
        return x

    def bind(self, cls_scores: Sequence[torch.Tensor], bbox_preds: Sequence[torch.Tensor],
             L_scores: Sequence[torch.Tensor], anchors: Sequence[torch.Tensor], img_shapes,
             scale_factors, image_ids=None) -> int:
        """Point the level table at this batch's tensors (no copies for contiguous fp32 CUDA
        tensors).  Returns the batch size."""
        sp = self.spec
        S = sp.num_levels
        if not (len(cls_scores) == len(bbox_preds) == len(L_scores) == len(anchors) == S):
            raise ValueError(f"expected {S} levels")
        B = int(cls_scores[0].shape[0])
        if B > self.max_batch:
            raise ValueError(f"batch {B} exceeds the scorer's max_batch {self.max_batch}")
        keep = []
        for s in range(S):
            (h, w), a = sp.featmaps[s], sp.num_anchors[s]
            cs = self._dev_f32(cls_scores[s], "cls_score")
            bp = self._dev_f32(bbox_preds[s], "bbox_pred")
            ls = self._dev_f32(L_scores[s], "L_score")
            an = self._dev_f32(anchors[s], "anchors")
            if tuple(cs.shape) != (B, a * sp.c_out, h, w):
                raise ValueError(f"level {s}: cls_score shape {tuple(cs.shape)} != {(B, a * sp.c_out, h, w)}")
            if tuple(bp.shape) != (B, a * 4, h, w):
                raise ValueError(f"level {s}: bbox_pred shape {tuple(bp.shape)} != {(B, a * 4, h, w)}")
            if tuple(ls.shape) != (B, a, h, w):
                raise ValueError(f"level {s}: L_score shape {tuple(ls.shape)} != {(B, a, h, w)}")
            if tuple(an.shape) != (h * w * a, 4):
                raise ValueError(f"level {s}: anchors shape {tuple(an.shape)} != {(h * w * a, 4)}")
            L = self._levels[s]
            L.logits, L.deltas, L.lam, L.anchors = cs.data_ptr(), bp.data_ptr(), ls.data_ptr(), an.data_ptr()
            L.H, L.W, L.A = h, w, a
            keep += [cs, bp, ls, an]
        shp = np.asarray([tuple(s)[:2] for s in img_shapes], dtype=np.float32).reshape(B, 2)
        sf = np.asarray([np.asarray(s, dtype=np.float32).reshape(4) for s in scale_factors], dtype=np.float32)
        self._img_shapes = torch.from_numpy(shp).to(self.device)
        self._scale_factors = torch.from_numpy(sf.reshape(B, 4)).to(self.device)
        if image_ids is None:
            self._ids = None
        elif torch.is_tensor(image_ids):
            self._ids = image_ids.to(device=self.device, dtype=torch.int64).contiguous()
        else:
            self._ids = torch.as_tensor(np.asarray(image_ids, dtype=np.int64)).to(self.device)
        keep += [self._img_shapes, self._scale_factors, self._ids]
        self._keep = keep
        self._B = B
        return B

    def bind_raw(self, level_ptrs: Sequence[Sequence[int]], B: int, img_shapes: torch.Tensor,
                 scale_factors: torch.Tensor, image_ids: Optional[torch.Tensor]) -> None:
        """Bind pre-validated device pointers (bench hot loop: no per-step tensor checks)."""
        for s, (lg, dl, lm, an) in enumerate(level_ptrs):
            L = self._levels[s]
            L.logits, L.deltas, L.lam, L.anchors = lg, dl, lm, an
            (L.H, L.W), L.A = self.spec.featmaps[s], self.spec.num_anchors[s]
        self._img_shapes, self._scale_factors, self._ids = img_shapes, scale_factors, image_ids
        self._B = B

    # ------------------------------------------------------------------ stages
    def _call(self, name: str, *args) -> None:
        """One C-ABI call with the scorer's GPU made current for its duration: the library launches on
        the current device, which need not be the one the scorer's tensors live on (rank r of a
        multi-GPU job, or Scorer(device='cuda:1') in a single process)."""
        with torch.cuda.device(self.device):
            _lib.check(getattr(self.lib, name)(*args), name)

    def _common(self):
        return (C.byref(self.cfg), self._levels, self._B)

    def _ws(self):
        return (self.workspace.data_ptr(), self.ws_bytes, self._stream())

    def k1(self) -> None:
        self._call("mehhua_k1_alpha_topk", *self._common(), self._img_shapes.data_ptr(),
                   self._scale_factors.data_ptr(), C.byref(self.bufs), *self._ws())

    def nms(self) -> None:
        self._call("mehhua_nms_objects", *self._common(), C.byref(self.bufs), *self._ws())

    def pairs(self) -> None:
        self._call("mehhua_iou_pairs", *self._common(), C.byref(self.bufs), *self._ws())

    def k2(self, inj_samples: Optional[torch.Tensor] = None, inj_off: Optional[torch.Tensor] = None) -> None:
        ids = self._ids.data_ptr() if self._ids is not None else None
        ip = inj_samples.data_ptr() if inj_samples is not None else None
        io = inj_off.data_ptr() if inj_off is not None else None
        self._call("mehhua_k2_dirichlet_epi", *self._common(), ids, ip, io, C.byref(self.bufs), *self._ws())

    def hua(self) -> None:
        self._call("mehhua_k3_hua", *self._common(), C.byref(self.bufs), *self._ws())

    def all_rows(self) -> None:
        self._call("mehhua_all_fg_rows", *self._common(), C.byref(self.bufs), *self._ws())

    def score_bound(self) -> None:
        """The whole path for the bound batch, one C call, no host sync."""
        ids = self._ids.data_ptr() if self._ids is not None else None
        if self.mode == "all":
            self._call("mehhua_score_batch_all", *self._common(), ids, C.byref(self.bufs), *self._ws())
            return
        self._call("mehhua_score_batch", *self._common(), self._img_shapes.data_ptr(),
                   self._scale_factors.data_ptr(), ids, C.byref(self.bufs), *self._ws())

    def read_status(self) -> int:
        st = C.c_uint32(0)
        self._call("mehhua_read_status", *self._common(), self.workspace.data_ptr(), self._stream(), C.byref(st))
        return int(st.value)

    def capture_counts(self) -> np.ndarray:
        """Rows parked per (image, level) by the last K1 call (capture mode of K1, csrc/k1_alpha_topk.cuh);
        -1 for levels that are not in capture mode.  int32 [B, S]."""
        out = np.zeros((self._B, self.spec.num_levels), dtype=np.int32)
        self._call("mehhua_debug_capture_counts", *self._common(), self.workspace.data_ptr(), self._stream(),
                   out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out

    def check_status(self) -> int:
        """Raise on data-dependent failures; returns the informational bits."""
        st = self.read_status()
        if st & _lib.ST_PAIR_OVERFLOW:
            what = "foreground priors" if self.mode == "all" else "(box, object) pairs"
            limit = self.spec.num_priors if self.mode == "all" else self.spec.k_tot * self.spec.max_per_img
            raise _lib.MehhuaError(f"an image produced more than pair_cap={self.pair_cap} {what}; re-create the "
                                   f"Scorer with a larger pair_cap (at most {limit} can occur for this geometry)")
        return st

    def detect(self, cls_scores, bbox_preds, anchors, img_shapes, scale_factors):
        """Detections only (K1 + K3a), the isEval=True route of _get_bboxes (Lambda_L2.py:383-384):
        list of (Tensor[n,5], LongTensor[n]).  No lambda map is needed; the logits tensor stands in
        for it (K1c only copies it into lam_rows, which this route ignores)."""
        self.bind(cls_scores, bbox_preds, [c[:, : c.shape[1] // self.spec.c_out] for c in cls_scores], anchors,
                  img_shapes, scale_factors)
        self.k1()
        self.nms()
        res = self.result()
        n_det = res.n_det.cpu().tolist()
        return [(res.dets[b, :n_det[b]].clone(), res.det_labels[b, :n_det[b]].long()) for b in range(res.B)]

    def result(self) -> BatchResult:
        B = self._B
        return BatchResult(B=B, **{k: v[:B] for k, v in self.t.items() if k != "pair_avg"})

    # ------------------------------------------------------------------ whole path
    def score(self, cls_scores, bbox_preds, L_scores, anchors, img_shapes, scale_factors, image_ids=None,
              check: bool = True) -> BatchResult:
        self.bind(cls_scores, bbox_preds, L_scores, anchors, img_shapes, scale_factors, image_ids)
        self.score_bound()
        if check:
            self.check_status()
        return self.result()


def pool_topk(scores: torch.Tensor, k: int, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Indices (int64, descending score, ties: larger index first) of the k largest scores among
    mask != 0.  K4; replaces arg[-k:] of update_X_L (utils/active_datasets.py:106-107, 124)."""
    lib = _lib.load()
    if not scores.is_cuda:
        raise _lib.MehhuaError("pool_topk needs the scores on a CUDA device; there is no CPU fallback")
    scores = scores.contiguous().float()
    n = scores.numel()
    k = int(min(max(k, 0), n))
    out = torch.empty(max(k, 1), dtype=torch.int64, device=scores.device)
    nsel = torch.zeros(1, dtype=torch.int32, device=scores.device)
    ws_bytes = int(lib.mehhua_pool_topk_workspace_bytes_k(n, k))      # 256 for small pools, the grid-wide form's buffers for large ones
    ws = (torch.zeros if ws_bytes <= 256 else torch.empty)(ws_bytes, dtype=torch.uint8, device=scores.device)
    mp = None
    if mask is not None:
        mask = mask.to(device=scores.device, dtype=torch.uint8).contiguous()
        mp = mask.data_ptr()
    st = torch.cuda.current_stream(scores.device).cuda_stream
    with torch.cuda.device(scores.device):       # the library launches on the current device
        _lib.check(lib.mehhua_k4_pool_topk(scores.data_ptr(), mp, n, k, out.data_ptr(), nsel.data_ptr(),
                                           ws.data_ptr(), ws_bytes, st), "mehhua_k4_pool_topk")
    return out[: int(nsel.item())]


def pair_uncertainty(rows: torch.Tensor, lam: torch.Tensor, pair_row: torch.Tensor, pair_obj: torch.Tensor,
                     params: ScoringParams, seed_ids=(0, 0), inj_samples: Optional[torch.Tensor] = None,
                     return_avg: bool = False, lib=None):
    """K2 on an explicit pair list of ONE (image, level): rows [K, C] scores, lam [K], pair_row /
    pair_obj [P] -> [P, 3] (total, aleatoric, epistemic).  Used by the ComputeObjUnc compatibility
    method, whose cluster masks come from the caller.  lambda' = mean(lam[pair_row]) / (lam + eps)
    * scale as in Lambda_L2.py:513-515.  return_avg: also the class means mean_t x_c [P, C]
    (Lambda_L2.py:521 `avg`).  lib: an alternative build of the library (experiments)."""
    lib = lib or _lib.load()
    if not rows.is_cuda:
        raise _lib.MehhuaError("pair_uncertainty needs CUDA tensors; there is no CPU fallback")
    dev = rows.device
    K, Cc = int(rows.shape[0]), int(rows.shape[1])
    P = int(pair_row.numel())
    cfg = _lib.Config()
    cfg.head, cfg.c_out, cfg.num_levels, cfg.nms_pre = 0, Cc, 1, -1
    cfg.score_thr, cfg.nms_iou, cfg.max_per_img = 0.05, 0.5, 1
    cfg.fg_thr, cfg.obj_thr, cfg.cluster_iou = params.fg_thr, params.obj_thr, params.cluster_iou
    cfg.lambda_scale, cfg.lambda_eps, cfg.use_lambda = params.lambda_scale, params.lambda_eps, int(params.use_lambda)
    cfg.n_samples = params.n_samples
    cfg.agg_object = cfg.agg_scale = cfg.agg_class = 0
    cfg.stds = (C.c_float * 4)(1, 1, 1, 1)
    cfg.wh_ratio_clip, cfg.rescale, cfg.pair_cap, cfg.seed = 16 / 1000, 0, max(P, 1), params.seed
    lv = _lib.LevelArray()
    lv[0].H, lv[0].W, lv[0].A = 1, K, 1
    rows = rows.float().contiguous()
    lam = lam.float().contiguous()
    prow = pair_row.to(device=dev, dtype=torch.int32).contiguous()
    pobj = pair_obj.to(device=dev, dtype=torch.int32).contiguous()
    poff = torch.tensor([0, P], dtype=torch.int32, device=dev)
    lmean = lam[pair_row.long()].mean().reshape(1).float() if P else torch.zeros(1, device=dev)
    unc = torch.zeros(max(P, 1), 3, device=dev)
    bufs = _lib.Buffers()
    bufs.score_rows, bufs.lam_rows, bufs.lam_mean = rows.data_ptr(), lam.data_ptr(), lmean.data_ptr()
    bufs.pair_row, bufs.pair_obj, bufs.pair_off, bufs.pair_unc = prow.data_ptr(), pobj.data_ptr(), poff.data_ptr(), unc.data_ptr()
    avg = None
    if return_avg:
        avg = torch.zeros(max(P, 1), Cc, device=dev)
        bufs.pair_avg = avg.data_ptr()
    ws_bytes = int(lib.mehhua_workspace_bytes(C.byref(cfg), lv, 1))
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev)
    ids = torch.tensor([int(seed_ids[0]) * 64 + int(seed_ids[1])], dtype=torch.int64, device=dev)
    ip = io = None
    keep = None
    if inj_samples is not None:
        keep = (inj_samples.to(dev).float().contiguous(), torch.zeros(1, dtype=torch.int64, device=dev))
        ip, io = keep[0].data_ptr(), keep[1].data_ptr()
    st = torch.cuda.current_stream(dev).cuda_stream
    if P:
        with torch.cuda.device(dev):             # the library launches on the current device
            _lib.check(lib.mehhua_k2_dirichlet_epi(C.byref(cfg), lv, 1, ids.data_ptr(), ip, io, C.byref(bufs),
                                                   ws.data_ptr(), ws_bytes, st), "mehhua_k2_dirichlet_epi")
        torch.cuda.current_stream(dev).synchronize()
    if return_avg:
        return unc[:P], avg[:P]
    return unc[:P]
