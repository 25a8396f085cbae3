"""Drop-in boundary: the reference's head / API methods for the scoring path, backed by the CUDA
kernels.  Same names, argument meaning and return types as the reference so that
tools/train_RetinaNet.py / tools/train_SSD.py keep working unchanged:

  B200ScoringMixin._get_bboxes          <- Lambda_L2Net._get_bboxes   (Lambda_L2.py:254-384)
                                           MyLSSDHead._get_bboxes     (My_L_ssd_head.py:315-433)
  B200ScoringMixin.ComputeObjUnc        <- Lambda_L2.py:489-537 / My_L_ssd_head.py:435-482
  B200ScoringMixin.AggregateObjScaleUnc <- Lambda_L2.py:597-619 / My_L_ssd_head.py:517-539
  calculate_uncertainty                 <- mmdet/apis/test.py:65-70, 90-135
  update_X_L                            <- mmdet/utils/active_datasets.py:102-135 (pool.py)

The mixin goes in front of the reference head in the MRO (`class Lambda_L2Net_B200(
B200ScoringMixin, Lambda_L2Net)`, see register_heads / INTEGRATION.md).  Taken over: the Entropy_NMS
and Entropy_ALL scoring routes and the plain evaluation route (isEval=True -> det_results from
K1 + K3a); everything else (Entropy_NoNMS, ONNX export, a non-default cfg) falls through to the
reference method via super().
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .pool import ResumeCycle, ResumeCycle_WorkDir, save_cycle, scoring_stage, update_X_L  # noqa: F401  (re-exported)
from .scoring import Scorer
from .specs import HEAD_RETINA, HEAD_SSD, DetectorSpec, ScoringParams, parse_agg_spec


def _spec_from_head(head, featmaps: Sequence[Tuple[int, int]], num_anchors: Sequence[int], img_hw) -> DetectorSpec:
    act = getattr(head, "last_activation", "relu")
    kind = HEAD_SSD if act == "softmax" else HEAD_RETINA
    c_out = int(head.cls_out_channels)
    cfg = head.test_cfg
    stds = tuple(float(v) for v in getattr(head.bbox_coder, "stds", (1.0, 1.0, 1.0, 1.0)))
    means = tuple(float(v) for v in getattr(head.bbox_coder, "means", (0.0, 0.0, 0.0, 0.0)))
    if any(m != 0.0 for m in means):
        raise NotImplementedError("non-zero bbox_coder target_means")
    nms = cfg["nms"] if isinstance(cfg, dict) else cfg.nms
    return DetectorSpec(
        name=f"head_{kind}_{c_out}_{tuple(featmaps)}", head=kind,
        num_classes=c_out - (1 if kind == HEAD_SSD else 0), img_hw=tuple(img_hw),
        strides=tuple(range(len(featmaps))), featmaps=tuple(tuple(f) for f in featmaps),
        num_anchors=tuple(num_anchors), target_stds=stds, score_thr=float(cfg.get("score_thr")),
        max_per_img=int(cfg.get("max_per_img")), nms_pre=int(cfg.get("nms_pre", -1)),
        nms_iou=float(nms.get("iou_threshold", 0.5)))


class B200ScoringMixin:
    """Put in front of Lambda_L2Net / MyLSSDHead.  Head attributes used: cls_out_channels,
    last_activation, test_cfg, bbox_coder.{means,stds} - exactly what the reference method reads."""

    mehhua_params = ScoringParams()     # reference constants; override per class / instance for ablations
    mehhua_max_batch = 8
    # Lambda_L2_ReLU / _ablation read the object / foreground thresholds from kwargs['score_thr'] and
    # the cluster IoU from kwargs['iou_thr'] (Lambda_L2_ReLU.py:150-154, 395-398); Lambda_L2Net and
    # MyLSSDHead ignore those kwargs and hard-code 0.3 / 0.5 (Lambda_L2.py:349, 500, 508)
    mehhua_thresholds_from_kwargs = False
    mehhua_fused_eval = True            # isEval route: detections from K1 + K3a instead of super()
    mehhua_all_pair_cap = None          # Entropy_ALL / Entropy_Avg: foreground priors per image the row buffers hold
                                        # (None = min(N, 32768); raise it up to N for heads with very many of them)
    _mehhua_scorers: Dict[tuple, Scorer]

    def _mehhua_scorer(self, cls_scores: List[torch.Tensor], img_hw, uPool2: str, clsW: bool, kwargs=None) -> Scorer:
        B = cls_scores[0].shape[0]
        p = self.mehhua_params
        if self.mehhua_thresholds_from_kwargs and kwargs:
            thr = kwargs.get("score_thr") or 0.3
            iou = kwargs.get("iou_thr") or 0.5
            p = ScoringParams(n_samples=p.n_samples, fg_thr=thr, obj_thr=thr, cluster_iou=iou, lambda_scale=p.lambda_scale,
                              lambda_eps=p.lambda_eps, use_lambda=p.use_lambda, seed=p.seed, activation=p.activation)
        featmaps = [tuple(c.shape[-2:]) for c in cls_scores]
        c_out = int(self.cls_out_channels)
        num_anchors = [c.shape[1] // c_out for c in cls_scores]
        device = cls_scores[0].device
        key = (tuple(featmaps), tuple(num_anchors), c_out, str(device), uPool2, bool(clsW),
               max(B, self.mehhua_max_batch), p.fg_thr, p.obj_thr, p.cluster_iou, p.activation)
        cache = self.__dict__.setdefault("_mehhua_scorers", {})
        if key not in cache:
            spec = _spec_from_head(self, featmaps, num_anchors, img_hw)
            params = ScoringParams(n_samples=p.n_samples, fg_thr=p.fg_thr, obj_thr=p.obj_thr,
                                   cluster_iou=p.cluster_iou, lambda_scale=p.lambda_scale,
                                   lambda_eps=p.lambda_eps, use_lambda=p.use_lambda, agg=uPool2,
                                   cls_w=bool(clsW), seed=p.seed, activation=p.activation)
            parse_agg_spec(uPool2)       # KeyError for specs without object/scale/class, as the reference
            cache[key] = Scorer(spec, params, max_batch=max(B, self.mehhua_max_batch), device=device)
        return cache[key]

    # ------------------------------------------------------------------ widest boundary
    def _get_bboxes(self, mlvl_cls_scores, mlvl_bbox_preds, mlvl_anchors, img_shapes, scale_factors, cfg,
                    rescale=False, with_nms=True, **kwargs):
        scoring = bool(kwargs.get("isUnc")) and kwargs.get("uPool") == "Entropy_NMS" and with_nms \
            and "L_scores" in kwargs and not torch.onnx.is_in_onnx_export()
        if bool(kwargs.get("isUnc")) and "L_scores" in kwargs and (cfg is None or cfg is self.test_cfg) and (
                (kwargs.get("uPool") == "Entropy_ALL" and not with_nms) or
                (kwargs.get("uPool") == "Entropy_Avg" and self.mehhua_thresholds_from_kwargs)):
            return self._mehhua_entropy_all(mlvl_cls_scores, mlvl_bbox_preds, mlvl_anchors, img_shapes,
                                            scale_factors, **kwargs)
        if (not kwargs.get("isUnc")) and with_nms and self.mehhua_fused_eval and (cfg is None or cfg is self.test_cfg) \
                and not torch.onnx.is_in_onnx_export() and mlvl_cls_scores[0].is_cuda:
            # evaluation route (isEval=True, isUnc=None; eval hook -> single_gpu_test): plain det_results
            hw = tuple(int(v) for v in img_shapes[0][:2])
            sc = self._mehhua_scorer(list(mlvl_cls_scores), hw, ScoringParams().agg, False)
            if bool(sc.cfg.rescale) != bool(rescale):
                sc.cfg.rescale = int(bool(rescale))
            return sc.detect(mlvl_cls_scores, mlvl_bbox_preds, mlvl_anchors, img_shapes, scale_factors)
        if not scoring or (cfg is not None and cfg is not self.test_cfg):
            return super()._get_bboxes(mlvl_cls_scores, mlvl_bbox_preds, mlvl_anchors, img_shapes, scale_factors,
                                       cfg, rescale, with_nms, **kwargs)
        if kwargs.get("scaleUnc"):
            raise NotImplementedError("scaleUnc=True is undefined on the Entropy_NMS route of the reference")
        B = mlvl_cls_scores[0].shape[0]
        hw = tuple(int(v) for v in img_shapes[0][:2])
        sc = self._mehhua_scorer(list(mlvl_cls_scores), hw, kwargs["uPool2"], kwargs.get("clsW", False), kwargs)
        if bool(sc.cfg.rescale) != bool(rescale):
            sc.cfg.rescale = int(bool(rescale))
        ids = kwargs.get("image_ids")
        if ids is None and "batchIdx" in kwargs:
            ids = [int(kwargs["batchIdx"]) * B + j for j in range(B)]    # apis/test.py:115 batchIdx=i
        sc.save_max_conf(bool(kwargs.get("saveMaxConf")))
        res = sc.score(mlvl_cls_scores, mlvl_bbox_preds, kwargs["L_scores"], mlvl_anchors, img_shapes,
                       scale_factors, image_ids=ids)
        n_det = res.n_det.cpu().tolist()
        det_results = [(res.dets[b, :n_det[b]].clone(), res.det_labels[b, :n_det[b]].long())
                       for b in range(B)]
        agged = [float(v) if v != 0 else 0 for v in res.image_scores.cpu().tolist()]
        if kwargs.get("saveMaxConf"):      # getMaxConf, fused into the logits pass (K1a)
            return det_results, agged, res.level_maxconf.max(dim=1)[0].tolist()
        return det_results, agged

    def _mehhua_entropy_all(self, mlvl_cls_scores, mlvl_bbox_preds, mlvl_anchors, img_shapes, scale_factors,
                            **kwargs):
        """Entropy_ALL route (Lambda_L2.py:281-283, 362-365 -> ComputeScaleUnc + AggregateScaleUnc) and, for the
        ablation heads, the Entropy_Avg route (Lambda_L2_ReLU.py:261-263 -> ComputeAvgUnc + AggregateAvgUnc):
        returns (det_results, AggedUnc) or, with scaleUnc=True on Entropy_ALL, (det_results, AggedUnc, scaleUnc)
        where scaleUnc[i][s] = {str(cls): (aleatoric, epistemic)} as ComputeScaleUnc builds it (Lambda_L2.py:377-378).
        The reference's det_results on the Entropy_ALL route are the raw (boxes [N,4], scores [N,C+1]) of every
        prior, which no caller reads (apis/test.py:116 only takes len()); they are returned here with zero rows."""
        avg_mode = kwargs.get("uPool") == "Entropy_Avg"
        if kwargs.get("scaleUnc") and avg_mode:
            raise NotImplementedError("scaleUnc=True is undefined on the Entropy_Avg route of the reference")
        B = mlvl_cls_scores[0].shape[0]
        dev = mlvl_cls_scores[0].device
        hw = tuple(int(v) for v in img_shapes[0][:2])
        featmaps = [tuple(c.shape[-2:]) for c in mlvl_cls_scores]
        c_out = int(self.cls_out_channels)
        num_anchors = [c.shape[1] // c_out for c in mlvl_cls_scores]
        agg = "Entropy_Avg" if avg_mode else kwargs["uPool2"]
        key = ("all", tuple(featmaps), tuple(num_anchors), c_out, str(dev), agg, max(B, self.mehhua_max_batch),
               self.mehhua_all_pair_cap)
        cache = self.__dict__.setdefault("_mehhua_scorers", {})
        if key not in cache:
            spec = _spec_from_head(self, featmaps, num_anchors, hw)
            p = self.mehhua_params
            if avg_mode:     # ComputeAvgUnc: alpha = relu(logits) * lambda', T = 50, thresholds hard-coded (Lambda_L2_ReLU.py:453-466)
                params = ScoringParams(n_samples=50, fg_thr=0.3, lambda_scale=p.lambda_scale, lambda_eps=p.lambda_eps,
                                       use_lambda=True, agg=agg, seed=p.seed, activation="relu")
            else:
                params = ScoringParams(n_samples=p.n_samples, fg_thr=p.fg_thr, lambda_scale=p.lambda_scale,
                                       lambda_eps=p.lambda_eps, use_lambda=p.use_lambda, agg=agg, seed=p.seed)
            cache[key] = Scorer(spec, params, max_batch=max(B, self.mehhua_max_batch), device=dev, mode="all",
                                pair_cap=self.mehhua_all_pair_cap)
        sc = cache[key]
        ids = kwargs.get("image_ids")
        if ids is None and "batchIdx" in kwargs:
            ids = [int(kwargs["batchIdx"]) * B + j for j in range(B)]
        sc.save_max_conf(bool(kwargs.get("saveMaxConf")))
        res = sc.score(mlvl_cls_scores, mlvl_bbox_preds, kwargs["L_scores"], mlvl_anchors, img_shapes,
                       scale_factors, image_ids=ids)
        dets = [(torch.zeros(0, 4, device=dev), torch.zeros(0, c_out + (0 if getattr(self, "last_activation", "relu") == "softmax" else 1), device=dev))
                for _ in range(B)]
        agged = [float(v) if v != 0 else 0 for v in res.image_scores.cpu().tolist()]
        if kwargs.get("scaleUnc"):
            g = res.group_unc.cpu()
            scale_unc = [[{f"{c}": (g[b, s, c, 1], g[b, s, c, 2]) for c in range(c_out) if g[b, s, c, 0] > 0}
                          for s in range(g.shape[1])] for b in range(B)]
            return dets, agged, scale_unc
        if kwargs.get("saveMaxConf"):
            return dets, agged, res.level_maxconf.max(dim=1)[0].tolist()
        return dets, agged

    # ------------------------------------------------------------------ narrowest boundary
    def ComputeObjUnc(self, mlvl_cls_scores, pos_bboxes, mlvl_scores, mlvl_Ls, mlvl_idces):
        """Same inputs / nested output as the reference: output[i][obj][s][str(cls)] = (ale, epi)
        0-d tensors.  Cluster masks and kept rows come from the caller (as in the reference); the
        Dirichlet sampling runs in K2."""
        S = len(mlvl_cls_scores)
        B = mlvl_cls_scores[0].shape[0]
        dev = mlvl_scores[0].device
        c_out = int(self.cls_out_channels)
        p = self.mehhua_params
        ssd = getattr(self, "last_activation", "relu") == "softmax"
        out = [[[{} for _ in range(S)] for _ in range(pos_bboxes[b].size(1))] for b in range(B)]
        ksz = [m.shape[1] for m in mlvl_scores]
        koff = np.concatenate([[0], np.cumsum(ksz)])
        rows_t = torch.cat(list(mlvl_scores), dim=1).float().contiguous()
        lam_t = torch.cat(list(mlvl_Ls), dim=1).float().contiguous()
        from .scoring import pair_uncertainty
        for s in range(S):
            x = mlvl_cls_scores[s]
            x = x.permute(0, 2, 3, 1).reshape(B, -1, c_out)
            conf = x.softmax(dim=2)
            conf = conf[..., :-1].max(dim=2)[0] if ssd else conf.max(dim=2)[0]
            level_fg = (conf > p.fg_thr).any(dim=1).cpu().tolist()
            for i in range(B):
                if not level_fg[i]:
                    continue
                rows = mlvl_scores[s][i]
                fgpos = pos_bboxes[i][koff[s]:koff[s + 1]] & (rows.max(dim=1)[0] > p.fg_thr)[:, None]
                nz = fgpos.nonzero()
                if nz.shape[0] == 0:
                    continue
                pidx, oidx = nz[:, 0], nz[:, 1]
                unc = pair_uncertainty(rows_t[i], lam_t[i], pidx + int(koff[s]), oidx, p, seed_ids=(i, s))
                pcls = rows[pidx].argmax(dim=1)
                for obj in oidx.unique():
                    om = oidx == obj
                    for c in pcls[om].unique():
                        m = om & (pcls == c)
                        out[i][obj][s][f"{c}"] = (unc[m, 1].mean(), unc[m, 2].mean())
        return out

    def AggregateObjScaleUnc(self, objScaleClsUnc, type, clsW=False, **kwargs):
        """Nested-dict input -> list of python floats (0 for images without objects).  The input is
        a host-side Python structure, so this compatibility method reduces it on the host; the fused
        route (_get_bboxes) aggregates on the GPU in K3c."""
        ops = parse_agg_spec(type)
        red = {0: lambda v: float(np.float32(np.sum(np.asarray(v, dtype=np.float32)))),
               1: lambda v: float(np.float32(np.mean(np.asarray(v, dtype=np.float32)))),
               2: lambda v: float(np.float32(np.max(np.asarray(v, dtype=np.float32))))}
        f_obj, f_scale, f_cls = red[ops[0]], red[ops[1]], red[ops[2]]
        output = []
        for img in objScaleClsUnc:
            per_obj, seen = [], set()
            for obj in img:
                per_lvl = []
                for lvl in obj:
                    vals = [float(epi) for (_, epi) in lvl.values()]
                    seen.update(lvl.keys())
                    if vals:
                        per_lvl.append(f_cls(vals))
                if per_lvl:
                    per_obj.append(f_scale(per_lvl))
            output.append(f_obj(per_obj) if per_obj else 0)
            if clsW:
                output[-1] *= len(seen)
        return output


def max_conf(scorer: Scorer):
    """getMaxConf (mmdet/utils/functions.py:467-476) of the batch the scorer last processed:
    (list of per-image maxima, Tensor[B, S] per level).  The maxima come out of the same pass over
    the logits as the top-k keys (K1a / KA1), not from a second softmax pass; the scorer must have
    had save_max_conf(True) set for that batch."""
    if not scorer.bufs.level_maxconf:
        raise _lib.MehhuaError("max_conf: save_max_conf(True) was not set on this scorer")
    out = scorer.result().level_maxconf
    return out.max(dim=-1)[0].tolist(), out.clone()


def calculate_uncertainty(cfg, model, data_loader, **kwargs):
    """Mirror of calculate_uncertainty -> Uncertainty_fns.Entropy_NMS -> single_gpu_uncertainty
    (apis/test.py:52-70, 90-135): loops the pool loader in order and returns the list of 0-d CPU
    tensors the AL scripts stack (tools/train_RetinaNet.py:242-245).  Unlike the reference it does
    not swallow a failing batch (apis/test.py:122-128 would silently misalign image indices)."""
    if cfg.uncertainty_pool == "Random":     # Uncertainty_fns.Random (apis/test.py:20-25)
        return torch.randperm(len(data_loader.dataset)).numpy()
    if cfg.uncertainty_pool not in ("Entropy_NMS", "Entropy_ALL", "Entropy_Avg"):     # apis/test.py:27-38, 52-63: one loop for all
        raise NotImplementedError(f"uncertainty_pool={cfg.uncertainty_pool!r}: Entropy_NMS / Entropy_ALL / Entropy_Avg / Random")
    if "scaleUnc" not in kwargs:
        raise KeyError("scaleUnc")          # the reference reads kwargs['scaleUnc'] unconditionally (:129)
    model.eval()
    uncertainties, maxconfs = [], []
    seen = 0        # running image offset: the Philox streams are keyed by the image's position in the pool, whatever
                    # the batch sizes (a short last batch must not reuse the ids of earlier images)
    with torch.no_grad():
        for i, data in enumerate(data_loader):
            data = dict(data)
            data["img"] = getattr(data["img"], "data", data["img"])
            data["img_metas"] = getattr(data["img_metas"], "data", data["img_metas"])
            metas = data["img_metas"][0] if isinstance(data["img_metas"], (list, tuple)) else data["img_metas"]
            nb = len(metas)
            extra = {} if "image_ids" in kwargs else {"image_ids": list(range(seen, seen + nb))}
            result, *unc = model(return_loss=False, rescale=True, isEval=False, batchIdx=i, isUnc=cfg.uncertainty_type,
                                 uPool=cfg.uncertainty_pool, uPool2=cfg.uncertainty_pool2, **data, **kwargs, **extra)
            seen += nb
            others = unc[1:]
            unc = unc[0]
            while isinstance(unc[0], list):
                unc = unc[0]
            if len(unc) != len(result):
                raise _lib.MehhuaError(f"batch {i}: {len(unc)} scores for {len(result)} images")
            uncertainties.extend(unc)
            if kwargs.get("saveMaxConf"):
                maxconfs.extend(others[0])
    out = torch.tensor(uncertainties)
    if kwargs.get("saveMaxConf"):
        return [out, maxconfs]
    return [x.cpu() for x in out]


# The reference's scoring heads and how their scoring methods relate to Lambda_L2Net's (compared by
# AST: `_get_bboxes`, `ComputeObjUnc`, `AggregateObjScaleUnc`, `ComputeScaleUnc`, `AggregateScaleUnc`):
#   identical text   : Lambda_L1Net (Lambda_L1.py), Lambda_MSLENet (Lambda_MSLE.py),
#                      Lambda_L2Net_reverse (Lambda_L2_reverseorder.py)
#   kwargs thresholds: Lambda_L2Net_ablation (Lambda_L2_ablation.py:  score_thr / iou_thr from kwargs, lambda' kept)
#   kwargs thresholds, alpha = score row (no lambda'): Lambda_L2Net_NoL (Lambda_L2_noL.py), Lambda_L2Net_ReLU
HEAD_VARIANTS = {
    # registered name: (module, class, thresholds from kwargs, use_lambda[, activation])
    "Lambda_L2Net_B200": ("Lambda_L2", "Lambda_L2Net", False, True),
    "Lambda_L1Net_B200": ("Lambda_L1", "Lambda_L1Net", False, True),
    "Lambda_MSLENet_B200": ("Lambda_MSLE", "Lambda_MSLENet", False, True),
    "Lambda_L2Net_reverse_B200": ("Lambda_L2_reverseorder", "Lambda_L2Net_reverse", False, True),
    "Lambda_L2Net_ablation_B200": ("Lambda_L2_ablation", "Lambda_L2Net_ablation", True, True),
    "Lambda_L2Net_NoL_B200": ("Lambda_L2_noL", "Lambda_L2Net_NoL", True, False),
    "Lambda_L2Net_ReLU_B200": ("Lambda_L2_ReLU", "Lambda_L2Net_ReLU", True, False),
    "MyLSSDHead_B200": ("My_L_ssd_head", "MyLSSDHead", False, True),
    # the base head with last_activation='relu': no scoring route of its own, its detection route uses the
    # evidential scores alpha = relu(logits) + 1 (L_anchor_head.py:401-406)
    "L_AnchorHead_B200": ("L_anchor_head", "L_AnchorHead", False, True, "relu_plus_one"),
}


def make_variant(name: str, base: type) -> type:
    """The B200 drop-in class for one of the reference's scoring heads (`base` = the reference class)."""
    _, _, from_kwargs, use_lambda, *rest = HEAD_VARIANTS[name]
    return type(name, (B200ScoringMixin, base), dict(
        mehhua_thresholds_from_kwargs=from_kwargs,
        mehhua_params=ScoringParams(use_lambda=use_lambda, activation=rest[0] if rest else "softmax")))


def register_heads():
    """Register the `*_B200` drop-in heads (HEAD_VARIANTS) with the reference's HEADS registry (needs
    the reference's mmdet + mmcv importable).  Select with `--bbox-head Lambda_L2Net_B200`
    (tools/train_RetinaNet.py:59,90) or `bbox_head=dict(type='Lambda_L2Net_B200', ...)`.
    Returns the registered classes by name."""
    import importlib
    from mmdet.models.builder import HEADS
    out = {}
    for name, (module, cls, *_rest) in HEAD_VARIANTS.items():
        base = getattr(importlib.import_module(f"mmdet.models.dense_heads.{module}"), cls)
        out[name] = HEADS.register_module()(make_variant(name, base))
    return out
