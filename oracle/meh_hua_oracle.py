"""CPU oracle for the MEH alpha -> Dirichlet epistemic uncertainty -> HUA -> pool top-k path.

TEST INFRASTRUCTURE ONLY.  This module restates, on torch-CPU / numpy, the algorithm of the
reference's scoring path so the CUDA kernels can be checked against it.  Only `tests/`,
`__graft_entry__.smoke()` and bench.py's `cpu_baseline` / `--impl reference` legs may import it;
the product package `aod_meh_hua_b200` never does (it fails loudly without its CUDA library).

Pinning: the reference carries NO tests, fixtures or golden vectors for this path (SURVEY 4,
8c) - "parity unpinned" by the reference's own suite.  The restatement is instead pinned against
outputs of the reference's own functions (`_get_bboxes`, `ComputeObjUnc`, `AggregateObjScaleUnc`,
`delta2bbox`, `bbox_overlaps`, `update_X_L`, anchor generators) executed unmodified in the build
container by AST extraction (`oracle/make_golden.py` -> `tests/golden/*.npz`), and against the
reference's three docstring known-answers (delta2bbox, AnchorGenerator).

Third-party arithmetic the reference calls but does not vendor (restated from the published
algorithms):
  * mmcv 1.3.8 `mmcv.ops.nms.batched_nms` (call sites core/post_processing/bbox_nms.py:2,84):
    class offset = label * (max coordinate + 1), greedy NMS in descending-score order,
    suppress IoU > thr, areas without +1.
  * ATen `_sample_dirichlet` behind torch.distributions.Dirichlet (Lambda_L2.py:519-520):
    the oracle calls the installed torch for it; samples can be captured / injected.

Reference line numbers below are relative to /root/reference/mmdet/.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch.distributions import Dirichlet

HEAD_RETINA, HEAD_SSD = 0, 1


# --------------------------------------------------------------------------------------------
# stage a2: logits -> score rows, per-level top-k      (models/dense_heads/Lambda_L2.py:264-297,
#                                                       My_L_ssd_head.py:325-351)
# --------------------------------------------------------------------------------------------
def flatten_level(x: torch.Tensor, k: int) -> torch.Tensor:
    """[B, A*k, H, W] -> [B, H*W*A, k]   (Lambda_L2.py:266-267, 278)."""
    b = x.shape[0]
    return x.permute(0, 2, 3, 1).reshape(b, -1, k)


def score_rows(cls_score: torch.Tensor, head: int, c_out: int, activation: str = "softmax") -> torch.Tensor:
    x = flatten_level(cls_score, c_out)
    if activation == "relu_plus_one":
        # the base head's evidential form, L_anchor_head.py:401-406 (gamma = 1: (1-gamma)*Smax is an exact 0)
        alphas = x.relu() + 1
        s = alphas.sum(dim=2, keepdim=True) + 1e-20
        return alphas / s
    p = x.softmax(dim=2)
    if head == HEAD_RETINA:
        # Lambda_L2.py:269-273 with gamma = 1: (1-gamma)*Smax contributes an exact 0
        s = p.sum(dim=2, keepdim=True) + 1e-20
        return p / (s + 1e-9)
    return p  # My_L_ssd_head.py:330


def topk_keys(scores: torch.Tensor, head: int) -> torch.Tensor:
    """Ranking key of a prior: max over foreground classes (Lambda_L2.py:286-289)."""
    if head == HEAD_RETINA:
        return scores.max(-1)[0]
    return scores[..., :-1].max(-1)[0]


def k_for_topk(k: int, size: int) -> int:
    """core/export/onnx_helper.py:61-78 outside ONNX export."""
    if k <= 0 or size <= 0:
        return -1
    return k if k < size else -1


# --------------------------------------------------------------------------------------------
# stage a3: delta2bbox                         (core/bbox/coder/delta_xywh_bbox_coder.py:205-267)
# --------------------------------------------------------------------------------------------
def delta2bbox(rois: torch.Tensor, deltas: torch.Tensor, stds, max_shape=None,
               means=(0.0, 0.0, 0.0, 0.0), wh_ratio_clip: float = 16 / 1000) -> torch.Tensor:
    """rois [..., N, 4], deltas [..., N, 4]; max_shape (H, W[, C]) or one per batch row."""
    means_t = deltas.new_tensor(means)
    stds_t = deltas.new_tensor(stds)
    d = deltas * stds_t + means_t
    dx, dy, dw, dh = d[..., 0], d[..., 1], d[..., 2], d[..., 3]
    px = (rois[..., 0] + rois[..., 2]) * 0.5
    py = (rois[..., 1] + rois[..., 3]) * 0.5
    pw = rois[..., 2] - rois[..., 0]
    ph = rois[..., 3] - rois[..., 1]
    dxw = pw * dx
    dyh = ph * dy
    r = float(np.abs(np.log(wh_ratio_clip)))
    dw = dw.clamp(min=-r, max=r)
    dh = dh.clamp(min=-r, max=r)
    gw = pw * dw.exp()
    gh = ph * dh.exp()
    gx = px + dxw
    gy = py + dyh
    box = torch.stack([gx - gw * 0.5, gy - gh * 0.5, gx + gw * 0.5, gy + gh * 0.5], dim=-1)
    if max_shape is not None:
        ms = box.new_tensor(max_shape)[..., :2]          # (H, W) or [B, 2]
        hi = torch.cat([ms, ms], dim=-1).flip(-1)        # (W, H, W, H)
        if hi.ndim == 2:
            hi = hi.unsqueeze(-2)
        zero = box.new_tensor(0)
        box = torch.where(box < zero, zero, box)
        box = torch.where(box > hi, hi, box)
    return box


# --------------------------------------------------------------------------------------------
# stage a6: IoU                               (core/bbox/iou_calculators/iou2d_calculator.py:206-252)
# --------------------------------------------------------------------------------------------
def bbox_overlaps(b1: torch.Tensor, b2: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """[m,4] x [n,4] -> IoU [m,n]; no +1, union floored at eps; empty inputs give empty output."""
    rows, cols = b1.shape[0], b2.shape[0]
    if rows * cols == 0:
        return b1.new_zeros((rows, cols))
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = torch.max(b1[:, None, :2], b2[None, :, :2])
    rb = torch.min(b1[:, None, 2:], b2[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    union = a1[:, None] + a2[None, :] - overlap
    union = torch.max(union, union.new_tensor([eps]))
    return overlap / union


# --------------------------------------------------------------------------------------------
# stage a5: multiclass NMS                    (core/post_processing/bbox_nms.py:34-93 + mmcv)
# --------------------------------------------------------------------------------------------
def greedy_nms(boxes: np.ndarray, scores: np.ndarray, thr: float, max_keep: int = -1) -> np.ndarray:
    """Published greedy NMS: visit in descending score, keep a box unless a previously kept box
    overlaps it with IoU > thr (areas (x2-x1)*(y2-y1), IoU = inter / (a_i + a_j - inter), fp32).
    Ties in score are visited in ascending index order (stable).  Early exit after max_keep keeps
    is exact because the caller truncates to the first max_keep anyway (bbox_nms.py:86-88)."""
    order = np.argsort(-scores, kind="stable")
    b = boxes[order].astype(np.float32)
    area = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])).astype(np.float32)
    alive = np.ones(len(order), dtype=bool)
    keep = []
    for i in range(len(order)):
        if not alive[i]:
            continue
        keep.append(order[i])
        if 0 < max_keep <= len(keep):
            break
        j = slice(i + 1, None)
        xx1 = np.maximum(b[i, 0], b[j, 0])
        yy1 = np.maximum(b[i, 1], b[j, 1])
        xx2 = np.minimum(b[i, 2], b[j, 2])
        yy2 = np.minimum(b[i, 3], b[j, 3])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = (w * h).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (area[i] + area[j] - inter)
        alive[j] &= ~(ovr > np.float32(thr))
    return np.asarray(keep, dtype=np.int64)


def multiclass_nms(boxes: torch.Tensor, scores: torch.Tensor, score_thr: float, iou_thr: float,
                   max_num: int):
    """boxes [K,4], scores [K, C+1] (last column = background, dropped).
    Returns dets [n,5], labels [n], keep (indices into the flattened K*C candidate grid after the
    score filter), cand_flat (flat K*C index of every candidate)."""
    num_classes = scores.shape[1] - 1
    sc = scores[:, :-1].reshape(-1)
    valid = (sc > score_thr).nonzero(as_tuple=False).squeeze(1)
    row = valid // num_classes
    lab = valid % num_classes
    cb = boxes[row]
    cs = sc[valid]
    if cb.numel() == 0:
        return torch.cat([cb, cs[:, None]], -1), lab, valid[:0], valid
    # mmcv batched_nms: boxes of class c are shifted by c * (max coordinate + 1)
    max_coord = cb.max()
    off = lab.to(cb) * (max_coord + torch.tensor(1).to(cb))
    shifted = cb + off[:, None]
    if cb.is_cuda:      # torch-eager-on-GPU baseline only: torchvision's stock CUDA NMS (same rule)
        from torchvision.ops import nms as tv_nms
        keep_t = tv_nms(shifted, cs, iou_thr)
    else:
        keep = greedy_nms(shifted.numpy(), cs.numpy(), iou_thr, max_num)
        keep_t = torch.from_numpy(keep)
    dets = torch.cat([cb[keep_t], cs[keep_t][:, None]], -1)
    if max_num > 0:
        dets, keep_t = dets[:max_num], keep_t[:max_num]
    return dets, lab[keep_t], keep_t, valid


# --------------------------------------------------------------------------------------------
# stage a2-a6 together: the part of _get_bboxes before ComputeObjUnc
#                                 (Lambda_L2.py:254-349, My_L_ssd_head.py:315-400)
# --------------------------------------------------------------------------------------------
def pre_stage(cls_scores: List[torch.Tensor], bbox_preds: List[torch.Tensor],
              L_scores: List[torch.Tensor], anchors: List[torch.Tensor], img_shapes, scale_factors,
              *, head: int, c_out: int, stds, nms_pre: int, score_thr: float, nms_iou: float,
              max_per_img: int, obj_thr: float = 0.3, cluster_iou: float = 0.5,
              rescale: bool = True, topk_override: Optional[List[torch.Tensor]] = None,
              activation: str = "softmax") -> Dict[str, object]:
    """topk_override (stage isolation): per-level [B, K_s] prior indices to use instead
    of this function's own topk result - lets a test feed a kernel's (near-tie re-ordered but
    set-identical) row order into the later oracle stages."""
    B = cls_scores[0].shape[0]
    lvl_scores, lvl_boxes, lvl_L, lvl_idx, lvl_keys = [], [], [], [], []
    for lv, (cls, reg, anc, lam) in enumerate(zip(cls_scores, bbox_preds, anchors, L_scores)):
        sc = score_rows(cls.float(), head, c_out, activation)
        lm = lam.permute(0, 2, 3, 1).reshape(B, -1)
        dl = flatten_level(reg.float(), 4)
        an = anc[None].expand_as(dl)
        n = dl.shape[1]
        k = k_for_topk(nms_pre, n)
        lvl_keys.append(topk_keys(sc, head))
        if k > 0:
            _, idx = topk_keys(sc, head).topk(k)
            if topk_override is not None:
                idx = topk_override[lv].long()
            bi = torch.arange(B, device=idx.device).view(-1, 1).expand_as(idx)
            an, dl, sc, lm = an[bi, idx, :], dl[bi, idx, :], sc[bi, idx, :], lm[bi, idx]
        else:
            idx = torch.arange(n, device=sc.device).view(1, -1).expand(B, n)
        lvl_boxes.append(delta2bbox(an, dl, stds, max_shape=img_shapes))
        lvl_scores.append(sc)
        lvl_L.append(lm)
        lvl_idx.append(idx)
    boxes = torch.cat(lvl_boxes, dim=1)
    if rescale:
        boxes = boxes / boxes.new_tensor(np.asarray(scale_factors, dtype=np.float32)).unsqueeze(1)
    scores = torch.cat(lvl_scores, dim=1)
    if head == HEAD_RETINA:
        scores_nms = torch.cat([scores, scores.new_zeros(B, scores.shape[1], 1)], dim=-1)
    else:
        scores_nms = scores
    dets, labels, keeps, pos = [], [], [], []
    for j in range(B):
        d, l, keep, cand = multiclass_nms(boxes[j], scores_nms[j], score_thr, nms_iou, max_per_img)
        dets.append(d)
        labels.append(l)
        keeps.append(cand[keep])                       # flat row*C + class index of each det
        objs = d[d[:, -1] > obj_thr][:, :4]            # GetObjectIdx, Lambda_L2.py:343-349
        pos.append(bbox_overlaps(boxes[j], objs) > cluster_iou)
    return dict(lvl_scores=lvl_scores, lvl_L=lvl_L, lvl_idx=lvl_idx, lvl_keys=lvl_keys, boxes=boxes, scores=scores,
                dets=dets, labels=labels, det_flat=keeps, pos_bboxes=pos)


# --------------------------------------------------------------------------------------------
# stage a8: ComputeObjUnc                    (Lambda_L2.py:489-537, My_L_ssd_head.py:435-482)
# --------------------------------------------------------------------------------------------
SampleFn = Callable[[torch.Tensor, int, int, int], torch.Tensor]   # (alpha[P,C], T, image, level)


def default_sampler(alpha: torch.Tensor, T: int, i: int, s: int) -> torch.Tensor:
    return Dirichlet(alpha).sample(torch.tensor([T]))


def uncertainty_from_samples(samples: torch.Tensor):
    """[T,P,C] -> (total, aleatoric, epistemic) per pair; Lambda_L2.py:521-525 (natural log)."""
    avg = samples.mean(dim=0)
    total = (-avg * avg.log()).sum(dim=1)
    ent = (-samples * samples.log()).sum(dim=-1)
    ale = ent.mean(dim=0)
    return total, ale, total - ale


def compute_obj_unc(cls_scores: List[torch.Tensor], pos_bboxes: List[torch.Tensor],
                    lvl_scores: List[torch.Tensor], lvl_L: List[torch.Tensor], *, head: int,
                    c_out: int, T: int = 500, fg_thr: float = 0.3, lambda_scale: float = 25.0,
                    lambda_eps: float = 1e-7, use_lambda: bool = True,
                    sampler: SampleFn = default_sampler, analytic: bool = False,
                    record_flat: Optional[bool] = None):
    """Returns (nested, flat): nested[i][obj][s][str(cls)] = (ale, epi) exactly as the reference
    builds it, and a flat per-(image, level) record list for stage-wise comparison.
    analytic=True replaces the T drawn samples by their T -> infinity limits (dirichlet_expectations):
    a deterministic stand-in for the Monte-Carlo step, used by the pool-level set-identity tests."""
    S = len(cls_scores)
    B = cls_scores[0].shape[0]
    n_obj = [p.size(1) for p in pos_bboxes]
    nested = [[[{} for _ in range(S)] for _ in range(n_obj[b])] for b in range(B)]
    flat = []
    level_fg = np.zeros((B, S), dtype=bool)
    start = 0
    for s in range(S):
        end = start + lvl_scores[s].size(1)          # StartEnd, utils/functions.py:438-444
        for i in range(B):
            x = cls_scores[s][i].permute(1, 2, 0).reshape(-1, c_out)
            p = x.softmax(dim=1)
            conf = p.max(dim=1)[0] if head == HEAD_RETINA else p[:, :-1].max(dim=1)[0]
            fg = conf > fg_thr
            level_fg[i, s] = bool(fg.any())
            if not level_fg[i, s]:
                continue
            rows = lvl_scores[s][i]
            pb = pos_bboxes[i][start:end]
            if len(pb.nonzero()) == 0:
                continue
            fgpos = pb & (rows.max(dim=1)[0] > fg_thr)[:, None].expand_as(pb)
            nz = fgpos.nonzero()
            pidx, oidx = nz[:, 0], nz[:, 1]
            if len(pidx) == 0:
                continue
            ps = rows[pidx]
            lam = lvl_L[s][i][pidx]
            lam_p = lam.mean() / (lam + lambda_eps) * lambda_scale
            alpha = ps * lam_p[:, None] if use_lambda else ps
            if analytic:
                h_, e_, _ = dirichlet_expectations(alpha.double().cpu().numpy())
                total = torch.from_numpy(h_.astype(np.float32)).to(alpha.device)
                ale = torch.from_numpy(e_.astype(np.float32)).to(alpha.device)
                epi = total - ale
            else:
                smp = sampler(alpha, T, i, s)
                total, ale, epi = uncertainty_from_samples(smp)
            pcls = ps.argmax(dim=1)
            for obj in oidx.unique():
                om = oidx == obj
                for c in ps[om].argmax(dim=1).unique():
                    m = om & (pcls == c)
                    nested[i][obj][s][f"{c}"] = (ale[m].mean(), epi[m].mean())
            # stage records for the parity tests (skipped by the eager-GPU baseline unless asked for)
            if record_flat or (record_flat is None and not ps.is_cuda):
                n_ = lambda t_: t_.detach().cpu().numpy()
                flat.append(dict(image=i, level=s, row=n_(pidx + start), obj=n_(oidx),
                                 cls=n_(pcls), lam_p=n_(lam_p), alpha=n_(alpha),
                                 total=n_(total), ale=n_(ale), epi=n_(epi)))
        start = end
    return nested, flat, level_fg


# --------------------------------------------------------------------------------------------
# stage a9/a10: HUA                           (utils/functions.py:425-436, Lambda_L2.py:597-619)
# --------------------------------------------------------------------------------------------
def extract_agg_func(spec: str):
    table = {"Sum": torch.sum, "Avg": torch.mean, "Max": torch.max}
    out = {}
    for name in ("object", "scale", "class"):
        for tok in spec.split("_"):
            if name in tok:
                out[name] = table[tok.replace(name, "")]
    return out


def aggregate_obj_scale_unc(nested, spec: str, cls_w: bool = False) -> List[float]:
    f = extract_agg_func(spec)
    out = []
    for img in nested:
        per_obj, seen = [], {}
        for obj in img:
            per_lvl = []
            for lvl in obj:
                vals = []
                for c, (ale, epi) in lvl.items():
                    vals.append(epi.item())
                    seen[c] = ""
                if vals:
                    per_lvl.append(f["class"](torch.tensor(vals)))
            if per_lvl:
                per_obj.append(f["scale"](torch.tensor(per_lvl)))
        out.append(f["object"](torch.tensor(per_obj)).item() if per_obj else 0)
        if cls_w:
            out[-1] *= len(seen)
    return out


# --------------------------------------------------------------------------------------------
# stage a13: Entropy_ALL mode - ComputeScaleUnc + AggregateScaleUnc
#                    (Lambda_L2.py:539-569, 636-691; My_L_ssd_head.py:484-515, 541-596)
# --------------------------------------------------------------------------------------------
def compute_scale_unc(cls_scores: List[torch.Tensor], L_scores: List[torch.Tensor], *, head: int, c_out: int,
                      T: int = 500, fg_thr: float = 0.3, lambda_scale: float = 25.0, lambda_eps: float = 1e-7,
                      use_lambda: bool = True, sampler: SampleFn = default_sampler):
    """Every prior with max softmax > fg_thr is sampled; no NMS / objects.  Returns (nested, flat):
    nested[i][s][str(cls)] = (ale, epi) as the reference builds it.  use_lambda=False is the
    Lambda_L2Net_NoL / Lambda_L2Net_ReLU form (alpha = softmax, Lambda_L2_noL.py ComputeScaleUnc)."""
    S = len(cls_scores)
    B = cls_scores[0].shape[0]
    nested = [[{} for _ in range(S)] for _ in range(B)]
    flat = []
    for s in range(S):
        for i in range(B):
            x = cls_scores[s][i].permute(1, 2, 0).reshape(-1, c_out)
            p = x.softmax(dim=1)
            conf = p.max(dim=1)[0] if head == HEAD_RETINA else p[:, :-1].max(dim=1)[0]
            fg = conf > fg_thr
            if not bool(fg.any()):
                continue
            lam = L_scores[s][i].permute(1, 2, 0).reshape(-1, 1)
            lam_p = lam.mean() / (lam + lambda_eps) * lambda_scale
            alpha = (p * lam_p)[fg] if use_lambda else p[fg]
            smp = sampler(alpha, T, i, s)
            total, ale, epi = uncertainty_from_samples(smp)
            pcls = alpha.argmax(dim=1)
            for c in pcls.unique():
                m = pcls == c
                nested[i][s][f"{c}"] = (ale[m].mean(), epi[m].mean())
            flat.append(dict(image=i, level=s, prior=fg.nonzero()[:, 0].numpy(), cls=pcls.numpy(),
                             alpha=alpha.numpy(), total=total.numpy(), ale=ale.numpy(), epi=epi.numpy()))
    return nested, flat


def aggregate_scale_unc(nested, kind: str):
    """The four hard-coded aggregation types of AggregateScaleUnc (numpy float64 means / sums of the
    group epistemic values); an unknown type yields an empty list, as in the reference."""
    table = {"scaleAvg_classAvg": (np.mean, np.mean), "scaleSum_classSum": (np.sum, np.sum),
             "scaleSum_classAvg": (np.sum, np.mean), "scaleAvg_classSum": (np.mean, np.sum)}
    if kind not in table:
        return []
    f_scale, f_cls = table[kind]
    out = []
    for img in nested:
        per_lvl = []
        for lvl in img:
            vals = [epi.item() for (_, (ale, epi)) in lvl.items()]
            if vals:
                per_lvl.append(f_cls(np.array(vals)))
        out.append(f_scale(np.array(per_lvl)) if per_lvl else 0)
    return out


# --------------------------------------------------------------------------------------------
# Entropy_Avg mode of the ablation heads - ComputeAvgUnc + AggregateAvgUnc
#                                                  (Lambda_L2_ReLU.py:446-474, 532-541)
# --------------------------------------------------------------------------------------------
def zero_alpha_sampler(alpha: torch.Tensor, T: int, i: int, s: int) -> torch.Tensor:
    """Dirichlet draws for rows that contain alpha == 0 (relu of a negative logit).  The reference was
    pinned to torch 1.5, which sampled such rows silently (a zero-shape gamma draw is 0 and the sample
    is clamped to FLT_MIN); later torch versions refuse them unless argument validation is off."""
    return Dirichlet(alpha, validate_args=False).sample(torch.tensor([T]))


def compute_avg_unc(cls_scores: List[torch.Tensor], L_scores: List[torch.Tensor], *, c_out: int, T: int = 50,
                    fg_thr: float = 0.3, lambda_scale: float = 25.0, lambda_eps: float = 1e-7,
                    sampler: SampleFn = zero_alpha_sampler):
    """r = relu(logits); prob = r / (sum r + 1e-9); every prior with max prob > fg_thr is sampled with
    alpha = r * lambda' (T = 50); output[i][s] = mean epistemic uncertainty of the level's foreground
    priors ({} when it has none).  Returns (nested, flat)."""
    S = len(cls_scores)
    B = cls_scores[0].shape[0]
    nested = [[{} for _ in range(S)] for _ in range(B)]
    flat = []
    for s in range(S):
        for i in range(B):
            x = cls_scores[s][i].permute(1, 2, 0).reshape(-1, c_out)
            r = F.relu(x)
            prob = r / (r.sum(dim=1, keepdim=True) + 1e-9)
            fg = prob.max(dim=1)[0] > fg_thr
            if not bool(fg.any()):
                continue
            lam = L_scores[s][i].permute(1, 2, 0).reshape(-1, 1)
            lam_p = lam.mean() / (lam + lambda_eps) * lambda_scale
            alpha = (r * lam_p)[fg]
            smp = sampler(alpha, T, i, s)
            total, ale, epi = uncertainty_from_samples(smp)
            nested[i][s] = epi.mean().item()
            flat.append(dict(image=i, level=s, prior=fg.nonzero()[:, 0].numpy(), cls=alpha.argmax(dim=1).numpy(),
                             alpha=alpha.numpy(), total=total.numpy(), ale=ale.numpy(), epi=epi.numpy()))
    return nested, flat


def aggregate_avg_unc(nested):
    """AggregateAvgUnc: the mean over the levels that hold a (truthy) value; NaN when none does."""
    out = []
    for img in nested:
        vals = [v for v in img if v]
        with np.errstate(invalid="ignore"), __import__("warnings").catch_warnings():
            __import__("warnings").simplefilter("ignore")
            out.append(np.array(vals).mean())
    return out


def score_batch_avg(batch: Dict[str, object], *, c_out: int, T: int = 50, fg_thr: float = 0.3,
                    lambda_scale: float = 25.0, lambda_eps: float = 1e-7, sampler: SampleFn = zero_alpha_sampler,
                    **_unused) -> Dict[str, object]:
    nested, flat = compute_avg_unc(batch["cls_scores"], batch["L_scores"], c_out=c_out, T=T, fg_thr=fg_thr,
                                   lambda_scale=lambda_scale, lambda_eps=lambda_eps, sampler=sampler)
    return dict(nested=nested, flat=flat, image_scores=aggregate_avg_unc(nested))


def score_batch_all(batch: Dict[str, object], *, head: int, c_out: int, T: int = 500, fg_thr: float = 0.3,
                    lambda_scale: float = 25.0, lambda_eps: float = 1e-7, kind: str = "scaleAvg_classAvg",
                    use_lambda: bool = True, sampler: SampleFn = default_sampler, **_unused) -> Dict[str, object]:
    nested, flat = compute_scale_unc(batch["cls_scores"], batch["L_scores"], head=head, c_out=c_out, T=T,
                                     fg_thr=fg_thr, lambda_scale=lambda_scale, lambda_eps=lambda_eps,
                                     use_lambda=use_lambda, sampler=sampler)
    return dict(nested=nested, flat=flat, image_scores=aggregate_scale_unc(nested, kind))


# --------------------------------------------------------------------------------------------
# the whole per-batch path (what _get_bboxes returns on the Entropy_NMS route)
# --------------------------------------------------------------------------------------------
def score_batch(batch: Dict[str, object], *, head: int, c_out: int, stds, nms_pre: int = 1000,
                score_thr: float = 0.05, nms_iou: float = 0.5, max_per_img: int = 100,
                T: int = 500, fg_thr: float = 0.3, obj_thr: float = 0.3, cluster_iou: float = 0.5,
                lambda_scale: float = 25.0, lambda_eps: float = 1e-7, use_lambda: bool = True,
                agg: str = "objectSum_scaleMax_classSum", cls_w: bool = False,
                sampler: SampleFn = default_sampler, rescale: bool = True,
                topk_override: Optional[List[torch.Tensor]] = None, analytic: bool = False,
                record_flat: Optional[bool] = None) -> Dict[str, object]:
    pre = pre_stage(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
                    batch["img_shapes"], batch["scale_factors"], head=head, c_out=c_out, stds=stds,
                    nms_pre=nms_pre, score_thr=score_thr, nms_iou=nms_iou, max_per_img=max_per_img,
                    obj_thr=obj_thr, cluster_iou=cluster_iou, rescale=rescale, topk_override=topk_override)
    nested, flat, level_fg = compute_obj_unc(
        batch["cls_scores"], pre["pos_bboxes"], pre["lvl_scores"], pre["lvl_L"], head=head,
        c_out=c_out, T=T, fg_thr=fg_thr, lambda_scale=lambda_scale, lambda_eps=lambda_eps,
        use_lambda=use_lambda, sampler=sampler, analytic=analytic, record_flat=record_flat)
    unc = aggregate_obj_scale_unc(nested, agg, cls_w)
    pre.update(nested=nested, flat=flat, level_fg=level_fg, image_scores=unc)
    return pre


def spec_kwargs(spec, params=None) -> Dict[str, object]:
    """kwargs of score_batch from a DetectorSpec (+ optional ScoringParams)."""
    kw = dict(head=spec.head, c_out=spec.c_out, stds=spec.target_stds, nms_pre=spec.nms_pre,
              score_thr=spec.score_thr, nms_iou=spec.nms_iou, max_per_img=spec.max_per_img)
    if params is not None:
        kw.update(T=params.n_samples, fg_thr=params.fg_thr, obj_thr=params.obj_thr,
                  cluster_iou=params.cluster_iou, lambda_scale=params.lambda_scale,
                  lambda_eps=params.lambda_eps, use_lambda=params.use_lambda, agg=params.agg,
                  cls_w=params.cls_w)
    return kw


# --------------------------------------------------------------------------------------------
# closed forms for moment matching of a free-running sampler (SURVEY 7 "Sampling parity")
# --------------------------------------------------------------------------------------------
def get_max_conf(cls_scores: List[torch.Tensor], n_cls: int):
    """getMaxConf (mmdet/utils/functions.py:467-476): per (image, level) the largest softmax
    probability over all priors and ALL classes (background included for SSD), and its maximum over
    levels.  Returns (list[float] per image, Tensor[B, S])."""
    B = cls_scores[0].size(0)
    out = torch.zeros(B, len(cls_scores), device=cls_scores[0].device)
    for s, x in enumerate(cls_scores):
        p = x.permute(0, 2, 3, 1).reshape(B, -1, n_cls).softmax(dim=-1)
        out[:, s] = p.reshape(B, -1).max(dim=-1)[0]
    return out.max(dim=-1)[0].tolist(), out


def dirichlet_expectations(alpha: np.ndarray):
    """alpha [P,C] (float64) -> (H(mean), E[entropy], epistemic_inf) per row:
    mean = alpha/alpha0; E[-sum x ln x] = psi(alpha0+1) - sum mean_c psi(alpha_c+1)."""
    from scipy.special import digamma
    a = np.asarray(alpha, dtype=np.float64)
    a0 = a.sum(axis=1, keepdims=True)
    m = a / a0
    with np.errstate(divide="ignore", invalid="ignore"):
        h = -(np.where(m > 0, m * np.log(m), 0.0)).sum(axis=1)
    e_ent = digamma(a0[:, 0] + 1) - (m * digamma(a + 1)).sum(axis=1)
    return h, e_ent, h - e_ent


# --------------------------------------------------------------------------------------------
# stage a12: pool selection                                  (utils/active_datasets.py:102-135)
# --------------------------------------------------------------------------------------------
def update_X_L(uncertainty, X_all, X_L, X_S_size, rng: Optional[np.random.RandomState] = None,
               **kwargs):
    """Restatement with the two host-RNG draws routed through `rng` (default: numpy's global
    state, as in the reference) so the deterministic top-k part can be compared on its own."""
    rs = np.random if rng is None else rng
    if torch.is_tensor(uncertainty):
        uncertainty = uncertainty.cpu().numpy()
    pool = np.array(list(set(X_all) - set(X_L)))
    u = uncertainty[pool]
    arg = u.argsort()
    if kwargs.get("zeroRate"):
        zeros = (u == 0).nonzero()[0]
        n_zero = int(X_S_size * kwargs["zeroRate"])
        n_top = X_S_size - n_zero
        n_zero = min(n_zero, len(zeros))
        mode = kwargs.get("useMaxConf", "False")
        if mode != "False":
            order = np.array(kwargs["maxconf"])[pool].argsort()
            zidx = order[:n_zero] if mode == "min" else order[-n_zero:]
        else:
            zidx = rs.choice(zeros, n_zero)
        picked = np.concatenate((pool[zidx], pool[arg[-n_top:]]))
    else:
        picked = pool[arg[-X_S_size:]]
    X_L_next = np.concatenate((X_L, picked))
    rest = np.array(list(set(X_all) - set(X_L_next)))
    rs.shuffle(rest)
    X_U_next = rest[:X_L_next.shape[0]]
    X_L_next.sort()
    X_U_next.sort()
    return X_L_next, X_U_next


def topk_part(uncertainty: np.ndarray, X_all, X_L, n_top: int) -> np.ndarray:
    """The deterministic part of update_X_L: ids of the n_top largest scores among unlabelled
    images (arg[-n_top:], active_datasets.py:107,124).  Returned as a sorted id array; defined
    only up to ties in the score."""
    pool = np.array(sorted(set(X_all) - set(X_L)))
    arg = np.argsort(uncertainty[pool], kind="stable")
    return np.sort(pool[arg[-n_top:]])


# --------------------------------------------------------------------------------------------
# guard-band report: how far every integer decision is from flipping (hard part 1)
# --------------------------------------------------------------------------------------------
def decision_margins(batch, out, *, head: int, c_out: int, nms_pre: int, score_thr: float,
                     fg_thr: float = 0.3, obj_thr: float = 0.3, cluster_iou: float = 0.5,
                     nms_iou: float = 0.5, **_unused) -> Dict[str, float]:
    """Smallest relative distance of any compared quantity to its threshold / neighbour.  A test
    may demand bit-exact integer outputs only on inputs whose margins exceed the fp32 noise of
    a re-ordered softmax (~1e-6)."""
    m = dict(topk_boundary=np.inf, topk_adjacent=np.inf, score_thr=np.inf, fg_thr=np.inf,
             level_fg=np.inf, obj_thr=np.inf, cluster_iou=np.inf, nms_iou=np.inf,
             nms_adjacent=np.inf, argmax=np.inf)

    def rel(a, b):
        return float(np.min(np.abs(a - b) / np.maximum(np.abs(b), 1e-30))) if np.size(a) else np.inf

    B = batch["cls_scores"][0].shape[0]
    for s, cls in enumerate(batch["cls_scores"]):
        sc = score_rows(cls.float(), head, c_out)
        keys = topk_keys(sc, head).numpy()
        p = flatten_level(cls.float(), c_out).softmax(dim=2)
        conf = (p.max(-1)[0] if head == HEAD_RETINA else p[..., :-1].max(-1)[0]).numpy()
        for i in range(B):
            m["level_fg"] = min(m["level_fg"], rel(conf[i], np.float32(fg_thr)))
            srt = np.sort(keys[i])[::-1]
            k = k_for_topk(nms_pre, len(srt))
            if k > 0:
                m["topk_boundary"] = min(m["topk_boundary"], (srt[k - 1] - srt[k]) / srt[k - 1])
                top = srt[:k]
                m["topk_adjacent"] = min(m["topk_adjacent"], float(np.min((top[:-1] - top[1:]) / top[:-1])))
    for i in range(B):
        sc = out["scores"][i].numpy()
        fgcols = sc if head == HEAD_RETINA else sc[:, :-1]
        m["score_thr"] = min(m["score_thr"], rel(fgcols, np.float32(score_thr)))
        m["fg_thr"] = min(m["fg_thr"], rel(sc.max(axis=1), np.float32(fg_thr)))
        top2 = np.sort(sc, axis=1)[:, -2:]
        m["argmax"] = min(m["argmax"], float(np.min((top2[:, 1] - top2[:, 0]) / top2[:, 1])))
        d = out["dets"][i].numpy()
        if len(d):
            m["obj_thr"] = min(m["obj_thr"], rel(d[:, 4], np.float32(obj_thr)))
            if len(d) > 1:
                m["nms_adjacent"] = min(m["nms_adjacent"], float(np.min((d[:-1, 4] - d[1:, 4]) / d[:-1, 4])))
            objs = out["dets"][i][out["dets"][i][:, 4] > obj_thr][:, :4]
            iou = bbox_overlaps(out["boxes"][i], objs).numpy()
            m["cluster_iou"] = min(m["cluster_iou"], rel(iou, np.float32(cluster_iou)))
            # NMS decisions among candidates at least as confident as the weakest kept det
            ncls = fgcols.shape[1]
            flat = fgcols.reshape(-1)
            cand = np.nonzero(flat >= d[-1, 4])[0]
            if len(cand) > 1:
                rows, labs = cand // ncls, cand % ncls
                bx = out["boxes"][i][torch.from_numpy(rows)]
                off = torch.from_numpy(labs).to(bx) * (bx.max() + 1)
                sh = bx + off[:, None]
                io = bbox_overlaps(sh, sh).numpy()
                io = io[np.triu_indices(len(cand), 1)]
                m["nms_iou"] = min(m["nms_iou"], rel(io[io > 0], np.float32(nms_iou)))
    return {k: float(v) for k, v in m.items()}


def compute_mi(members, n_cls: int = 20):
    """Restatement of ComputeMI (mmdet/apis/CalEnsembleUnc.py:166-181) / ComputeMCDropoutMI
    (mmdet/apis/CalMCDropoutUnc.py:185-201) for any number of members.  members[m][s]: float32
    [B, A*n_cls, H, W].  Returns (image scores float32[B], level values float32[B, S])."""
    M, S, B = len(members), len(members[0]), members[0][0].shape[0]
    buffer = torch.zeros(B, S)
    for s in range(S):
        for b in range(B):
            preds = torch.stack([torch.sigmoid(members[m][s][b]).permute(1, 2, 0).reshape(-1, n_cls) for m in range(M)])
            avg = preds.mean(dim=0)                                   # :174 / :195
            total = (-avg * avg.log()).sum(dim=1)                     # :175
            ent = (-preds * preds.log()).sum(dim=-1)                  # :176
            buffer[b, s] = (total - ent.mean(dim=0)).mean()           # :177-179
    return buffer.mean(dim=-1), buffer                                # :180


MI_SHAPES = [(9, 11), (5, 6), (3, 3), (2, 2), (1, 1)]      # level feature maps of the mutual-information goldens


def mi_inputs(seed: int, members: int, n_cls: int = 20, anchors: int = 3, batch: int = 2):
    """Seeded member logits of the MI goldens (numpy RandomState: the same bits on every platform):
    members x levels of float32 [batch, anchors*n_cls, H, W], member m = a shared map + its own perturbation."""
    rs = np.random.RandomState(seed)
    out = [[None] * len(MI_SHAPES) for _ in range(members)]
    for s, (h, w) in enumerate(MI_SHAPES):
        shared = rs.standard_normal((batch, anchors * n_cls, h, w)) * 2.0 - 2.5
        for m in range(members):
            out[m][s] = torch.from_numpy((shared + 0.7 * rs.standard_normal(shared.shape)).astype(np.float32))
    return out
