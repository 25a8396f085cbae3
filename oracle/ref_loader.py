"""Load the reference's OWN scoring functions, unmodified, by AST extraction (build container only).

TEST INFRASTRUCTURE ONLY - used by oracle/make_golden.py to mint golden vectors and by the
optional `-m "not gpu"` test that re-validates the restatement when /root/reference is mounted.
`/root/reference` does not exist on the GPU box; nothing imported at GPU-test / bench time may
call into this module.

`import mmdet` is impossible here (mmcv is not installed), so the needed functions and
methods are parsed out of the reference sources with `ast`, their decorators dropped
(`@force_fp32`, `@mmcv.jit`, registry decorators) and executed in a namespace that holds only
torch / numpy plus the shims listed below.  No reference source is copied into this repository:
the code objects are built in memory from the files where they lie.

Shims (the only code on the path that is not the reference's):
  * mmcv.ops.nms.batched_nms  -> class-offset + torchvision.ops.nms   (mmcv 1.3.8, un-vendored)
  * self.bbox_coder.decode    -> the reference's delta2bbox with the config's means / stds
  * self.assigner.iou_calculator -> the reference's bbox_overlaps after dropping a 5th column
    (BboxOverlaps2D.__call__, core/bbox/iou_calculators/iou2d_calculator.py:48-53)
"""
from __future__ import annotations

import ast
import os
import types
from typing import Dict, Iterable

import numpy as np
import torch
import torch.nn.functional as F
from torch.distributions import Dirichlet

REF_ROOT = os.environ.get("MEHHUA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "mmdet/models/dense_heads/Lambda_L2.py"))


def _parse(rel: str) -> ast.Module:
    with open(os.path.join(REF_ROOT, rel), "r") as f:
        return ast.parse(f.read(), filename=rel)


def _compile_defs(nodes: Iterable[ast.AST], ns: Dict[str, object], rel: str) -> None:
    for node in nodes:
        node.decorator_list = []
        mod = ast.Module(body=[node], type_ignores=[])
        ast.fix_missing_locations(mod)
        exec(compile(mod, filename=f"<reference:{rel}>", mode="exec"), ns)


def load_functions(rel: str, names: Iterable[str], ns: Dict[str, object]) -> None:
    """Module-level functions `names` of reference file `rel` -> ns."""
    want = set(names)
    tree = _parse(rel)
    nodes = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    missing = want - {n.name for n in nodes}
    if missing:
        raise KeyError(f"{rel}: functions not found: {sorted(missing)}")
    _compile_defs(nodes, ns, rel)


def load_methods(rel: str, cls_name: str, names: Iterable[str], ns: Dict[str, object]) -> Dict[str, object]:
    """Methods `names` of class `cls_name` -> dict of plain functions (first arg = self)."""
    want = set(names)
    tree = _parse(rel)
    for n in tree.body:
        if isinstance(n, ast.ClassDef) and n.name == cls_name:
            nodes = [m for m in n.body if isinstance(m, ast.FunctionDef) and m.name in want]
            missing = want - {m.name for m in nodes}
            if missing:
                raise KeyError(f"{rel}:{cls_name}: methods not found: {sorted(missing)}")
            local = dict(ns)
            _compile_defs(nodes, local, rel)
            return {m.name: local[m.name] for m in nodes}
    raise KeyError(f"{rel}: class {cls_name} not found")


def _batched_nms_shim(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    from torchvision.ops import nms as tv_nms
    cfg = dict(nms_cfg)
    assert cfg.pop("type", "nms") == "nms"
    thr = cfg.pop("iou_threshold")
    max_coordinate = boxes.max()
    offsets = idxs.to(boxes) * (max_coordinate + torch.tensor(1).to(boxes))
    keep = tv_nms(boxes + offsets[:, None], scores, thr)
    return torch.cat([boxes[keep], scores[keep][:, None]], -1), keep


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def base_namespace() -> Dict[str, object]:
    ns: Dict[str, object] = dict(torch=torch, np=np, F=F, Dirichlet=Dirichlet, os=os)
    ns["batched_nms"] = _batched_nms_shim
    load_functions("mmdet/core/bbox/iou_calculators/iou2d_calculator.py",
                   ["fp16_clamp", "bbox_overlaps"], ns)
    load_functions("mmdet/core/post_processing/bbox_nms.py", ["multiclass_nms"], ns)
    load_functions("mmdet/core/export/onnx_helper.py", ["get_k_for_topk"], ns)
    load_functions("mmdet/core/bbox/coder/delta_xywh_bbox_coder.py", ["delta2bbox"], ns)
    load_functions("mmdet/utils/functions.py", ["ExtractAggFunc", "StartEnd", "getMaxConf"], ns)
    load_functions("mmdet/utils/active_datasets.py", ["update_X_L"], ns)
    return ns


def make_head(kind: str, c_out: int, stds, score_thr: float, max_per_img: int, nms_pre: int = 1000,
              nms_iou: float = 0.5, ns: Dict[str, object] = None):
    """A stub `self` carrying the reference's real _get_bboxes / ComputeObjUnc /
    AggregateObjScaleUnc bound as methods.  kind: 'retina' (Lambda_L2Net) | 'ssd' (MyLSSDHead) |
    'retina_relu' (Lambda_L2Net_ReLU, Lambda_L2_ReLU.py:146-276, 395-444)."""
    ns = dict(base_namespace() if ns is None else ns)
    if kind == "retina":
        rel, cls, act = "mmdet/models/dense_heads/Lambda_L2.py", "Lambda_L2Net", "relu"
    elif kind == "ssd":
        rel, cls, act = "mmdet/models/dense_heads/My_L_ssd_head.py", "MyLSSDHead", "softmax"
        ns["ignoreBG"] = False          # My_L_ssd_head.py:19
    elif kind == "retina_relu":         # ablation head: thresholds from kwargs, alpha = scores (no lambda')
        rel, cls, act = "mmdet/models/dense_heads/Lambda_L2_ReLU.py", "Lambda_L2Net_ReLU", "relu"
    elif kind == "retina_nol":          # Lambda_L2Net_NoL: thresholds from kwargs, alpha = scores (no lambda')
        rel, cls, act = "mmdet/models/dense_heads/Lambda_L2_noL.py", "Lambda_L2Net_NoL", "relu"
    elif kind == "retina_ablation":     # ablation head: thresholds from kwargs, lambda' kept
        rel, cls, act = "mmdet/models/dense_heads/Lambda_L2_ablation.py", "Lambda_L2Net_ablation", "relu"
    elif kind == "base_relu":           # the base head: detection route with alpha = relu + 1 (L_anchor_head.py:358-464)
        rel, cls, act = "mmdet/models/dense_heads/L_anchor_head.py", "L_AnchorHead", "relu"
    else:
        raise ValueError(kind)
    names = ["_get_bboxes", "ComputeObjUnc", "AggregateObjScaleUnc", "ComputeScaleUnc",
             "AggregateScaleUnc"]
    if kind in ("retina_relu", "retina_nol"):
        names += ["ComputeAvgUnc", "AggregateAvgUnc"]
    if kind == "base_relu":
        names = ["_get_bboxes"]
        # the base method imports `get_k_for_topk` from mmdet.core.export inside its body (L_anchor_head.py:413);
        # mmdet is not importable here, so a stub package that holds the reference's own function stands in
        import sys
        if "mmdet" not in sys.modules:
            for mod in ("mmdet", "mmdet.core", "mmdet.core.export"):
                sys.modules[mod] = types.ModuleType(mod)
            sys.modules["mmdet.core.export"].get_k_for_topk = ns["get_k_for_topk"]
            sys.modules["mmdet"].core = sys.modules["mmdet.core"]
            sys.modules["mmdet.core"].export = sys.modules["mmdet.core.export"]
    fns = load_methods(rel, cls, names, ns)
    head = types.SimpleNamespace()
    head.cls_out_channels = c_out
    head.last_activation = act
    head.test_cfg = _Cfg(nms_pre=nms_pre, min_bbox_size=0, score_thr=score_thr,
                         nms=dict(type="nms", iou_threshold=nms_iou), max_per_img=max_per_img)
    d2b = ns["delta2bbox"]
    head.bbox_coder = types.SimpleNamespace(
        decode=lambda rois, deltas, max_shape=None: d2b(rois, deltas, (0., 0., 0., 0.), tuple(stds), max_shape))
    ovl = ns["bbox_overlaps"]

    def iou_calc(b1, b2, mode="iou", is_aligned=False):
        if b2.size(-1) == 5:
            b2 = b2[..., :4]
        if b1.size(-1) == 5:
            b1 = b1[..., :4]
        return ovl(b1, b2, mode, is_aligned)

    head.assigner = types.SimpleNamespace(iou_calculator=iou_calc)
    for k, f in fns.items():
        setattr(head, k, types.MethodType(f, head))
    head._ns = ns
    head._fn_globals = fns["ComputeObjUnc" if "ComputeObjUnc" in fns else "_get_bboxes"].__globals__
    return head


def load_anchor_generators():
    """The reference's AnchorGenerator / SSDAnchorGenerator classes with mmcv stubs."""
    ns: Dict[str, object] = dict(torch=torch, np=np)
    import warnings
    from torch.nn.modules.utils import _pair
    ns["warnings"] = warnings
    ns["_pair"] = _pair
    ns["mmcv"] = types.SimpleNamespace(
        is_tuple_of=lambda seq, t: isinstance(seq, tuple) and all(isinstance(v, t) for v in seq))
    rel = "mmdet/core/anchor/anchor_generator.py"
    tree = _parse(rel)
    for n in tree.body:
        if isinstance(n, ast.ClassDef) and n.name in ("AnchorGenerator", "SSDAnchorGenerator"):
            n.decorator_list = []
            mod = ast.Module(body=[n], type_ignores=[])
            ast.fix_missing_locations(mod)
            exec(compile(mod, filename=f"<reference:{rel}>", mode="exec"), ns)
    return ns["AnchorGenerator"], ns["SSDAnchorGenerator"]
