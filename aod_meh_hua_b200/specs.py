"""Detector shapes and scoring constants for the MEH -> uncertainty -> HUA pool-scoring path.

Everything here is *parameters of the path*, not model code: the level geometry that the
RetinaNet / SSD dense heads emit and the constants the reference hard-codes or reads from
its two active-learning configs.

Reference (paths relative to /root/reference):
  configs/_base_/Config_RetinaNet.py:41-62,79-85   (Lambda_L2Net head, test_cfg)
  configs/_base_/Config_SSD.py:40-54,68-74         (MyLSSDHead head, test_cfg)
  configs/ssd/ssd512_coco.py:1-17                  (SSD512 geometry)
  mmdet/models/dense_heads/Lambda_L2.py:349,500,508,514-515,520   (0.3 / 0.5 / 25 / 1e-7 / T=500)
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Sequence, Tuple

HEAD_RETINA = 0  # Lambda_L2Net: C_out = num_classes, score = p / (sum(p)+1e-20+1e-9)
HEAD_SSD = 1     # MyLSSDHead:  C_out = num_classes + 1 (background last), score = p

AGG_SUM, AGG_AVG, AGG_MAX = 0, 1, 2
AGG_POOL = 3      # class axis of the Entropy_ALL family only: no class split (Entropy_Avg)
ACT_SOFTMAX, ACT_RELU_PLUS_ONE, ACT_RELU = 0, 1, 2
_ACTIVATIONS = {"softmax": ACT_SOFTMAX, "relu_plus_one": ACT_RELU_PLUS_ONE, "relu": ACT_RELU}
_AGG_TOKENS = {"Sum": AGG_SUM, "Avg": AGG_AVG, "Max": AGG_MAX}


def parse_agg_spec(spec: str) -> Tuple[int, int, int]:
    """'objectSum_scaleMax_classSum' -> (object_op, scale_op, class_op).

    Mirrors ExtractAggFunc (mmdet/utils/functions.py:425-436): the string is split on '_';
    a token that *contains* 'object' / 'scale' / 'class' names that level's reducer, the
    remainder of the token must be one of Sum / Avg / Max (KeyError otherwise, as in the
    reference); a later token overrides an earlier one.  AggregateObjScaleUnc needs all
    three levels (Lambda_L2.py:610-614) and raises KeyError when one is missing.
    """
    found = {}
    for name in ("object", "scale", "class"):
        for tok in spec.split("_"):
            if name in tok:
                found[name] = _AGG_TOKENS[tok.replace(name, "")]
    return found["object"], found["scale"], found["class"]


_SCALE_TYPES = {"scaleAvg_classAvg": (AGG_AVG, AGG_AVG), "scaleSum_classSum": (AGG_SUM, AGG_SUM),
                "scaleSum_classAvg": (AGG_SUM, AGG_AVG), "scaleAvg_classSum": (AGG_AVG, AGG_SUM)}


def parse_scale_agg(kind: str) -> Tuple[int, int, int]:
    """Entropy_ALL aggregation type -> (object_op, scale_op, class_op).  AggregateScaleUnc
    (Lambda_L2.py:636-691) hard-codes exactly four types; there is a single pseudo-object, so the
    object reducer is Sum.  (The reference silently returns an empty list for any other string,
    which crashes its caller later; here it is a ValueError.)"""
    if kind == "Entropy_Avg":        # ComputeAvgUnc / AggregateAvgUnc (Lambda_L2_ReLU.py:446-474, 532-541): pooled level means, mean over levels
        return AGG_SUM, AGG_AVG, AGG_POOL
    if kind not in _SCALE_TYPES:
        raise ValueError(f"unknown Entropy_ALL aggregation type {kind!r}; one of {sorted(_SCALE_TYPES)}")
    sc, cl = _SCALE_TYPES[kind]
    return AGG_SUM, sc, cl


@dataclasses.dataclass(frozen=True)
class DetectorSpec:
    """Geometry + constants of one detector configuration on the scoring path."""

    name: str
    head: int                       # HEAD_RETINA / HEAD_SSD
    num_classes: int                # foreground classes C
    img_hw: Tuple[int, int]         # padded network input (H, W) = img_shape used for clipping
    strides: Tuple[int, ...]
    featmaps: Tuple[Tuple[int, int], ...]   # (H_s, W_s) per level
    num_anchors: Tuple[int, ...]    # A per level
    target_stds: Tuple[float, float, float, float]
    score_thr: float                # test_cfg.score_thr
    max_per_img: int                # test_cfg.max_per_img
    nms_pre: int = 1000
    nms_iou: float = 0.5
    # anchor generator parameters
    retina_octave_base_scale: int = 4
    retina_scales_per_octave: int = 3
    retina_ratios: Tuple[float, ...] = (0.5, 1.0, 2.0)
    ssd_input_size: int = 300
    ssd_ratio_range: Tuple[float, float] = (0.15, 0.9)
    ssd_ratios: Tuple[Tuple[int, ...], ...] = ()
    # synthetic-pool parameters
    gt_range: Tuple[int, int] = (1, 8)

    @property
    def c_out(self) -> int:
        return self.num_classes + (1 if self.head == HEAD_SSD else 0)

    @property
    def num_levels(self) -> int:
        return len(self.strides)

    @property
    def level_sizes(self) -> List[int]:
        return [h * w * a for (h, w), a in zip(self.featmaps, self.num_anchors)]

    @property
    def num_priors(self) -> int:
        return sum(self.level_sizes)

    @property
    def level_k(self) -> List[int]:
        """Rows each level contributes after the per-level top-k (get_k_for_topk:
        mmdet/core/export/onnx_helper.py:61-78 -> k only when k < size)."""
        return [min(n, self.nms_pre) if self.nms_pre > 0 else n for n in self.level_sizes]

    @property
    def k_tot(self) -> int:
        return sum(self.level_k)

    def k1_bytes_per_image(self) -> int:
        """Algorithmic HBM bytes of the alpha/top-k stage per image:
        logits read once + lambda + gathered deltas + written score rows / lambda / boxes."""
        n, c, k = self.num_priors, self.c_out, self.k_tot
        return 4 * n * c + 4 * n + 16 * k + k * (4 * c + 4 + 16)


def _retina_featmaps(h: int, w: int, n: int = 5, first_stride: int = 8):
    fh, fw = math.ceil(h / first_stride), math.ceil(w / first_stride)
    out = []
    for _ in range(n):
        out.append((fh, fw))
        fh, fw = math.ceil(fh / 2), math.ceil(fw / 2)
    return tuple(out)


def retina_spec(name: str, h: int, w: int, num_classes: int, **kw) -> DetectorSpec:
    strides = (8, 16, 32, 64, 128)
    return DetectorSpec(
        name=name, head=HEAD_RETINA, num_classes=num_classes, img_hw=(h, w), strides=strides,
        featmaps=_retina_featmaps(h, w), num_anchors=(9,) * 5,
        target_stds=(1.0, 1.0, 1.0, 1.0), score_thr=0.05, max_per_img=100, **kw)


def ssd_spec(name: str, size: int, num_classes: int, **kw) -> DetectorSpec:
    if size == 300:
        strides = (8, 16, 32, 64, 100, 300)
        fm = (38, 19, 10, 5, 3, 1)
        ratios = ((2,), (2, 3), (2, 3), (2, 3), (2,), (2,))
        # Config_SSD.py:49 uses (0.15, 0.9) with VOC, which selects the "SSD300 COCO" size table
        rng = kw.pop("ssd_ratio_range", (0.15, 0.9))
    elif size == 512:
        strides = (8, 16, 32, 64, 128, 256, 512)
        fm = (64, 32, 16, 8, 4, 2, 1)
        ratios = ((2,), (2, 3), (2, 3), (2, 3), (2, 3), (2,), (2,))
        rng = kw.pop("ssd_ratio_range", (0.1, 0.9))
    else:
        raise ValueError("SSD input size must be 300 or 512")
    na = tuple(2 + 2 * len(r) for r in ratios)
    return DetectorSpec(
        name=name, head=HEAD_SSD, num_classes=num_classes, img_hw=(size, size), strides=strides,
        featmaps=tuple((f, f) for f in fm), num_anchors=na,
        target_stds=(0.1, 0.1, 0.2, 0.2), score_thr=0.02, max_per_img=200,
        ssd_input_size=size, ssd_ratio_range=rng, ssd_ratios=ratios, **kw)


# BASELINE.json configs (SURVEY section 8 shape table)
SPECS = {
    "cfg1_retina_r50_512_voc": retina_spec("cfg1_retina_r50_512_voc", 512, 512, 20),
    "cfg2_ssd300_voc": ssd_spec("cfg2_ssd300_voc", 300, 20),
    "cfg3_retina_r50_800x1344_coco": retina_spec("cfg3_retina_r50_800x1344_coco", 800, 1344, 80),
    "cfg3p_retina_r50_800x800_coco": retina_spec("cfg3p_retina_r50_800x800_coco", 800, 800, 80),
    "cfg4_ssd512_coco": ssd_spec("cfg4_ssd512_coco", 512, 80),
    "cfg5_retina_r101_1344_coco": retina_spec("cfg5_retina_r101_1344_coco", 1344, 1344, 80,
                                              gt_range=(40, 100)),
    # small shapes for parity tests (oracle finishes in seconds)
    "tiny_retina_voc": retina_spec("tiny_retina_voc", 128, 160, 20, gt_range=(1, 4)),
    "tiny_retina_coco": retina_spec("tiny_retina_coco", 96, 128, 80, gt_range=(1, 3)),
    "tiny_ssd_voc": ssd_spec("tiny_ssd_voc", 300, 20, gt_range=(1, 4)),
    # class counts without a template instantiation (generic kernels)
    "tiny_retina_c12": retina_spec("tiny_retina_c12", 96, 128, 12, gt_range=(1, 3)),
    "tiny_ssd_c7": ssd_spec("tiny_ssd_c7", 300, 7, gt_range=(1, 4)),
}


def get_spec(name: str) -> DetectorSpec:
    return SPECS[name]


@dataclasses.dataclass(frozen=True)
class ScoringParams:
    """Constants of ComputeObjUnc / AggregateObjScaleUnc.  Defaults = reference;
    the ablation heads (Lambda_L2_ReLU.py:150-154,395-425) vary obj_thr / cluster_iou /
    use_lambda / n_samples."""

    n_samples: int = 500            # Lambda_L2.py:520
    fg_thr: float = 0.3             # Lambda_L2.py:500,508 (level-FG and per-box FG), strict >
    obj_thr: float = 0.3            # Lambda_L2.py:349 score_thr of GetObjectIdx, strict >
    cluster_iou: float = 0.5        # Lambda_L2.py:349 iou_thr of GetObjectIdx, strict >
    lambda_scale: float = 25.0      # Lambda_L2.py:515
    lambda_eps: float = 1e-7        # Lambda_L2.py:514
    use_lambda: bool = True         # False = Lambda_L2_noL.py:531
    agg: str = "objectSum_scaleMax_classSum"   # Config_RetinaNet.py:18
    cls_w: bool = False             # tools/train_RetinaNet.py:30 clsW
    seed: int = 20                  # tools/train_RetinaNet.py:80-87
    # how a class row is formed from the logits: "softmax" (the scoring heads), "relu_plus_one" (the base head's
    # evidential form, L_anchor_head.py:401-406; detection route), "relu" (Entropy_Avg, Lambda_L2_ReLU.py:453-455)
    activation: str = "softmax"

    @property
    def activation_code(self) -> int:
        return _ACTIVATIONS[self.activation]
