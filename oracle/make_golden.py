"""Mint golden vectors by running the REFERENCE'S OWN functions (build container only).

    python -m oracle.make_golden            # writes tests/golden/*.npz

The reference has no fixtures for this path, so the goldens are outputs of its own
`_get_bboxes` -> `ComputeObjUnc` -> `AggregateObjScaleUnc`, `delta2bbox`, anchor generators and
`update_X_L`, AST-loaded unmodified from /root/reference (oracle/ref_loader.py) and executed on
seeded synthetic head outputs (aod_meh_hua_b200/synth.py).  Inputs are NOT stored: tests regenerate
them from the seed and verify a checksum.  Dirichlet draws come from torch's CPU generator seeded
with `sample_seed`; alpha per (image, level) and every derived value are stored.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F
from torch.distributions import Dirichlet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from aod_meh_hua_b200.specs import HEAD_RETINA, get_spec  # noqa: E402
from aod_meh_hua_b200.synth import SyntheticPool  # noqa: E402
from aod_meh_hua_b200 import anchors as my_anchors  # noqa: E402
from oracle import ref_loader as RL  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, spec, image ids, pool seed, sample seed, scale factor, uPool2, clsW
    ("retina_voc", "tiny_retina_voc", [0, 1], 20, 1234, (1.0, 1.0, 1.0, 1.0), "objectSum_scaleMax_classSum", False),
    ("retina_coco", "tiny_retina_coco", [0, 1, 2], 20, 1234, (1.0, 1.0, 1.0, 1.0), "objectSum_scaleMax_classSum", False),
    ("retina_coco_aniso", "tiny_retina_coco", [3, 4], 20, 99, (1.07, 0.94, 1.07, 0.94), "objectAvg_scaleSum_classMax", True),
    ("ssd_voc", "tiny_ssd_voc", [0, 1], 20, 1234, (1.0, 1.0, 1.0, 1.0), "objectSum_scaleMax_classSum", False),
    # one BASELINE.json configuration at full size (config 1: RetinaNet R50-FPN 512x512, 20 VOC classes, 49 104 priors)
    ("full_cfg1_retina_voc", "cfg1_retina_r50_512_voc", [0, 1], 20, 1234, (1.0, 1.0, 1.0, 1.0), "objectSum_scaleMax_classSum", False),
    # ... and the SSD head at full size (config 2: SSD300 VGG16, 21 outputs, 8 732 priors)
    ("full_cfg2_ssd300_voc", "cfg2_ssd300_voc", [0, 1], 20, 4321, (1.0, 1.0, 1.0, 1.0), "objectSum_scaleMax_classSum", False),
    # ... and the headline shape (config 3: RetinaNet R50-FPN 800x1344 COCO, 201 600 priors x 80 classes), one image
    ("full_cfg3_retina_coco", "cfg3_retina_r50_800x1344_coco", [0], 20, 777, (1.0, 1.0, 1.0, 1.0), "objectSum_scaleMax_classSum", False),
]


VARIANT_CASES = [
    # the ablation head Lambda_L2Net_ReLU: thresholds from kwargs (Lambda_L2_ReLU.py:150-154, 397-400),
    # alpha = score row without lambda' (:425-427); Lambda_L2Net_ablation: the same thresholds with lambda' kept.
    # name, spec, ids, pool seed, sample seed, score_thr, iou_thr, head kind, use_lambda
    ("relu_retina_coco", "tiny_retina_coco", [0, 1, 2], 20, 77, 0.4, 0.6, "retina_relu", False),
    ("ablation_retina_voc", "tiny_retina_voc", [0, 1], 20, 78, 0.35, 0.9, "retina_ablation", True),
]


ALL_CASES = [
    # name, spec, image ids, pool seed, sample seed, aggregation type
    ("all_retina_coco", "tiny_retina_coco", [0, 1, 2], 20, 5, "scaleAvg_classAvg"),
    ("all_ssd_voc", "tiny_ssd_voc", [0, 1, 2], 20, 5, "scaleSum_classAvg"),
    # the same route at full size (config 1: 49 104 priors, hundreds of foreground priors per image)
    ("all_full_cfg1_retina_voc", "cfg1_retina_r50_512_voc", [0, 1], 20, 8, "scaleSum_classSum"),
]

ALL_VARIANT_CASES = [
    # Entropy_ALL route of Lambda_L2Net_NoL (alpha = softmax, no lambda'): name, spec, ids, seeds, type, head kind
    ("all_nol_retina_coco", "tiny_retina_coco", [0, 1, 2], 20, 6, "scaleSum_classSum", "retina_nol"),
]


RPO_CASES = [
    # detection route of the base head L_AnchorHead with last_activation='relu': alpha = relu(logits) + 1
    # (L_anchor_head.py:358-464): name, spec, image ids, pool seed
    ("rpo_retina_voc", "tiny_retina_voc", [0, 1, 2], 20),
]

AVG_CASES = [
    # Entropy_Avg route of Lambda_L2Net_ReLU (ComputeAvgUnc + AggregateAvgUnc): name, spec, image ids, pool seed, sample seed
    ("avg_retina_voc", "tiny_retina_voc", [0, 1, 2], 20, 11),
]


def batch_checksum(batch) -> str:
    h = hashlib.sha256()
    for key in ("cls_scores", "bbox_preds", "L_scores", "anchors"):
        for t in batch[key]:
            h.update(t.contiguous().numpy().tobytes())
    return h.hexdigest()


def run_reference_case(spec_name, gids, pool_seed, sample_seed, sf, upool2, clsw, kind=None, score_thr=0.3, iou_thr=0.9):
    spec = get_spec(spec_name)
    batch = SyntheticPool(spec, seed0=pool_seed, scale_factor=sf).batch(gids)
    kind = kind or ("retina" if spec.head == HEAD_RETINA else "ssd")
    head = RL.make_head(kind, spec.c_out, spec.target_stds, spec.score_thr, spec.max_per_img, spec.nms_pre,
                        spec.nms_iou)
    rec = []

    class RecordingDirichlet:
        def __init__(self, alpha):
            self.alpha = alpha
            self.dist = Dirichlet(alpha)

        def sample(self, n):
            s = self.dist.sample(n)
            rec.append((self.alpha.clone(), s))
            return s

    head._fn_globals["Dirichlet"] = RecordingDirichlet
    captured = {}
    real_compute = head.ComputeObjUnc

    def compute_spy(*a, **k):
        captured["pos_bboxes"] = [p.clone() for p in a[1]]
        captured["nested"] = real_compute(*a, **k)
        return captured["nested"]

    head.ComputeObjUnc = compute_spy
    kw = dict(isUnc="Epistemic", uPool="Entropy_NMS", uPool2=upool2, L_scores=batch["L_scores"], isEval=False,
              showNMS=False, saveUnc=False, saveMaxConf=False, clsW=clsw, scaleUnc=False, score_thr=score_thr,
              iou_thr=iou_thr, batchIdx=0, return_box=False)
    torch.manual_seed(sample_seed)
    dets, unc = head._get_bboxes(batch["cls_scores"], batch["bbox_preds"], batch["anchors"], batch["img_shapes"],
                                 [np.asarray(s, dtype=np.float32) for s in batch["scale_factors"]], None, True, True,
                                 **kw)
    g = dict(checksum=np.frombuffer(bytes.fromhex(batch_checksum(batch)), dtype=np.uint8),
             image_scores=np.asarray(unc, dtype=np.float64), n_images=np.int64(len(gids)))
    for b, (d, l) in enumerate(dets):
        g[f"dets_{b}"] = d.numpy()
        g[f"labels_{b}"] = l.numpy()
        g[f"pos_nz_{b}"] = captured["pos_bboxes"][b].nonzero().numpy().astype(np.int32)
        g[f"pos_shape_{b}"] = np.asarray(captured["pos_bboxes"][b].shape, dtype=np.int64)
    groups = []
    for b, img in enumerate(captured["nested"]):
        for o, obj in enumerate(img):
            for s, lvl in enumerate(obj):
                for c, (ale, epi) in lvl.items():
                    groups.append((b, o, s, int(c), float(ale), float(epi)))
    g["groups"] = np.asarray(groups, dtype=np.float64).reshape(-1, 6)
    h = hashlib.sha256()
    for k, (alpha, smp) in enumerate(rec):
        g[f"alpha_{k}"] = alpha.numpy()
        h.update(smp.numpy().tobytes())
    g["n_blocks"] = np.int64(len(rec))
    g["samples_sha256"] = np.frombuffer(h.digest(), dtype=np.uint8)
    if rec:  # one small block of raw samples so injection can be tested without torch's RNG
        k = int(np.argmin([a.shape[0] for a, _ in rec]))
        g["sample_block_index"] = np.int64(k)
        g["sample_block"] = rec[k][1].numpy()
    return g


def run_reference_all_case(spec_name, gids, pool_seed, sample_seed, kind, head_kind=None):
    """Entropy_ALL route of the reference's _get_bboxes (with_nms=False) -> ComputeScaleUnc ->
    AggregateScaleUnc, all four aggregation types on the same draws."""
    spec = get_spec(spec_name)
    batch = SyntheticPool(spec, seed0=pool_seed).batch(gids)
    head = RL.make_head(head_kind or ("retina" if spec.head == HEAD_RETINA else "ssd"), spec.c_out, spec.target_stds,
                        spec.score_thr, spec.max_per_img, spec.nms_pre, spec.nms_iou)
    g = dict(checksum=np.frombuffer(bytes.fromhex(batch_checksum(batch)), dtype=np.uint8))
    captured = {}
    real = head.ComputeScaleUnc

    def spy(*a, **k):
        captured["nested"] = real(*a, **k)
        return captured["nested"]

    head.ComputeScaleUnc = spy
    for typ in ("scaleAvg_classAvg", "scaleSum_classSum", "scaleSum_classAvg", "scaleAvg_classSum"):
        kw = dict(isUnc="Epistemic", uPool="Entropy_ALL", uPool2=typ, L_scores=batch["L_scores"], isEval=False,
                  showNMS=False, saveUnc=False, saveMaxConf=False, clsW=False, scaleUnc=False, batchIdx=0)
        torch.manual_seed(sample_seed)
        dets, unc = head._get_bboxes(batch["cls_scores"], batch["bbox_preds"], batch["anchors"], batch["img_shapes"],
                                     [np.asarray(s, dtype=np.float32) for s in batch["scale_factors"]], None, True,
                                     False, **kw)
        g[f"scores_{typ}"] = np.asarray(unc, dtype=np.float64)
    groups = []
    for b, img in enumerate(captured["nested"]):
        for s, lvl in enumerate(img):
            for c, (ale, epi) in lvl.items():
                groups.append((b, s, int(c), float(ale), float(epi)))
    g["groups"] = np.asarray(groups, dtype=np.float64).reshape(-1, 5)
    g["det_shapes"] = np.asarray([list(d[0].shape) + list(d[1].shape) for d in dets], dtype=np.int64)
    return g


def kat_goldens():
    ns = RL.base_namespace()
    g = {}
    # docstring known-answer of delta2bbox (delta_xywh_bbox_coder.py:190-203)
    rois = torch.Tensor([[0., 0., 1., 1.], [0., 0., 1., 1.], [0., 0., 1., 1.], [5., 5., 5., 5.]])
    deltas = torch.Tensor([[0., 0., 0., 0.], [1., 1., 1., 1.], [0., 0., 2., -1.], [0.7, -1.9, -0.5, 0.3]])
    g["d2b_rois"], g["d2b_deltas"] = rois.numpy(), deltas.numpy()
    g["d2b_out"] = ns["delta2bbox"](rois, deltas, max_shape=(32, 32, 3)).numpy()
    # seeded random decode incl. clamping and clipping, SSD stds
    gen = torch.Generator().manual_seed(7)
    r = torch.rand(64, 2, generator=gen) * 200
    rois2 = torch.cat([r, r + torch.rand(64, 2, generator=gen) * 100 + 1], dim=1)
    d2 = torch.randn(64, 4, generator=gen) * 3
    g["d2b2_rois"], g["d2b2_deltas"] = rois2.numpy(), d2.numpy()
    g["d2b2_out"] = ns["delta2bbox"](rois2, d2, (0., 0., 0., 0.), (0.1, 0.1, 0.2, 0.2), (120, 160, 3)).numpy()
    # bbox_overlaps on seeded boxes (+ empty second operand)
    b1, b2 = rois2[:40], rois2[30:]
    g["iou_b1"], g["iou_b2"] = b1.numpy(), b2.numpy()
    g["iou_out"] = ns["bbox_overlaps"](b1, b2).numpy()
    g["iou_empty_shape"] = np.asarray(ns["bbox_overlaps"](b1, b2[:0]).shape, dtype=np.int64)
    # anchors: reference generators vs this repo's, per named config (stored as checksums)
    AG, SAG = RL.load_anchor_generators()
    import warnings
    for name in ("cfg1_retina_r50_512_voc", "cfg2_ssd300_voc", "cfg3_retina_r50_800x1344_coco", "cfg4_ssd512_coco",
                 "tiny_retina_voc", "tiny_ssd_voc"):
        spec = get_spec(name)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if spec.head == HEAD_RETINA:
                gen_ = AG(strides=list(spec.strides), ratios=list(spec.retina_ratios),
                          octave_base_scale=spec.retina_octave_base_scale,
                          scales_per_octave=spec.retina_scales_per_octave)
            else:
                gen_ = SAG(strides=list(spec.strides), ratios=[list(r) for r in spec.ssd_ratios],
                           basesize_ratio_range=spec.ssd_ratio_range, input_size=spec.ssd_input_size,
                           scale_major=False)
            ref = gen_.grid_anchors(list(spec.featmaps), device="cpu")
        mine = my_anchors.grid_anchors(spec)
        h = hashlib.sha256()
        for r_, m_ in zip(ref, mine):
            assert torch.equal(r_, m_), f"{name}: anchors differ from the reference generator"
            h.update(r_.numpy().tobytes())
        g[f"anchors_sha256_{name}"] = np.frombuffer(h.digest(), dtype=np.uint8)
    # docstring known-answer of AnchorGenerator (anchor_generator.py:41-57)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g["anchor_kat"] = AG([16], [1.], [1.], [9]).grid_anchors([(2, 2)], device="cpu")[0].numpy()
    # ExtractAggFunc tokens
    import torch as _t
    names = {_t.sum: 0, _t.mean: 1, _t.max: 2}
    for spec_str in ("objectSum_scaleMax_classSum", "objectAvg_scaleSum_classMax", "objectMax_scaleAvg_classAvg"):
        f = ns["ExtractAggFunc"](spec_str)
        g[f"agg_{spec_str}"] = np.asarray([names[f["object"]], names[f["scale"]], names[f["class"]]], dtype=np.int64)
    # update_X_L on a seeded pool with exact zeros (utils/active_datasets.py:102-135)
    rs = np.random.RandomState(3)
    n = 2000
    unc = rs.rand(n).astype(np.float32)
    unc[rs.rand(n) < 0.25] = 0.0
    X_all = np.arange(n)
    X_L = np.sort(rs.choice(n, 100, replace=False))
    np.random.seed(11)
    xl, xu = ns["update_X_L"](unc.copy(), X_all, X_L.copy(), 80, zeroRate=0.15, maxconf=None, useMaxConf="False")
    g["sel_unc"], g["sel_X_L"] = unc, X_L
    g["sel_X_L_next"], g["sel_X_U_next"] = xl, xu
    np.random.seed(11)
    xl2, xu2 = ns["update_X_L"](unc.copy(), X_all, X_L.copy(), 80)
    g["sel2_X_L_next"], g["sel2_X_U_next"] = xl2, xu2
    # update_X_L with the zero-score picks taken by max confidence (useMaxConf 'min' / 'max', :113-119)
    maxconf = rs.rand(n).astype(np.float32)
    g["sel_maxconf"] = maxconf
    for mode in ("min", "max"):
        np.random.seed(11)
        xl3, xu3 = ns["update_X_L"](unc.copy(), X_all, X_L.copy(), 80, zeroRate=0.15, maxconf=maxconf.tolist(),
                                    useMaxConf=mode)
        g[f"sel_{mode}_X_L_next"], g[f"sel_{mode}_X_U_next"] = xl3, xu3
    # getMaxConf (utils/functions.py:467-476) on the tiny batches
    RL.load_functions("mmdet/utils/functions.py", ["getMaxConf"], ns)
    for spec_name in ("tiny_retina_coco", "tiny_ssd_voc"):
        spec = get_spec(spec_name)
        batch = SyntheticPool(spec, seed0=20).batch([0, 1, 2])
        per_img, per_level = ns["getMaxConf"](batch["cls_scores"], spec.c_out)
        g[f"maxconf_{spec_name}"] = np.asarray(per_img, dtype=np.float64)
        g[f"maxconf_levels_{spec_name}"] = per_level.numpy()
    return g


def run_reference_rpo_case(spec_name, gids, pool_seed):
    """det_results of L_AnchorHead._get_bboxes (with_nms=True, rescale=True) with last_activation='relu'."""
    spec = get_spec(spec_name)
    batch = SyntheticPool(spec, seed0=pool_seed).batch(gids)
    head = RL.make_head("base_relu", spec.c_out, spec.target_stds, spec.score_thr, spec.max_per_img, spec.nms_pre,
                        spec.nms_iou)
    dets = head._get_bboxes(batch["cls_scores"], batch["bbox_preds"], batch["anchors"], batch["img_shapes"],
                            [np.asarray(s, dtype=np.float32) for s in batch["scale_factors"]], None, True, True)
    g = dict(checksum=np.frombuffer(bytes.fromhex(batch_checksum(batch)), dtype=np.uint8))
    for b, (d, l) in enumerate(dets):
        g[f"dets_{b}"] = d.numpy()
        g[f"labels_{b}"] = l.numpy()
    return g


def run_reference_avg_case(spec_name, gids, pool_seed, sample_seed):
    """Entropy_Avg route of Lambda_L2Net_ReLU._get_bboxes.  The reference (pinned to torch 1.5) builds
    Dirichlet(alpha) with alpha == 0 entries (relu of a negative logit); torch >= 1.8 refuses that unless argument
    validation is off, so the class is handed to the reference code with validate_args=False - the one shim here."""
    spec = get_spec(spec_name)
    batch = SyntheticPool(spec, seed0=pool_seed).batch(gids)
    head = RL.make_head("retina_relu", spec.c_out, spec.target_stds, spec.score_thr, spec.max_per_img, spec.nms_pre,
                        spec.nms_iou)
    head._fn_globals["Dirichlet"] = lambda alpha: Dirichlet(alpha, validate_args=False)
    captured = {}
    real = head.ComputeAvgUnc

    def spy(*a, **k):
        captured["nested"] = real(*a, **k)
        return captured["nested"]

    head.ComputeAvgUnc = spy
    kw = dict(isUnc="Epistemic", uPool="Entropy_Avg", uPool2="", L_scores=batch["L_scores"], isEval=False, showNMS=False,
              saveUnc=False, saveMaxConf=False, clsW=False, scaleUnc=False, batchIdx=0, score_thr=0.3, iou_thr=0.5,
              return_box=False)
    torch.manual_seed(sample_seed)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dets, unc = head._get_bboxes(batch["cls_scores"], batch["bbox_preds"], batch["anchors"], batch["img_shapes"],
                                     [np.asarray(s, dtype=np.float32) for s in batch["scale_factors"]], None, True, True, **kw)
    levels = np.asarray([[(v if v else np.nan) for v in img] for img in captured["nested"]], dtype=np.float64)
    return dict(checksum=np.frombuffer(bytes.fromhex(batch_checksum(batch)), dtype=np.uint8),
                image_scores=np.asarray(unc, dtype=np.float64), level_means=levels)


def mi_goldens():
    """ComputeMI (apis/CalEnsembleUnc.py:166-181) and ComputeMCDropoutMI (apis/CalMCDropoutUnc.py:185-201), the
    reference's own functions, on seeded member logits."""
    ns = dict(torch=torch, F=F, np=np)
    import warnings
    RL.load_functions("mmdet/apis/CalEnsembleUnc.py", ["ComputeMI"], ns)
    from oracle.meh_hua_oracle import mi_inputs
    ens = mi_inputs(301, 3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g = dict(ensemble=np.asarray(ns["ComputeMI"](*ens, nCls=20), dtype=np.float64))
        RL.load_functions("mmdet/apis/CalMCDropoutUnc.py", ["ComputeMCDropoutMI"], ns)
        mc = mi_inputs(302, 7)
        g["mcdropout"] = np.asarray(ns["ComputeMCDropoutMI"](*mc, nCls=20), dtype=np.float64)
    g["ensemble_checksum"] = np.frombuffer(hashlib.sha256(b"".join(t.numpy().tobytes() for m in ens for t in m)).digest(), dtype=np.uint8)
    g["mcdropout_checksum"] = np.frombuffer(hashlib.sha256(b"".join(t.numpy().tobytes() for m in mc for t in m)).digest(), dtype=np.uint8)
    return g


def main():
    assert RL.available(), "reference tree not mounted"
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only_kats = "--only-kats" in sys.argv
    for name, spec_name, gids, pseed, sseed, sf, up2, clsw in ([] if only_kats else CASES):
        g = run_reference_case(spec_name, gids, pseed, sseed, sf, up2, clsw)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **g)
        print(f"{path}: scores {g['image_scores']}, {os.path.getsize(path)} bytes")
    for name, spec_name, gids, pseed, sseed, thr, iou, kind, _ in ([] if only_kats or "--skip-variants" in sys.argv else VARIANT_CASES):
        g = run_reference_case(spec_name, gids, pseed, sseed, (1.0, 1.0, 1.0, 1.0), "objectSum_scaleMax_classSum", False,
                               kind=kind, score_thr=thr, iou_thr=iou)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **g)
        print(f"{path}: scores {g['image_scores']}, {os.path.getsize(path)} bytes")
    for name, spec_name, gids, pseed, sseed, kind in ([] if only_kats else ALL_CASES):
        g = run_reference_all_case(spec_name, gids, pseed, sseed, kind)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **g)
        print(f"{path}: {g['scores_' + kind]}, {os.path.getsize(path)} bytes")
    for name, spec_name, gids, pseed, sseed, kind, head_kind in ([] if only_kats else ALL_VARIANT_CASES):
        g = run_reference_all_case(spec_name, gids, pseed, sseed, kind, head_kind)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **g)
        print(f"{path}: {g['scores_' + kind]}, {os.path.getsize(path)} bytes")
    for name, spec_name, gids, pseed in ([] if only_kats else RPO_CASES):
        g = run_reference_rpo_case(spec_name, gids, pseed)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **g)
        print(f"{path}: {[len(g[f'dets_{b}']) for b in range(len(gids))]} detections, {os.path.getsize(path)} bytes")
    for name, spec_name, gids, pseed, sseed in ([] if only_kats else AVG_CASES):
        g = run_reference_avg_case(spec_name, gids, pseed, sseed)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **g)
        print(f"{path}: {g['image_scores']}, {os.path.getsize(path)} bytes")
    if not only_kats:
        path = os.path.join(GOLDEN_DIR, "mi_baselines.npz")
        g = mi_goldens()
        np.savez_compressed(path, **g)
        print(f"{path}: ensemble {g['ensemble']}, mcdropout {g['mcdropout']}")
    path = os.path.join(GOLDEN_DIR, "kats.npz")
    np.savez_compressed(path, **kat_goldens())
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
