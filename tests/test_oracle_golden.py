"""CPU: the oracle restatement against golden vectors minted from the reference's own functions
(oracle/make_golden.py).  No GPU, no /root/reference needed at test time."""
import hashlib
import os

import numpy as np
import pytest
import torch

from aod_meh_hua_b200 import anchors as A
from aod_meh_hua_b200.specs import ScoringParams, get_spec, parse_agg_spec
from aod_meh_hua_b200.synth import SyntheticPool
from oracle import meh_hua_oracle as O
from oracle.make_golden import ALL_CASES, ALL_VARIANT_CASES, AVG_CASES, CASES, RPO_CASES, VARIANT_CASES, batch_checksum

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, f"{name}.npz"))


@pytest.mark.parametrize("case", VARIANT_CASES, ids=[c[0] for c in VARIANT_CASES])
def test_ablation_head_parameters_match_the_reference_variant(case):
    """Lambda_L2Net_ReLU (Lambda_L2_ReLU.py:146-276, 395-444) and Lambda_L2Net_ablation
    (Lambda_L2_ablation.py): thresholds from kwargs drive the object filter, the cluster IoU and both
    foreground tests; the ReLU head's alpha is the score row without lambda'.  The restatement with
    ScoringParams(fg_thr = obj_thr = score_thr, cluster_iou = iou_thr, use_lambda = ...) reproduces
    the variant heads' own outputs."""
    name, spec_name, gids, pseed, sseed, thr, iou, _, use_lambda = case
    _check_against_golden(name, spec_name, gids, pseed, sseed, (1.0, 1.0, 1.0, 1.0),
                          ScoringParams(fg_thr=thr, obj_thr=thr, cluster_iou=iou, use_lambda=use_lambda))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_restatement_matches_reference_outputs(case):
    name, spec_name, gids, pseed, sseed, sf, up2, clsw = case
    _check_against_golden(name, spec_name, gids, pseed, sseed, sf, ScoringParams(agg=up2, cls_w=clsw))


def _check_against_golden(name, spec_name, gids, pseed, sseed, sf, params):
    g = _load(name)
    spec = get_spec(spec_name)
    batch = SyntheticPool(spec, seed0=pseed, scale_factor=sf).batch(gids)
    assert bytes.fromhex(batch_checksum(batch)) == g["checksum"].tobytes(), "synthetic inputs drifted"
    blocks = []

    def sampler(alpha, T, i, s):
        smp = O.default_sampler(alpha, T, i, s)
        blocks.append((alpha.clone(), smp))
        return smp

    torch.manual_seed(sseed)
    out = O.score_batch(batch, sampler=sampler, **O.spec_kwargs(spec, params))
    # detections and cluster membership: bit-exact
    for b in range(len(gids)):
        assert np.array_equal(out["dets"][b].numpy(), g[f"dets_{b}"])
        assert np.array_equal(out["labels"][b].numpy(), g[f"labels_{b}"])
        assert tuple(out["pos_bboxes"][b].shape) == tuple(g[f"pos_shape_{b}"])
        assert np.array_equal(out["pos_bboxes"][b].nonzero().numpy().astype(np.int32), g[f"pos_nz_{b}"])
    # alpha per (image, level) block, in the reference's visiting order: bit-exact
    assert len(blocks) == int(g["n_blocks"])
    h = hashlib.sha256()
    for k, (alpha, smp) in enumerate(blocks):
        assert np.array_equal(alpha.numpy(), g[f"alpha_{k}"])
        h.update(smp.numpy().tobytes())
    same_rng = h.digest() == g["samples_sha256"].tobytes()
    # group table keys exact; values exact when torch's CPU generator reproduced the draws
    groups = []
    for b, img in enumerate(out["nested"]):
        for o, obj in enumerate(img):
            for s, lvl in enumerate(obj):
                for c, (ale, epi) in lvl.items():
                    groups.append((b, o, s, int(c), float(ale), float(epi)))
    groups = np.asarray(groups, dtype=np.float64).reshape(-1, 6)
    assert np.array_equal(groups[:, :4], g["groups"][:, :4])
    if same_rng:
        assert np.array_equal(groups[:, 4:], g["groups"][:, 4:])
        assert np.array_equal(np.asarray(out["image_scores"], dtype=np.float64), g["image_scores"])
    else:  # different torch build: the draws differ, compare statistically
        np.testing.assert_allclose(np.asarray(out["image_scores"]), g["image_scores"], rtol=0.1)


@pytest.mark.parametrize("case", [c + (None,) for c in ALL_CASES] + ALL_VARIANT_CASES, ids=lambda c: c[0])
def test_entropy_all_restatement_matches_reference_outputs(case):
    name, spec_name, gids, pseed, sseed, kind, head_kind = case
    use_lambda = head_kind != "retina_nol"
    g = _load(name)
    spec = get_spec(spec_name)
    batch = SyntheticPool(spec, seed0=pseed).batch(gids)
    assert bytes.fromhex(batch_checksum(batch)) == g["checksum"].tobytes(), "synthetic inputs drifted"
    for typ in ("scaleAvg_classAvg", "scaleSum_classSum", "scaleSum_classAvg", "scaleAvg_classSum"):
        torch.manual_seed(sseed)
        out = O.score_batch_all(batch, kind=typ, **O.spec_kwargs(spec, ScoringParams(use_lambda=use_lambda)))
        groups = np.asarray([(b, s, int(c), float(ale), float(epi)) for b, img in enumerate(out["nested"])
                             for s, lvl in enumerate(img) for c, (ale, epi) in lvl.items()], dtype=np.float64).reshape(-1, 5)
        assert np.array_equal(groups[:, :3], g["groups"][:, :3])          # (image, level, class) keys: exact
        if np.array_equal(groups[:, 3:], g["groups"][:, 3:]):             # torch's CPU generator reproduced the draws
            assert np.array_equal(np.asarray(out["image_scores"], dtype=np.float64), g[f"scores_{typ}"])
        else:
            np.testing.assert_allclose(np.asarray(out["image_scores"], dtype=np.float64), g[f"scores_{typ}"], rtol=0.1)
    assert O.aggregate_scale_unc(out["nested"], "objectSum_scaleMax_classSum") == []   # unknown type, as the reference


def test_injected_block_reproduces_reference_uncertainty():
    """uncertainty_from_samples on the stored raw sample block gives the reference's numbers."""
    g = _load("retina_coco")
    k = int(g["sample_block_index"])
    smp = torch.from_numpy(g["sample_block"])
    assert smp.shape[0] == 500 and smp.shape[1] == g[f"alpha_{k}"].shape[0]
    total, ale, epi = O.uncertainty_from_samples(smp)
    assert torch.isfinite(epi).all()
    assert float(smp.min()) >= 1.17549435e-38 and float(smp.max()) <= 0.99999994


def test_kats():
    g = _load("kats")
    out = O.delta2bbox(torch.from_numpy(g["d2b_rois"]), torch.from_numpy(g["d2b_deltas"]), (1., 1., 1., 1.),
                       max_shape=(32, 32, 3))
    assert np.array_equal(out.numpy(), g["d2b_out"])
    # the docstring's printed values (delta_xywh_bbox_coder.py:198-203)
    np.testing.assert_allclose(out.numpy(), [[0, 0, 1, 1], [0.1409, 0.1409, 2.8591, 2.8591],
                                             [0, 0.3161, 4.1945, 0.6839], [5, 5, 5, 5]], atol=5e-5)
    out2 = O.delta2bbox(torch.from_numpy(g["d2b2_rois"]), torch.from_numpy(g["d2b2_deltas"]), (0.1, 0.1, 0.2, 0.2),
                        max_shape=(120, 160, 3))
    assert np.array_equal(out2.numpy(), g["d2b2_out"])
    iou = O.bbox_overlaps(torch.from_numpy(g["iou_b1"]), torch.from_numpy(g["iou_b2"]))
    assert np.array_equal(iou.numpy(), g["iou_out"])
    assert tuple(O.bbox_overlaps(torch.from_numpy(g["iou_b1"]), torch.zeros(0, 4)).shape) == tuple(g["iou_empty_shape"])
    for name in ("cfg1_retina_r50_512_voc", "cfg2_ssd300_voc", "cfg3_retina_r50_800x1344_coco", "cfg4_ssd512_coco",
                 "tiny_retina_voc", "tiny_ssd_voc"):
        h = hashlib.sha256()
        for a in A.grid_anchors(get_spec(name)):
            h.update(a.numpy().tobytes())
        assert h.digest() == g[f"anchors_sha256_{name}"].tobytes(), name
    # AnchorGenerator docstring KAT (anchor_generator.py:41-46)
    want = np.array([[-4.5, -4.5, 4.5, 4.5], [11.5, -4.5, 20.5, 4.5], [-4.5, 11.5, 4.5, 20.5], [11.5, 11.5, 20.5, 20.5]],
                    dtype=np.float32)
    assert np.array_equal(g["anchor_kat"], want)
    for spec_str in ("objectSum_scaleMax_classSum", "objectAvg_scaleSum_classMax", "objectMax_scaleAvg_classAvg"):
        assert list(parse_agg_spec(spec_str)) == list(g[f"agg_{spec_str}"])
    with pytest.raises(KeyError):
        parse_agg_spec("scaleAvg_classAvg")      # no object token -> KeyError, as in the reference


def test_update_X_L_restatement():
    g = _load("kats")
    unc, X_L = g["sel_unc"], g["sel_X_L"]
    X_all = np.arange(len(unc))
    np.random.seed(11)
    xl, xu = O.update_X_L(unc.copy(), X_all, X_L.copy(), 80, zeroRate=0.15, maxconf=None, useMaxConf="False")
    assert np.array_equal(xl, g["sel_X_L_next"]) and np.array_equal(xu, g["sel_X_U_next"])
    np.random.seed(11)
    xl2, xu2 = O.update_X_L(unc.copy(), X_all, X_L.copy(), 80)
    assert np.array_equal(xl2, g["sel2_X_L_next"]) and np.array_equal(xu2, g["sel2_X_U_next"])
    # the deterministic top-k part is a subset of the selected set
    top = O.topk_part(unc, X_all, X_L, 80)
    assert set(top.tolist()) <= set(xl2.tolist())
    vals = np.sort(unc[np.setdiff1d(X_all, X_L)])[-80:]
    assert np.array_equal(np.sort(unc[top]), vals)


def test_empty_and_degenerate_inputs():
    # no detections above obj_thr -> score exactly 0 (Lambda_L2.py:615-616)
    spec = get_spec("tiny_retina_coco")
    batch = SyntheticPool(spec, seed0=20).batch([0])
    for t in batch["cls_scores"]:
        t.zero_()                                    # uniform softmax: p = 1/80 everywhere
    out = O.score_batch(batch, **O.spec_kwargs(spec, ScoringParams(n_samples=4)))
    assert out["image_scores"] == [0]
    assert out["dets"][0].shape == (0, 5) and out["pos_bboxes"][0].shape[1] == 0
    assert not out["level_fg"].any()


@pytest.mark.parametrize("case", RPO_CASES, ids=[c[0] for c in RPO_CASES])
def test_relu_plus_one_detection_route_matches_the_base_head(case):
    """alpha = relu(logits) + 1, score = alpha / (sum alpha + 1e-20) (L_anchor_head.py:401-406): the oracle's
    pre-stage with activation='relu_plus_one' against det_results of the AST-loaded L_AnchorHead._get_bboxes."""
    name, spec_name, gids, pseed = case
    g = _load(name)
    spec = get_spec(spec_name)
    batch = SyntheticPool(spec, seed0=pseed).batch(gids)
    assert bytes.fromhex(batch_checksum(batch)) == g["checksum"].tobytes(), "synthetic inputs drifted"
    kw = O.spec_kwargs(spec)
    pre = O.pre_stage(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"], batch["img_shapes"],
                      batch["scale_factors"], activation="relu_plus_one", **kw)
    for b in range(len(gids)):
        assert np.array_equal(pre["dets"][b].numpy(), g[f"dets_{b}"])
        assert np.array_equal(pre["labels"][b].numpy(), g[f"labels_{b}"])


@pytest.mark.parametrize("case", AVG_CASES, ids=[c[0] for c in AVG_CASES])
def test_entropy_avg_restatement_matches_reference_outputs(case):
    """ComputeAvgUnc + AggregateAvgUnc (Lambda_L2_ReLU.py:446-474, 532-541): same torch seed, same visiting order
    -> the same draws -> identical level means and image scores."""
    name, spec_name, gids, pseed, sseed = case
    g = _load(name)
    spec = get_spec(spec_name)
    batch = SyntheticPool(spec, seed0=pseed).batch(gids)
    assert bytes.fromhex(batch_checksum(batch)) == g["checksum"].tobytes(), "synthetic inputs drifted"
    torch.manual_seed(sseed)
    out = O.score_batch_avg(batch, c_out=spec.c_out)
    levels = np.asarray([[(v if v else np.nan) for v in img] for img in out["nested"]], dtype=np.float64)
    assert np.array_equal(np.isnan(levels), np.isnan(g["level_means"]))
    if np.array_equal(levels[~np.isnan(levels)], g["level_means"][~np.isnan(levels)]):      # torch build reproduced the draws
        assert np.array_equal(np.asarray(out["image_scores"], dtype=np.float64), g["image_scores"])
    else:
        np.testing.assert_allclose(np.asarray(out["image_scores"]), g["image_scores"], rtol=0.3)


def test_mutual_information_restatement_matches_the_reference_baselines():
    """ComputeMI (apis/CalEnsembleUnc.py:166-181) and ComputeMCDropoutMI (apis/CalMCDropoutUnc.py:185-201): the
    oracle's restatement on the seeded member logits against the outputs of the reference's own functions."""
    g = _load("mi_baselines")
    for key, seed, members in (("ensemble", 301, 3), ("mcdropout", 302, 7)):
        x = O.mi_inputs(seed, members)
        digest = hashlib.sha256(b"".join(t.numpy().tobytes() for m in x for t in m)).digest()
        assert bytes(g[f"{key}_checksum"]) == digest
        scores, levels = O.compute_mi(x, 20)
        assert levels.shape == (2, len(O.MI_SHAPES))
        np.testing.assert_array_equal(scores.double().numpy(), g[key])
