"""GPU: the drop-in boundary (aod_meh_hua_b200/dropin.py) behaves like the reference methods it
replaces - same call signatures, return types, fall-through and error behaviour - checked against
the CPU oracle.  mmdet/mmcv are not installed here, so the mixin sits in front of a stub head that
carries exactly the attributes the reference method reads."""
import types

import numpy as np
import pytest
import torch

from aod_meh_hua_b200 import _lib
from aod_meh_hua_b200.dropin import B200ScoringMixin, calculate_uncertainty, update_X_L
from aod_meh_hua_b200.specs import HEAD_RETINA, ScoringParams
from oracle import meh_hua_oracle as O
from tests.helpers import make_batch, run_oracle

pytestmark = pytest.mark.gpu


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class _ReferenceHeadStub:
    """Stands in for Lambda_L2Net / MyLSSDHead: the method the mixin falls through to."""

    def _get_bboxes(self, *args, **kwargs):
        return "reference-route"


def _head(spec):
    class Head(B200ScoringMixin, _ReferenceHeadStub):
        pass

    h = Head()
    h.cls_out_channels = spec.c_out
    h.last_activation = "relu" if spec.head == HEAD_RETINA else "softmax"
    h.test_cfg = _Cfg(nms_pre=spec.nms_pre, min_bbox_size=0, score_thr=spec.score_thr,
                      nms=dict(type="nms", iou_threshold=spec.nms_iou), max_per_img=spec.max_per_img)
    h.bbox_coder = types.SimpleNamespace(means=(0., 0., 0., 0.), stds=spec.target_stds)
    return h


def _cuda(batch):
    dev = "cuda:0"
    return ([t.to(dev) for t in batch["cls_scores"]], [t.to(dev) for t in batch["bbox_preds"]],
            [t.to(dev) for t in batch["L_scores"]], [t.to(dev) for t in batch["anchors"]])


KW = dict(isUnc="Epistemic", uPool="Entropy_NMS", uPool2="objectSum_scaleMax_classSum", isEval=False, showNMS=False,
          saveUnc=False, saveMaxConf=False, clsW=False, scaleUnc=False, score_thr=0.3, iou_thr=0.9, batchIdx=0,
          return_box=False)


@pytest.mark.parametrize("spec_name", ["tiny_retina_coco", "tiny_ssd_voc"])
def test_get_bboxes_route(spec_name):
    spec, batch = make_batch(spec_name, [0, 1])
    out, _ = run_oracle(spec, batch, ScoringParams())
    head = _head(spec)
    cls, reg, lam, anc = _cuda(batch)
    sf = [np.asarray(s, dtype=np.float32) for s in batch["scale_factors"]]
    dets, unc = head._get_bboxes(cls, reg, anc, batch["img_shapes"], sf, None, True, True, L_scores=lam, **KW)
    assert isinstance(unc, list) and len(unc) == 2 and all(isinstance(v, (float, int)) for v in unc)
    for b, (d, l) in enumerate(dets):
        assert d.shape[1] == 5 and l.dtype == torch.int64 and d.is_cuda
        assert np.array_equal(l.cpu().numpy(), out["labels"][b].numpy())
        np.testing.assert_allclose(d.cpu().numpy(), out["dets"][b].numpy(), rtol=1e-5, atol=1e-4)
    # free-running sampler vs the oracle's own Monte-Carlo draw: same estimator, different stream
    np.testing.assert_allclose(unc, out["image_scores"], rtol=0.15, atol=0.05)
    # evaluation route (isEval=True, isUnc=None): plain det_results from the fused decode + NMS
    ev = head._get_bboxes(cls, reg, anc, batch["img_shapes"], sf, None, True, True, isUnc=None, isEval=True)
    assert len(ev) == 2
    for b, (d, l) in enumerate(ev):
        assert np.array_equal(l.cpu().numpy(), out["labels"][b].numpy())
        np.testing.assert_allclose(d.cpu().numpy(), out["dets"][b].numpy(), rtol=1e-5, atol=1e-4)
    head.mehhua_fused_eval = False      # opt-out: every other route falls through to the reference method
    assert head._get_bboxes(cls, reg, anc, batch["img_shapes"], sf, None, True, True, isUnc=None, isEval=True) == "reference-route"
    kw2 = dict(KW, uPool="Entropy_NoNMS")
    assert head._get_bboxes(cls, reg, anc, batch["img_shapes"], sf, None, True, False, L_scores=lam, **kw2) == "reference-route"
    # the Entropy_ALL route (with_nms=False, one of the four scale/class types) is taken over as well
    torch.manual_seed(7)
    want_all = O.score_batch_all(batch, kind="scaleSum_classAvg", **O.spec_kwargs(spec, ScoringParams()))["image_scores"]
    kw3 = dict(KW, uPool="Entropy_ALL", uPool2="scaleSum_classAvg")
    dets_all, unc_all = head._get_bboxes(cls, reg, anc, batch["img_shapes"], sf, None, True, False, L_scores=lam, **kw3)
    assert len(dets_all) == 2 and len(unc_all) == 2
    np.testing.assert_allclose(unc_all, np.asarray(want_all, dtype=np.float64), rtol=0.15, atol=0.03)
    # scaleUnc=True: the third return item is ComputeScaleUnc's nested structure (Lambda_L2.py:377-378)
    d3, u3, sc3 = head._get_bboxes(cls, reg, anc, batch["img_shapes"], sf, None, True, False, L_scores=lam,
                                   **dict(kw3, scaleUnc=True))
    assert len(sc3) == 2 and all(len(img) == spec.num_levels for img in sc3)
    torch.manual_seed(7)
    nested = O.score_batch_all(batch, kind="scaleSum_classAvg", **O.spec_kwargs(spec, ScoringParams()))["nested"]
    for b in range(2):
        for s in range(spec.num_levels):
            assert set(sc3[b][s].keys()) == set(nested[b][s].keys())
            for c, (ale, epi) in sc3[b][s].items():
                assert torch.is_tensor(epi) and epi.dim() == 0
    # and the value the reference's AggregateScaleUnc makes of that structure is the returned score
    np.testing.assert_allclose(O.aggregate_scale_unc(sc3, "scaleSum_classAvg"), u3, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        head._get_bboxes(cls, reg, anc, batch["img_shapes"], sf, None, True, False, L_scores=lam,
                         **dict(KW, uPool="Entropy_ALL", uPool2="objectSum_scaleMax_classSum"))
    with pytest.raises(KeyError):       # spec without an 'object' token, as ExtractAggFunc + AggregateObjScaleUnc
        head._get_bboxes(cls, reg, anc, batch["img_shapes"], sf, None, True, True, L_scores=lam,
                         **dict(KW, uPool2="scaleAvg_classAvg"))


def test_compute_obj_unc_and_aggregate_boundary():
    spec, batch = make_batch("tiny_retina_coco", [0, 1, 2])
    params = ScoringParams(n_samples=4000)
    out, _ = run_oracle(spec, batch, params)
    head = _head(spec)
    head.mehhua_params = params
    cls, reg, lam, anc = _cuda(batch)
    dev = "cuda:0"
    pos = [p.to(dev) for p in out["pos_bboxes"]]
    rows = [r.to(dev) for r in out["lvl_scores"]]
    lams = [l.to(dev) for l in out["lvl_L"]]
    nested = head.ComputeObjUnc(cls, pos, rows, lams, None)
    want = out["nested"]
    assert len(nested) == len(want)
    for img_g, img_w in zip(nested, want):
        assert len(img_g) == len(img_w)
        for og, ow in zip(img_g, img_w):
            for lg, lw in zip(og, ow):
                assert set(lg.keys()) == set(lw.keys())
                for k in lg:
                    ale_g, epi_g = lg[k]
                    ale_w, epi_w = lw[k]
                    assert ale_g.ndim == 0 and epi_g.ndim == 0
                    np.testing.assert_allclose(float(ale_g), float(ale_w), rtol=0.1, atol=0.02)
                    np.testing.assert_allclose(float(epi_g), float(epi_w), rtol=0.25, atol=0.02)
    # AggregateObjScaleUnc on the oracle's own nested table reproduces the oracle's scores
    for agg, clsw in (("objectSum_scaleMax_classSum", False), ("objectAvg_scaleSum_classMax", True)):
        got = head.AggregateObjScaleUnc(want, agg, clsW=clsw)
        np.testing.assert_allclose(got, O.aggregate_obj_scale_unc(want, agg, clsw), rtol=1e-6)
    assert head.AggregateObjScaleUnc([[]], "objectSum_scaleMax_classSum") == [0]


def test_calculate_uncertainty_loop_and_errors():
    spec, batch = make_batch("tiny_retina_coco", [0, 1, 2, 3])
    head = _head(spec)
    cls, reg, lam, anc = _cuda(batch)
    sf = [np.asarray(s, dtype=np.float32) for s in batch["scale_factors"]]

    class Model:
        def eval(self):
            return self

        def __call__(self, return_loss, rescale, isEval, batchIdx, img, img_metas, **kw):
            lo = batchIdx * 2
            r = head._get_bboxes([c[lo:lo + 2] for c in cls], [c[lo:lo + 2] for c in reg], anc,
                                 batch["img_shapes"][lo:lo + 2], sf[lo:lo + 2], None, rescale, True,
                                 L_scores=[c[lo:lo + 2] for c in lam], isEval=isEval, batchIdx=batchIdx, **kw)
            return (r[0], r[1])

    cfg = _Cfg(uncertainty_pool="Entropy_NMS", uncertainty_type="Epistemic", uncertainty_pool2="objectSum_scaleMax_classSum")
    loader = [dict(img=torch.zeros(2), img_metas=[{}, {}]) for _ in range(2)]
    kw = dict(return_box=False, showNMS=False, saveUnc=False, saveMaxConf=False, clsW=False, scaleUnc=False,
              score_thr=0.3, iou_thr=0.9)
    unc = calculate_uncertainty(cfg, Model(), loader, **kw)
    assert len(unc) == 4 and all(torch.is_tensor(u) and u.ndim == 0 and not u.is_cuda for u in unc)
    stacked = torch.stack(unc).numpy()          # what tools/train_RetinaNet.py:242-245 does
    assert stacked.dtype == np.float32 and np.all(stacked >= 0)
    with pytest.raises(KeyError):               # the reference reads kwargs['scaleUnc'] unconditionally
        calculate_uncertainty(cfg, Model(), loader, **{k: v for k, v in kw.items() if k != "scaleUnc"})

    class Broken(Model):
        def __call__(self, *a, **k):
            raise RuntimeError("boom")
    with pytest.raises(RuntimeError):           # never swallowed (cf. apis/test.py:122-128)
        calculate_uncertainty(cfg, Broken(), loader, **kw)


def test_update_X_L_matches_reference_semantics():
    rs = np.random.RandomState(3)
    n = 3000
    unc = rs.permutation(n).astype(np.float32) / n + 0.001      # tie-free, non-zero
    unc[rs.choice(n, 600, replace=False)] = 0.0                 # plus exact zeros (drive zeroRate)
    X_all = np.arange(n)
    X_L = np.sort(rs.choice(n, 150, replace=False))
    for kw in (dict(), dict(zeroRate=0.15, maxconf=None, useMaxConf="False")):
        np.random.seed(11)
        want = O.update_X_L(unc.copy(), X_all, X_L.copy(), 120, **kw)
        np.random.seed(11)
        got = update_X_L(unc.copy(), X_all, X_L.copy(), 120, **kw)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    got = update_X_L(torch.from_numpy(unc), X_all, X_L.copy(), 120)      # tensor input, as the AL script may pass
    assert got[0].shape[0] == 270 and np.all(np.diff(got[0]) > 0)


def test_host_buffer_entry_point_matches_device_path():
    import ctypes as C
    from aod_meh_hua_b200.scoring import Scorer
    spec, batch = make_batch("tiny_retina_coco", [0, 1, 2])
    params = ScoringParams(n_samples=64)
    sc = Scorer(spec, params, max_batch=3, device="cuda:0")
    res = sc.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"], batch["img_shapes"],
                   batch["scale_factors"], image_ids=batch["gids"])
    want = res.image_scores.cpu().numpy().copy()
    lib = _lib.load()
    ctx = C.c_void_p()
    _lib.check(lib.mehhua_host_ctx_create(C.byref(sc.cfg), sc._shape_levels, 3, C.byref(ctx)), "ctx")
    lv = _lib.LevelArray()
    keep = []
    for s in range(spec.num_levels):
        a, b_, c, d = (batch["cls_scores"][s].contiguous(), batch["bbox_preds"][s].contiguous(),
                       batch["L_scores"][s].contiguous(), batch["anchors"][s].contiguous())
        keep += [a, b_, c, d]
        lv[s].logits, lv[s].deltas, lv[s].lam, lv[s].anchors = a.data_ptr(), b_.data_ptr(), c.data_ptr(), d.data_ptr()
        (lv[s].H, lv[s].W), lv[s].A = spec.featmaps[s], spec.num_anchors[s]
    shp = np.asarray([[spec.img_hw[0], spec.img_hw[1]]] * 3, dtype=np.float32)
    sf = np.ones((3, 4), dtype=np.float32)
    ids = np.asarray(batch["gids"], dtype=np.int64)
    out = np.zeros(3, dtype=np.float32)
    st = C.c_uint32(0)
    _lib.check(lib.mehhua_score_batch_host(ctx, lv, 3, shp.ctypes.data, sf.ctypes.data, ids.ctypes.data,
                                           out.ctypes.data, C.byref(st)), "host")
    # same Philox keys (seed, image id, row, object) -> bit-identical scores through either entry point
    assert np.array_equal(out, want) and st.value & _lib.ST_PAIR_OVERFLOW == 0
    # pageable (above) and page-locked callers get the same answer; mehhua_host_pin registers an existing allocation
    big = batch["cls_scores"][0].contiguous()
    assert lib.mehhua_host_is_pinned(big.data_ptr()) == 0
    _lib.check(lib.mehhua_host_pin(big.data_ptr(), big.numel() * 4), "pin")
    assert lib.mehhua_host_is_pinned(big.data_ptr()) == 1
    lv[0].logits = big.data_ptr()
    out2 = np.zeros(3, dtype=np.float32)
    _lib.check(lib.mehhua_score_batch_host(ctx, lv, 3, shp.ctypes.data, sf.ctypes.data, ids.ctypes.data,
                                           out2.ctypes.data, C.byref(st)), "host")
    _lib.check(lib.mehhua_host_unpin(big.data_ptr()), "unpin")
    assert np.array_equal(out2, want)
    lib.mehhua_host_ctx_destroy(ctx)


def test_degenerate_images_score_zero():
    """Uniform logits: no prior above fg_thr, no detection above obj_thr -> score exactly 0.0
    (Lambda_L2.py:615-616); also exercises the dense-tie slow path of the radix select."""
    from aod_meh_hua_b200.scoring import Scorer
    spec, batch = make_batch("tiny_retina_coco", [0, 1])
    for t in batch["cls_scores"]:
        t.zero_()
    sc = Scorer(spec, ScoringParams(n_samples=16), max_batch=2, device="cuda:0")
    res = sc.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"], batch["img_shapes"],
                   batch["scale_factors"], check=False)
    st = sc.read_status()
    assert res.image_scores.cpu().tolist() == [0.0, 0.0]
    assert res.n_det.cpu().tolist() == [0, 0] and res.pair_off[:, -1].cpu().tolist() == [0, 0]
    assert not res.level_fg.cpu().any()
    assert st & _lib.ST_PAIR_OVERFLOW == 0
    idx = res.topk_idx[0, :1000].cpu().numpy()
    assert len(set(idx.tolist())) == 1000


def test_all_equal_keys_take_the_tie_breaking_passes():
    """36 864 identical keys in one level (> the 4096-entry select buffer): the select must order
    exact ties by position and return a valid, deterministic top-1000."""
    from aod_meh_hua_b200.scoring import Scorer
    spec, batch = make_batch("cfg1_retina_r50_512_voc", [0])
    for t in batch["cls_scores"]:
        t.zero_()
    sc = Scorer(spec, ScoringParams(n_samples=8), max_batch=1, device="cuda:0")
    res = sc.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"], batch["img_shapes"],
                   batch["scale_factors"], check=False)
    assert sc.read_status() & _lib.ST_PAIR_OVERFLOW == 0
    assert res.image_scores.cpu().tolist() == [0.0]
    A, HW = spec.num_anchors[0], spec.featmaps[0][0] * spec.featmaps[0][1]
    idx = res.topk_idx[0, :1000].cpu().numpy()
    # exact ties go to the lowest positions of the anchor-major key array: anchor 0, hw = 0..999
    assert np.array_equal(idx, np.arange(1000) * A)
    res2 = sc.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"], batch["img_shapes"],
                    batch["scale_factors"], check=False)
    assert np.array_equal(res2.topk_idx.cpu().numpy(), res.topk_idx.cpu().numpy())


def test_ablation_variant_parameters():
    """Lambda_L2_ReLU-style variants: thresholds from kwargs (score_thr -> object / foreground
    threshold, iou_thr -> cluster IoU) and no lambda scaling (Lambda_L2_noL.py:531)."""
    from aod_meh_hua_b200.scoring import Scorer
    from tests.helpers import check_topk_order, injection_buffers
    spec, batch = make_batch("tiny_retina_coco", [0, 1, 2])
    params = ScoringParams(fg_thr=0.2, obj_thr=0.2, cluster_iou=0.4, use_lambda=False, n_samples=50)
    out, rec = run_oracle(spec, batch, params)
    sc = Scorer(spec, params, max_batch=3, device="cuda:0")
    B = sc.bind(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"], batch["img_shapes"],
                batch["scale_factors"], image_ids=batch["gids"])
    sc.k1(); torch.cuda.synchronize()
    override, swapped = check_topk_order(spec, out, sc.result().topk_idx.cpu().numpy())
    if swapped:
        out, rec = run_oracle(spec, batch, params, topk_override=override)
    sc.nms(); sc.pairs()
    inj, off = injection_buffers(spec, rec, B, sc.device)
    sc.k2(inj, off); sc.hua()
    np.testing.assert_allclose(sc.result().image_scores.cpu().numpy(), np.asarray(out["image_scores"], dtype=np.float32),
                               rtol=1e-5, atol=1e-5)
    # the mixin picks the thresholds up from kwargs when the variant flag is set
    head = _head(spec)
    head.mehhua_thresholds_from_kwargs = True
    head.mehhua_params = ScoringParams(use_lambda=False)
    cls, reg, lam, anc = _cuda(batch)
    sf = [np.asarray(s, dtype=np.float32) for s in batch["scale_factors"]]
    kw = dict(KW, score_thr=0.2, iou_thr=0.4)
    _, unc = head._get_bboxes(cls, reg, anc, batch["img_shapes"], sf, None, True, True, L_scores=lam, **kw)
    want = run_oracle(spec, batch, ScoringParams(fg_thr=0.2, obj_thr=0.2, cluster_iou=0.4, use_lambda=False))[0]["image_scores"]
    np.testing.assert_allclose(unc, want, rtol=0.15, atol=0.05)


@pytest.mark.parametrize("spec_name", ["tiny_retina_coco", "tiny_ssd_voc", "tiny_retina_c12", "tiny_ssd_c7"])
def test_max_conf_fused_into_the_logits_pass(spec_name):
    """getMaxConf (utils/functions.py:467-476): level maxima come out of K1a (Entropy_NMS) and KA1
    (Entropy_ALL); golden = the reference function's own output, generic class counts vs the oracle."""
    import os
    from aod_meh_hua_b200.dropin import max_conf
    from aod_meh_hua_b200.scoring import Scorer
    spec, batch = make_batch(spec_name, [0, 1, 2])
    want_img, want_lvl = O.get_max_conf(batch["cls_scores"], spec.c_out)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kats.npz"))
    if f"maxconf_levels_{spec_name}" in g.files:
        assert np.array_equal(want_lvl.numpy(), g[f"maxconf_levels_{spec_name}"])
        want_lvl = torch.from_numpy(g[f"maxconf_levels_{spec_name}"])
    for mode, agg in (("nms", "objectSum_scaleMax_classSum"), ("all", "scaleSum_classSum")):
        sc = Scorer(spec, ScoringParams(n_samples=8, agg=agg), max_batch=3, mode=mode)
        with pytest.raises(_lib.MehhuaError):
            max_conf(sc)
        sc.save_max_conf(True)
        sc.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"], batch["img_shapes"],
                 batch["scale_factors"])
        per_img, per_lvl = max_conf(sc)
        np.testing.assert_allclose(per_lvl.cpu().numpy(), want_lvl.numpy(), rtol=1e-5)
        np.testing.assert_allclose(per_img, want_lvl.max(dim=1)[0].numpy(), rtol=1e-5)
    # the head method returns it as the third item when saveMaxConf is set (Lambda_L2.py:376-381)
    head = _head(spec)
    cls, reg, lam, anc = _cuda(batch)
    sf = [np.asarray(s, dtype=np.float32) for s in batch["scale_factors"]]
    kw = dict(KW, saveMaxConf=True)
    r = head._get_bboxes(cls, reg, anc, batch["img_shapes"], sf, None, True, True, L_scores=lam, **kw)
    assert len(r) == 3 and isinstance(r[2], list) and len(r[2]) == 3
    np.testing.assert_allclose(r[2], want_img, rtol=1e-5)
    kw_all = dict(kw, uPool="Entropy_ALL", uPool2="scaleAvg_classAvg")
    r = head._get_bboxes(cls, reg, anc, batch["img_shapes"], sf, None, True, False, L_scores=lam, **kw_all)
    assert len(r) == 3
    np.testing.assert_allclose(r[2], want_img, rtol=1e-5)


def test_update_X_L_max_conf_modes_match_the_reference():
    """useMaxConf 'min' / 'max' (active_datasets.py:113-119) against the reference's own outputs."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kats.npz"))
    unc, X_L, maxconf = g["sel_unc"], g["sel_X_L"], g["sel_maxconf"]
    X_all = np.arange(len(unc))
    # the golden pool has tied zero scores only below the selected range: the top part is tie-free
    for mode in ("min", "max"):
        np.random.seed(11)
        xl, xu = update_X_L(unc.copy(), X_all, X_L.copy(), 80, zeroRate=0.15, maxconf=maxconf.tolist(), useMaxConf=mode)
        assert np.array_equal(xl, g[f"sel_{mode}_X_L_next"]) and np.array_equal(xu, g[f"sel_{mode}_X_U_next"])
    np.random.seed(11)
    xl, xu = update_X_L(unc.copy(), X_all, X_L.copy(), 80, zeroRate=0.15, maxconf=None, useMaxConf="False")
    assert np.array_equal(xl, g["sel_X_L_next"]) and np.array_equal(xu, g["sel_X_U_next"])


def test_scoring_stage_writes_the_cycle_files(tmp_path):
    """tools/train_RetinaNet.py:231-251: calculate_uncertainty -> update_X_L -> X_L/X_U/Unc .npy,
    readable back through ResumeCycle; Entropy_ALL and Random pools use the same loop."""
    from aod_meh_hua_b200.dropin import ResumeCycle, scoring_stage
    spec, batch = make_batch("tiny_retina_coco", [0, 1, 2, 3, 4, 5])
    head = _head(spec)
    cls, reg, lam, anc = _cuda(batch)
    sf = [np.asarray(s, dtype=np.float32) for s in batch["scale_factors"]]

    class Model:
        def eval(self):
            return self

        def __call__(self, return_loss, rescale, isEval, batchIdx, img, img_metas, **kw):
            lo = batchIdx * 2
            return head._get_bboxes([c[lo:lo + 2] for c in cls], [c[lo:lo + 2] for c in reg], anc,
                                    batch["img_shapes"][lo:lo + 2], sf[lo:lo + 2], None, rescale,
                                    kw.get("uPool") != "Entropy_ALL", L_scores=[c[lo:lo + 2] for c in lam],
                                    isEval=isEval, batchIdx=batchIdx, **kw)

    class Loader(list):
        dataset = list(range(6))

    loader = Loader(dict(img=torch.zeros(2), img_metas=[{}, {}]) for _ in range(3))
    X_all, X_L = np.arange(6), np.array([1])
    for pool, up2, smc in (("Entropy_NMS", "objectSum_scaleMax_classSum", False),
                           ("Entropy_NMS", "objectSum_scaleMax_classSum", True),
                           ("Entropy_ALL", "scaleSum_classSum", False)):
        cfg = _Cfg(uncertainty_pool=pool, uncertainty_type="Epistemic", uncertainty_pool2=up2, X_S_size=2,
                   work_dir=str(tmp_path))
        np.random.seed(5)
        xl, xu, unc = scoring_stage(cfg, Model(), loader, X_all, X_L, cycle=0, saveMaxConf=smc,
                                    useMaxConf="min" if smc else "False")
        assert unc.dtype == np.float32 and unc.shape == (6,) and np.all(unc >= 0)
        assert len(xl) == 3 and 1 in xl and np.all(np.diff(xl) > 0)
        rl, ru = ResumeCycle(cfg, 1, 1)
        assert np.array_equal(rl, xl) and np.array_equal(ru, xu)
        assert np.array_equal(np.load(tmp_path / "Unc_1.npy"), unc)
        assert ResumeCycle(cfg, 0, 1) == (False, False)
    cfg = _Cfg(uncertainty_pool="Random", uncertainty_type="Epistemic", uncertainty_pool2="x", X_S_size=2, work_dir=str(tmp_path))
    perm = calculate_uncertainty(cfg, Model(), loader, scaleUnc=False)
    assert sorted(perm.tolist()) == list(range(6))


@pytest.mark.parametrize("B,max_batch,explicit_ids", [(9, 9, True), (10, 12, True), (11, 16, False)])
def test_host_buffer_chunked_upload_matches_device_path(B, max_batch, explicit_ids):
    """mehhua_score_batch_host uploads a batch in up to four image chunks and scores chunk j while
    chunk j+1 is on the link: scores (and the default Philox ids = position in the BATCH) must not
    depend on the chunking, also when B < max_batch and the last chunk is ragged."""
    import ctypes as C
    from aod_meh_hua_b200.scoring import Scorer
    gids = list(range(B))
    spec, batch = make_batch("tiny_retina_coco", gids)
    params = ScoringParams(n_samples=32)
    sc = Scorer(spec, params, max_batch=B, device="cuda:0")
    res = sc.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"], batch["img_shapes"],
                   batch["scale_factors"], image_ids=gids)       # ids 0..B-1 = the host entry's default
    want = res.image_scores.cpu().numpy().copy()
    lib = _lib.load()
    ctx = C.c_void_p()
    _lib.check(lib.mehhua_host_ctx_create(C.byref(sc.cfg), sc._shape_levels, max_batch, C.byref(ctx)), "ctx")
    lv = _lib.LevelArray()
    keep = []
    for s in range(spec.num_levels):
        a, b_, c, d = (batch["cls_scores"][s].contiguous().pin_memory(), batch["bbox_preds"][s].contiguous().pin_memory(),
                       batch["L_scores"][s].contiguous().pin_memory(), batch["anchors"][s].contiguous())
        keep += [a, b_, c, d]
        lv[s].logits, lv[s].deltas, lv[s].lam, lv[s].anchors = a.data_ptr(), b_.data_ptr(), c.data_ptr(), d.data_ptr()
        (lv[s].H, lv[s].W), lv[s].A = spec.featmaps[s], spec.num_anchors[s]
    shp = np.asarray([[spec.img_hw[0], spec.img_hw[1]]] * B, dtype=np.float32)
    sf = np.ones((B, 4), dtype=np.float32)
    ids = np.arange(B, dtype=np.int64)
    st = C.c_uint32(0)
    for rep in range(2):                                 # second call: anchors already resident
        out = np.zeros(B, dtype=np.float32)
        if rep == 1:
            for s in range(spec.num_levels):
                lv[s].anchors = None
        _lib.check(lib.mehhua_score_batch_host(ctx, lv, B, shp.ctypes.data, sf.ctypes.data,
                                               ids.ctypes.data if explicit_ids else None, out.ctypes.data, C.byref(st)), "host")
        assert np.array_equal(out, want), (rep, out, want)
        assert st.value & (_lib.ST_PAIR_OVERFLOW | _lib.ST_BAD_ALPHA) == 0
    lib.mehhua_host_ctx_destroy(ctx)
