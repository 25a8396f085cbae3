"""GPU parity: every stage of the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Integer decisions (top-k membership and order, level-FG flags, NMS keep list,
pair list, class keys) are compared exactly; floats within rel 1e-5 (north_star tolerance)."""
import dataclasses

import numpy as np
import pytest
import torch

from aod_meh_hua_b200.scoring import Scorer
from aod_meh_hua_b200.specs import HEAD_RETINA, ScoringParams
from tests.helpers import assert_epi_close, check_topk_order, injection_buffers, make_batch, oracle_pairs, run_oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-5

CASES = [
    ("tiny_retina_voc", [0, 1], (1.0, 1.0, 1.0, 1.0)),
    ("tiny_retina_coco", [0, 1, 2], (1.0, 1.0, 1.0, 1.0)),
    ("tiny_ssd_voc", [0, 1], (1.0, 1.0, 1.0, 1.0)),
    ("tiny_retina_coco", [3, 4], (1.07, 0.94, 1.07, 0.94)),
    ("tiny_retina_c12", [0, 1, 2], (1.0, 1.0, 1.0, 1.0)),       # generic (runtime class count) kernels
    ("tiny_ssd_c7", [0, 1], (1.0, 1.0, 1.0, 1.0)),
]


def _run(spec_name, gids, sf, params=None, spec=None):
    if spec is None:
        spec, batch = make_batch(spec_name, gids, scale_factor=sf)
    else:
        from aod_meh_hua_b200.synth import SyntheticPool
        batch = SyntheticPool(spec, seed0=20, device="cpu", scale_factor=sf).batch(list(gids))
    params = params or ScoringParams()
    out, rec = run_oracle(spec, batch, params)
    sc = Scorer(spec, params, max_batch=len(gids), device="cuda:0")
    B = sc.bind(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
                batch["img_shapes"], batch["scale_factors"], image_ids=batch["gids"])
    sc.k1()
    torch.cuda.synchronize()
    # stage isolation: near-tied neighbours in the top-k order may swap (last-ulp softmax rounding);
    # the index set must be identical, and the later oracle stages then run in the kernel's row order
    override, swapped = check_topk_order(spec, out, sc.result().topk_idx.cpu().numpy())
    if swapped:
        out, rec = run_oracle(spec, batch, params, topk_override=override)
    sc.nms()
    sc.pairs()
    inj, off = injection_buffers(spec, rec, B, sc.device)
    sc.k2(inj, off)
    sc.hua()
    torch.cuda.synchronize()
    st = sc.check_status()
    res = sc.result()
    return spec, batch, out, rec, res, st


def _random_specs():
    """Seeded random detector geometries: odd image sizes (ragged last tiles, 1-row levels), class
    counts with and without a template instantiation, small nms_pre so that sparse-gather, dense-rescan
    and keep-everything levels all occur, shifted thresholds."""
    from aod_meh_hua_b200.specs import retina_spec, ssd_spec
    rs = np.random.RandomState(77)
    out = []
    for i in range(8):
        h, w = int(rs.randint(40, 200)), int(rs.randint(40, 230))
        c = int(rs.choice([2, 3, 5, 20, 33, 80]))
        spec = retina_spec(f"rand_retina_{i}", h, w, c, nms_pre=int(rs.choice([17, 64, 200, 1000])),
                           gt_range=(1, 4))
        spec = dataclasses.replace(spec, score_thr=float(rs.choice([0.02, 0.05, 0.2])),
                                   max_per_img=int(rs.choice([5, 30, 100])), nms_iou=float(rs.choice([0.3, 0.5, 0.7])))
        out.append((spec, [i, i + 1, i + 2][: int(rs.randint(1, 4))],
                    (1.0, 1.0, 1.0, 1.0) if i % 2 else (1.31, 0.77, 1.31, 0.77)))
    for i in range(3):
        spec = ssd_spec(f"rand_ssd_{i}", 300, int(rs.choice([1, 4, 20])), nms_pre=int(rs.choice([30, 150, 1000])),
                        gt_range=(1, 4))
        spec = dataclasses.replace(spec, max_per_img=int(rs.choice([10, 200])))
        out.append((spec, [i, i + 5], (1.0, 1.0, 1.0, 1.0)))
    return out


@pytest.mark.parametrize("spec,gids,sf", _random_specs(), ids=lambda v: getattr(v, "name", None))
def test_stagewise_parity_random_geometry(spec, gids, sf):
    _check_stagewise(*_run(None, gids, sf, spec=spec), gids)


@pytest.mark.parametrize("spec_name,gids,sf", CASES)
def test_stagewise_parity(spec_name, gids, sf):
    _check_stagewise(*_run(spec_name, gids, sf), gids)


def _check_stagewise(spec, batch, out, rec, res, st, gids):
    B, S, C = len(gids), spec.num_levels, spec.c_out
    koff = np.concatenate([[0], np.cumsum(spec.level_k)])

    # --- K1: per-level top-k index order, rows, lambda, boxes, level-FG flags
    idx = res.topk_idx.cpu().numpy()
    for s in range(S):
        want = out["lvl_idx"][s].numpy()
        got = idx[:, koff[s]:koff[s + 1]]
        assert np.array_equal(got, want), f"level {s}: top-k rows differ from the (order-isolated) oracle"
    np.testing.assert_allclose(res.score_rows.cpu().numpy(), out["scores"].numpy(), rtol=RTOL, atol=1e-9)
    np.testing.assert_array_equal(res.lam_rows.cpu().numpy(), torch.cat(out["lvl_L"], dim=1).numpy())
    np.testing.assert_allclose(res.boxes.cpu().numpy(), out["boxes"].numpy(), rtol=RTOL, atol=1e-4)
    np.testing.assert_array_equal(res.level_fg.cpu().numpy().astype(bool), out["level_fg"])
    np.testing.assert_array_equal(res.row_argmax.cpu().numpy(), out["scores"].argmax(dim=2).numpy())

    # --- K3a: detections (kept set, order, labels), object count
    n_det = res.n_det.cpu().numpy()
    n_obj = res.n_obj.cpu().numpy()
    for b in range(B):
        d_want, l_want = out["dets"][b].numpy(), out["labels"][b].numpy()
        assert n_det[b] == len(d_want)
        assert np.array_equal(res.det_flat[b, :n_det[b]].cpu().numpy(), out["det_flat"][b].numpy())
        assert np.array_equal(res.det_labels[b, :n_det[b]].cpu().numpy(), l_want)
        np.testing.assert_allclose(res.dets[b, :n_det[b]].cpu().numpy(), d_want, rtol=RTOL, atol=1e-4)
        assert n_obj[b] == int((d_want[:, 4] > 0.3).sum())

    # --- K3b: ordered pair list, class keys, lambda means
    poff = res.pair_off.cpu().numpy()
    for b in range(B):
        row, obj, cls, tot, ale, epi, recs = oracle_pairs(out, b)
        n = poff[b, S]
        assert n == len(row)
        assert np.array_equal(res.pair_row[b, :n].cpu().numpy(), row)
        assert np.array_equal(res.pair_obj[b, :n].cpu().numpy(), obj)
        assert np.array_equal(res.pair_cls[b, :n].cpu().numpy(), cls)
        # --- K2 with the oracle's samples injected
        unc = res.pair_unc[b, :n].cpu().numpy()
        np.testing.assert_allclose(unc[:, 0], tot, rtol=RTOL, atol=2e-6)
        np.testing.assert_allclose(unc[:, 1], ale, rtol=RTOL, atol=2e-6)
        assert_epi_close(unc[:, 2], tot, ale, epi)

    # --- K3c: image scores
    np.testing.assert_allclose(res.image_scores.cpu().numpy(),
                               np.asarray(out["image_scores"], dtype=np.float32), rtol=RTOL, atol=1e-5)


@pytest.mark.parametrize("spec_name", ["tiny_retina_voc", "tiny_ssd_voc", "tiny_retina_coco"])
def test_integer_outputs_are_bit_exact_on_guard_banded_data(spec_name):
    """SURVEY 8d / VERDICT r1 weak #3: on inputs whose every threshold decision keeps a 1e-5 relative margin
    (tests/helpers.make_guarded_batch, margins from the oracle's decision_margins) the integer outputs agree
    with the oracle EXACTLY: top-k set (and order wherever neighbouring keys are 2e-6 apart), level flags, class
    keys, NMS keep list AND order, object count, the ordered pair list - no tolerance, no skipped image."""
    from tests.helpers import GUARD, GUARD_ORDER, make_guarded_batch
    spec, batch, margins = make_guarded_batch(spec_name, [0, 1, 2])
    assert min(v for k, v in margins.items() if k not in ("topk_adjacent", "nms_adjacent")) >= GUARD, margins
    assert margins["nms_adjacent"] >= GUARD_ORDER
    params = ScoringParams()
    out, rec = run_oracle(spec, batch, params)
    sc = Scorer(spec, params, max_batch=3, device="cuda:0")
    B = sc.bind(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
                batch["img_shapes"], batch["scale_factors"], image_ids=batch["gids"])
    sc.k1()
    torch.cuda.synchronize()
    override, swapped = check_topk_order(spec, out, sc.result().topk_idx.cpu().numpy())     # set exact; order modulo 2e-6 ties
    if swapped:
        out, rec = run_oracle(spec, batch, params, topk_override=override)
    sc.nms()
    sc.pairs()
    torch.cuda.synchronize()
    res = sc.result()
    S = spec.num_levels
    assert np.array_equal(res.level_fg.cpu().numpy().astype(bool), out["level_fg"])
    assert np.array_equal(res.row_argmax.cpu().numpy(), out["scores"].argmax(dim=2).numpy())
    n_det, n_obj, poff = res.n_det.cpu().numpy(), res.n_obj.cpu().numpy(), res.pair_off.cpu().numpy()
    for b in range(B):
        assert n_det[b] == len(out["dets"][b])
        assert np.array_equal(res.det_flat[b, :n_det[b]].cpu().numpy(), out["det_flat"][b].numpy())
        assert np.array_equal(res.det_labels[b, :n_det[b]].cpu().numpy(), out["labels"][b].numpy())
        assert n_obj[b] == int((out["dets"][b][:, 4] > 0.3).sum())
        row, obj, cls, *_ = oracle_pairs(out, b)
        n = poff[b, S]
        assert n == len(row)
        assert np.array_equal(res.pair_row[b, :n].cpu().numpy(), row)
        assert np.array_equal(res.pair_obj[b, :n].cpu().numpy(), obj)
        assert np.array_equal(res.pair_cls[b, :n].cpu().numpy(), cls)


@pytest.mark.parametrize("agg", ["objectSum_scaleMax_classSum", "objectAvg_scaleSum_classMax",
                                 "objectMax_scaleAvg_classAvg"])
def test_aggregation_matrix(agg):
    params = ScoringParams(agg=agg, cls_w=(agg != "objectSum_scaleMax_classSum"))
    spec, batch, out, rec, res, st = _run("tiny_retina_coco", [0, 1, 2], (1.0, 1.0, 1.0, 1.0), params)
    np.testing.assert_allclose(res.image_scores.cpu().numpy(),
                               np.asarray(out["image_scores"], dtype=np.float32), rtol=RTOL, atol=1e-5)


def test_free_running_sampler_matches_closed_form():
    """Philox/Marsaglia-Tsang sampler: per-pair total / aleatoric against the Dirichlet closed forms
    (SURVEY 7 'sampling parity').  MC error at T samples is O(1/sqrt(T)); use a large T."""
    from oracle import meh_hua_oracle as O
    params = ScoringParams(n_samples=20000)
    spec, batch = make_batch("tiny_retina_coco", [0, 1, 2])
    oracle_params = ScoringParams(n_samples=8)
    out, rec = run_oracle(spec, batch, oracle_params)
    sc = Scorer(spec, params, max_batch=3, device="cuda:0")
    res = sc.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
                   batch["img_shapes"], batch["scale_factors"], image_ids=batch["gids"])
    torch.cuda.synchronize()
    poff = res.pair_off.cpu().numpy()
    S = spec.num_levels
    checked = 0
    for b in range(3):
        row, obj, cls, _, _, _, recs = oracle_pairs(out, b)
        n = poff[b, S]
        assert n == len(row)
        if n == 0:
            continue
        alpha = np.concatenate([r["alpha"] for r in recs]).astype(np.float64)
        h, e_ent, epi_inf = O.dirichlet_expectations(alpha)
        unc = res.pair_unc[b, :n].cpu().numpy().astype(np.float64)
        # aleatoric is an unbiased MC mean; total is H(mean) with O(C/T) bias
        np.testing.assert_allclose(unc[:, 1], e_ent, rtol=0.05, atol=5e-3)
        np.testing.assert_allclose(unc[:, 0], h, rtol=0.05, atol=5e-3)
        checked += n
    assert checked > 0


def test_philox_known_answers():
    import ctypes as C
    from aod_meh_hua_b200 import _lib
    lib = _lib.load()
    # Random123 kat_vectors, philox4x32 with 10 rounds and with 7 (the variant K2's sampler runs)
    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8),
         (0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd),
         (0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1), (0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a)),
    ]
    for ctr, key, want10, want7 in kats:
        out = (C.c_uint32 * 8)()
        _lib.check(lib.mehhua_debug_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out), "philox")
        assert tuple(out)[:4] == want10 and tuple(out)[4:] == want7


def test_pool_topk_matches_stable_argsort():
    from aod_meh_hua_b200.scoring import pool_topk
    g = torch.Generator().manual_seed(5)
    n = 100_000
    scores = torch.rand(n, generator=g)
    scores[torch.rand(n, generator=g) < 0.3] = 0.0          # many exact ties at zero
    scores[1000:1010] = scores[5]                           # a tie group above zero
    mask = (torch.rand(n, generator=g) < 0.8)
    for k in (1, 413, 2500, 9000):
        got = pool_topk(scores.cuda(), k, mask.cuda()).cpu().numpy()
        cand = np.nonzero(mask.numpy())[0]
        order = np.argsort(scores.numpy()[cand], kind="stable")
        want = cand[order[-k:]][::-1]
        assert np.array_equal(got, want)
    # fewer candidates than k
    few = torch.zeros(n, dtype=torch.bool)
    few[[3, 77, 9000]] = True
    got = pool_topk(scores.cuda(), 10, few.cuda()).cpu().numpy()
    assert sorted(got.tolist()) == [3, 77, 9000]


def test_sampler_moments_across_alpha_regimes():
    """K2's gamma samplers (Ahrens-Dieter GS for alpha < 1, Marsaglia-Tsang for alpha >= 1) against
    the Dirichlet closed forms, on hand-picked alpha rows that straddle the algorithm boundary and
    reach the tiny-alpha regime of softmax rows.  T = 200k samples -> MC error ~1e-3."""
    from aod_meh_hua_b200.scoring import pair_uncertainty
    from oracle import meh_hua_oracle as O
    rows = [
        [0.5, 0.5, 0.5, 0.5, 0.5, 0.5],
        [5.0, 0.01, 0.001, 2.0, 0.3, 0.7],
        [0.9, 0.95, 0.999, 1.0, 1.001, 1.05],
        [1e-4, 1e-3, 30.0, 1e-5, 1e-2, 1e-6],
        [0.3, 0.3, 0.3, 0.3, 0.3, 0.3],
        [12.0, 7.0, 3.0, 1.5, 1.0, 25.0],
        [0.05, 0.02, 0.08, 0.6, 0.11, 0.04],
    ]
    alpha = torch.tensor(rows, dtype=torch.float32, device="cuda:0")
    P = alpha.shape[0]
    params = ScoringParams(n_samples=200_000, use_lambda=False)
    unc = pair_uncertainty(alpha, torch.ones(P, device="cuda:0"), torch.arange(P), torch.zeros(P, dtype=torch.long),
                           params, seed_ids=(3, 1)).cpu().numpy().astype(np.float64)
    h, e_ent, epi = O.dirichlet_expectations(np.asarray(rows, dtype=np.float64))
    np.testing.assert_allclose(unc[:, 0], h, rtol=5e-3, atol=2e-3)        # total = H(mean_t x)
    np.testing.assert_allclose(unc[:, 1], e_ent, rtol=5e-3, atol=2e-3)    # aleatoric = E[H(x)]
    np.testing.assert_allclose(unc[:, 2], epi, rtol=2e-2, atol=3e-3)
    # a second seed gives a different stream but the same moments
    unc2 = pair_uncertainty(alpha, torch.ones(P, device="cuda:0"), torch.arange(P), torch.zeros(P, dtype=torch.long),
                            params, seed_ids=(4, 1)).cpu().numpy().astype(np.float64)
    assert not np.array_equal(unc, unc2)
    np.testing.assert_allclose(unc2[:, 1], e_ent, rtol=5e-3, atol=2e-3)


@pytest.mark.parametrize("spec_name,kind", [("tiny_retina_coco", "scaleAvg_classAvg"), ("tiny_ssd_voc", "scaleSum_classSum"),
                                            ("tiny_retina_c12", "scaleSum_classSum"), ("tiny_ssd_c7", "scaleAvg_classAvg"),
                                            ("tiny_retina_voc", "scaleSum_classAvg"), ("tiny_retina_coco", "scaleAvg_classSum")])
def test_entropy_all_mode_parity(spec_name, kind):
    """Entropy_ALL route (ComputeScaleUnc + AggregateScaleUnc): foreground prior lists, class keys,
    alpha inputs, per-prior uncertainties under injection and image scores against the oracle."""
    from oracle import meh_hua_oracle as O
    from tests.helpers import Recorder
    spec, batch = make_batch(spec_name, [0, 1, 2])
    params = ScoringParams(agg=kind)
    rec = Recorder(77)
    out = O.score_batch_all(batch, kind=kind, sampler=rec, **O.spec_kwargs(spec, ScoringParams()))
    sc = Scorer(spec, params, max_batch=3, device="cuda:0", mode="all")
    B = sc.bind(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
                batch["img_shapes"], batch["scale_factors"], image_ids=batch["gids"])
    sc.all_rows()
    inj, off = injection_buffers(spec, rec, B, sc.device)
    sc.k2(inj, off)
    sc.hua()
    torch.cuda.synchronize()
    assert sc.check_status() & 1 == 0
    res = sc.result()
    S = spec.num_levels
    poff = res.pair_off.cpu().numpy()
    for b in range(B):
        recs = sorted([r for r in out["flat"] if r["image"] == b], key=lambda r: r["level"])
        n = poff[b, S]
        assert n == sum(len(r["prior"]) for r in recs)
        for r in recs:
            a, e = poff[b, r["level"]], poff[b, r["level"] + 1]
            assert np.array_equal(res.topk_idx[b, a:e].cpu().numpy(), r["prior"])
            assert np.array_equal(res.pair_cls[b, a:e].cpu().numpy(), r["cls"])
            # alpha = p * lambda' with lambda' from the mean over ALL priors of the level
            lam = res.lam_rows[b, a:e].cpu().numpy()
            lam_p = res.lam_mean[b, r["level"]].item() / (lam + np.float32(1e-7)) * np.float32(25.0)
            alpha = res.score_rows[b, a:e].cpu().numpy() * lam_p[:, None]
            np.testing.assert_allclose(alpha, r["alpha"], rtol=2e-5, atol=1e-12)
            unc = res.pair_unc[b, a:e].cpu().numpy()
            np.testing.assert_allclose(unc[:, 0], r["total"], rtol=RTOL, atol=2e-6)
            np.testing.assert_allclose(unc[:, 1], r["ale"], rtol=RTOL, atol=2e-6)
            assert_epi_close(unc[:, 2], r["total"], r["ale"], r["epi"])
    np.testing.assert_allclose(res.image_scores.cpu().numpy(), np.asarray(out["image_scores"], dtype=np.float64),
                               rtol=RTOL, atol=1e-5)
    # free-running path end to end
    res2 = sc.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"], batch["img_shapes"],
                    batch["scale_factors"], image_ids=batch["gids"])
    np.testing.assert_allclose(res2.image_scores.cpu().numpy(), np.asarray(out["image_scores"], dtype=np.float64),
                               rtol=0.2, atol=0.03)


def _check_all_mode(spec, batch, out, rec, sc, B, lam_alpha=None):
    """Shared body: foreground prior lists in prior order, class keys, per-prior uncertainties under
    injection against the oracle's flat records."""
    S = spec.num_levels
    res = sc.result()
    poff = res.pair_off.cpu().numpy()
    for b in range(B):
        recs = sorted([r for r in out["flat"] if r["image"] == b], key=lambda r: r["level"])
        n = poff[b, S]
        assert n == sum(len(r["prior"]) for r in recs)
        for r in recs:
            a, e = poff[b, r["level"]], poff[b, r["level"] + 1]
            assert np.array_equal(res.topk_idx[b, a:e].cpu().numpy(), r["prior"])
            unc = res.pair_unc[b, a:e].cpu().numpy()
            np.testing.assert_allclose(unc[:, 0], r["total"], rtol=RTOL, atol=2e-6)
            np.testing.assert_allclose(unc[:, 1], r["ale"], rtol=RTOL, atol=2e-6)
            assert_epi_close(unc[:, 2], r["total"], r["ale"], r["epi"])
    return res


def test_entropy_all_handles_any_number_of_foreground_priors():
    """round 1 capped the Entropy_ALL route at 8192 / 16384 foreground priors per image
    (a shared-memory sort).  Here more than half of a 49 104-prior image is foreground (the head has no
    background class, so `max softmax > 0.3` can hold anywhere): 25 000+ rows per image, T = 8 samples
    injected, every per-prior value and the image score against the oracle; the group table (the scaleUnc
    return item) against the oracle's nested dicts."""
    from oracle import meh_hua_oracle as O
    from tests.helpers import Recorder
    spec, batch = make_batch("cfg1_retina_r50_512_voc", [0, 1])
    g = torch.Generator().manual_seed(3)
    for s, t in enumerate(batch["cls_scores"]):            # push one class of ~55 % of the priors up
        Bn, ch, h, w = t.shape
        a = ch // spec.c_out
        v = t.view(Bn, a, spec.c_out, h, w)
        hit = torch.rand(Bn, a, h, w, generator=g) < 0.55
        cls = torch.randint(0, spec.c_out, (Bn, a, h, w), generator=g)
        v.scatter_add_(2, cls.unsqueeze(2), (hit.float() * 5.0).unsqueeze(2))
    params = ScoringParams(agg="scaleSum_classAvg", n_samples=8)
    rec = Recorder(5)
    out = O.score_batch_all(batch, kind=params.agg, sampler=rec, **O.spec_kwargs(spec, params))
    n_fg = [sum(len(r["prior"]) for r in out["flat"] if r["image"] == b) for b in range(2)]
    assert min(n_fg) > 20000
    sc = Scorer(spec, params, max_batch=2, device="cuda:0", mode="all", pair_cap=spec.num_priors)
    B = sc.bind(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
                batch["img_shapes"], batch["scale_factors"], image_ids=batch["gids"])
    sc.all_rows()
    inj, off = injection_buffers(spec, rec, B, sc.device)
    sc.k2(inj, off)
    sc.hua()
    torch.cuda.synchronize()
    assert sc.check_status() & 1 == 0
    res = _check_all_mode(spec, batch, out, rec, sc, B)
    np.testing.assert_allclose(res.image_scores.cpu().numpy(), np.asarray(out["image_scores"], dtype=np.float64),
                               rtol=RTOL, atol=1e-5)
    grp = res.group_unc.cpu().numpy()
    for b in range(B):
        for s in range(spec.num_levels):
            want = out["nested"][b][s]
            assert set(np.nonzero(grp[b, s, :, 0])[0].tolist()) == {int(c) for c in want}
            for c, (ale, epi) in want.items():
                np.testing.assert_allclose(grp[b, s, int(c), 1:], [float(ale), float(epi)], rtol=2e-5, atol=2e-6)
    # a row buffer that is too small is reported, not silently truncated into a wrong score
    small = Scorer(spec, params, max_batch=2, device="cuda:0", mode="all", pair_cap=4096)
    small.bind(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
               batch["img_shapes"], batch["scale_factors"], image_ids=batch["gids"])
    small.all_rows()
    with pytest.raises(Exception, match="pair_cap"):
        small.check_status()


@pytest.mark.parametrize("spec_name", ["tiny_retina_voc", "tiny_retina_coco", "tiny_retina_c12", "cfg1_retina_r50_512_voc"])
def test_relu_plus_one_scores_and_detections(spec_name):
    """MEHHUA_ACT_RELU_PLUS_ONE (the base head's evidential form, L_anchor_head.py:401-406): K1 rows, top-k,
    boxes and the K3a detections against the oracle's pre-stage with the same activation (typed and generic
    class counts, capture and gather levels)."""
    from oracle import meh_hua_oracle as O
    spec, batch = make_batch(spec_name, [0, 1])
    params = ScoringParams(activation="relu_plus_one")
    pre = O.pre_stage(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"], batch["img_shapes"],
                      batch["scale_factors"], activation="relu_plus_one", **O.spec_kwargs(spec))
    sc = Scorer(spec, params, max_batch=2, device="cuda:0")
    sc.bind(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"], batch["img_shapes"],
            batch["scale_factors"], image_ids=batch["gids"])
    sc.k1()
    sc.nms()
    torch.cuda.synchronize()
    res = sc.result()
    override, _ = check_topk_order(spec, pre, res.topk_idx.cpu().numpy())
    pre = O.pre_stage(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"], batch["img_shapes"],
                      batch["scale_factors"], activation="relu_plus_one", topk_override=override, **O.spec_kwargs(spec))
    np.testing.assert_allclose(res.score_rows.cpu().numpy(), pre["scores"].numpy(), rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(res.boxes.cpu().numpy(), pre["boxes"].numpy(), rtol=RTOL, atol=1e-4)
    n_det = res.n_det.cpu().numpy()
    for b in range(2):
        assert n_det[b] == len(pre["dets"][b])
        assert np.array_equal(res.det_flat[b, :n_det[b]].cpu().numpy(), pre["det_flat"][b].numpy())
        assert np.array_equal(res.det_labels[b, :n_det[b]].cpu().numpy(), pre["labels"][b].numpy())
        np.testing.assert_allclose(res.dets[b, :n_det[b]].cpu().numpy(), pre["dets"][b].numpy(), rtol=RTOL, atol=1e-4)


@pytest.mark.parametrize("spec_name", ["tiny_retina_voc", "tiny_retina_c12", "cfg1_retina_r50_512_voc"])
def test_entropy_avg_mode_parity(spec_name):
    """Entropy_Avg route of the ablation heads (ComputeAvgUnc + AggregateAvgUnc, Lambda_L2_ReLU.py:446-474,
    532-541): relu rows, FG = max(relu / (sum + 1e-9)) > 0.3, alpha = relu * lambda' (zeros allowed), T = 50,
    per-level pooled mean, mean over levels - the oracle's draws injected."""
    from oracle import meh_hua_oracle as O
    from tests.helpers import Recorder
    spec, batch = make_batch(spec_name, [0, 1, 2])
    params = ScoringParams(agg="Entropy_Avg", n_samples=50, activation="relu")
    rec = Recorder(9, sampler=O.zero_alpha_sampler)
    out = O.score_batch_avg(batch, c_out=spec.c_out, T=50, sampler=rec)
    assert sum(len(r["prior"]) for r in out["flat"]) > 0
    sc = Scorer(spec, params, max_batch=3, device="cuda:0", mode="all")
    B = sc.bind(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
                batch["img_shapes"], batch["scale_factors"], image_ids=batch["gids"])
    sc.all_rows()
    inj, off = injection_buffers(spec, rec, B, sc.device)
    sc.k2(inj, off)
    sc.hua()
    torch.cuda.synchronize()
    assert sc.check_status() & 1 == 0
    res = _check_all_mode(spec, batch, out, rec, sc, B)
    want = np.asarray(out["image_scores"], dtype=np.float64)
    got = res.image_scores.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_allclose(got[~np.isnan(want)], want[~np.isnan(want)], rtol=RTOL, atol=1e-6)
    # rows are relu(logits): alpha = row * lambda' matches the oracle's alpha, zeros included
    S = spec.num_levels
    poff = res.pair_off.cpu().numpy()
    for r in out["flat"]:
        b, a, e = r["image"], poff[r["image"], r["level"]], poff[r["image"], r["level"] + 1]
        lam = res.lam_rows[b, a:e].cpu().numpy()
        lam_p = res.lam_mean[b, r["level"]].item() / (lam + np.float32(1e-7)) * np.float32(25.0)
        np.testing.assert_allclose(res.score_rows[b, a:e].cpu().numpy() * lam_p[:, None], r["alpha"], rtol=2e-5, atol=1e-12)
    # free-running sampler end to end (T = 50: a noisy estimator on both sides)
    res2 = sc.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"], batch["img_shapes"],
                    batch["scale_factors"], image_ids=batch["gids"])
    g2 = res2.image_scores.cpu().numpy()
    np.testing.assert_allclose(g2[~np.isnan(want)], want[~np.isnan(want)], rtol=0.25, atol=0.03)


@pytest.mark.parametrize("spec_name,gid", [("cfg3_retina_r50_800x1344_coco", 7), ("cfg4_ssd512_coco", 3),
                                           ("cfg1_retina_r50_512_voc", 5), ("cfg2_ssd300_voc", 11),
                                           ("cfg3p_retina_r50_800x800_coco", 2), ("cfg5_retina_r101_1344_coco", 1)])
def test_full_size_configs_end_to_end(spec_name, gid):
    """BASELINE.json shapes at full size (one image each, the oracle needs seconds): the whole path
    with injected samples against the oracle, plus size-independent properties - descending top-k
    keys, detections in descending score, idempotence, batch-composition independence."""
    spec, batch, out, rec, res, st = _run(spec_name, [gid], (1.0, 1.0, 1.0, 1.0))
    S = spec.num_levels
    np.testing.assert_allclose(res.image_scores.cpu().numpy(), np.asarray(out["image_scores"], dtype=np.float32),
                               rtol=RTOL, atol=1e-5)
    n = int(res.n_det[0])
    assert n == len(out["dets"][0])
    assert np.array_equal(res.det_flat[0, :n].cpu().numpy(), out["det_flat"][0].numpy())
    d = res.dets[0, :n, 4].cpu().numpy()
    assert np.all(d[:-1] >= d[1:])
    koff = np.concatenate([[0], np.cumsum(spec.level_k)])
    rm = res.row_max[0].cpu().numpy()
    for s in range(S):
        if spec.level_sizes[s] > spec.level_k[s] and spec.head == HEAD_RETINA:
            k = rm[koff[s]:koff[s + 1]]                 # Retina: ranking key == row max, bit for bit
            assert np.all(k[:-1] >= k[1:])
    # free-running: same image scored alone and inside a batch of 3, twice -> bit-identical score
    spec2, batch3 = make_batch(spec_name, [gid + 1, gid, gid + 2])
    sc = Scorer(spec, ScoringParams(n_samples=100), max_batch=3, device="cuda:0")
    r3 = sc.score(batch3["cls_scores"], batch3["bbox_preds"], batch3["L_scores"], batch3["anchors"],
                  batch3["img_shapes"], batch3["scale_factors"], image_ids=batch3["gids"]).image_scores.cpu().numpy().copy()
    r3b = sc.score(batch3["cls_scores"], batch3["bbox_preds"], batch3["L_scores"], batch3["anchors"],
                   batch3["img_shapes"], batch3["scale_factors"], image_ids=batch3["gids"]).image_scores.cpu().numpy().copy()
    r1 = sc.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
                  batch["img_shapes"], batch["scale_factors"], image_ids=batch["gids"]).image_scores.cpu().numpy().copy()
    assert np.array_equal(r3, r3b)
    assert r1[0] == r3[1]


def test_many_pairs_take_the_object_range_rounds():
    """Lowered fg / cluster thresholds give ~26 000 (box, object) pairs for one image - more than the
    K3c shared-memory sort holds (8192), so the object-range rounds are exercised; samples injected,
    T reduced to keep the oracle light.  Also covers non-default thresholds on every stage."""
    params = ScoringParams(n_samples=4, fg_thr=0.1, cluster_iou=0.2)
    spec, batch, out, rec, res, st = _run("tiny_retina_voc", [0], (1.0, 1.0, 1.0, 1.0), params)
    n_pairs = int(res.pair_off[0, spec.num_levels])
    assert n_pairs == sum(len(r["row"]) for r in out["flat"])
    assert n_pairs > 8192, n_pairs
    row, obj, cls, tot, ale, epi, recs = oracle_pairs(out, 0)
    assert np.array_equal(res.pair_row[0, :n_pairs].cpu().numpy(), row)
    assert np.array_equal(res.pair_obj[0, :n_pairs].cpu().numpy(), obj)
    assert_epi_close(res.pair_unc[0, :n_pairs, 2].cpu().numpy(), tot, ale, epi)
    np.testing.assert_allclose(res.image_scores.cpu().numpy(), np.asarray(out["image_scores"], dtype=np.float32),
                               rtol=RTOL, atol=1e-4)


def test_repeated_calls_are_bitwise_reproducible():
    """The same Scorer called again (same images, then other images) gives bit-identical outputs to a
    fresh Scorer: no state leaks between calls through the workspace."""
    spec, b1 = make_batch("cfg1_retina_r50_512_voc", [0, 1])
    _, b2 = make_batch("cfg1_retina_r50_512_voc", [2, 3])
    params = ScoringParams(n_samples=40)
    fresh = lambda: Scorer(spec, params, max_batch=2, device="cuda:0")
    def run(sc, b):
        r = sc.score(b["cls_scores"], b["bbox_preds"], b["L_scores"], b["anchors"], b["img_shapes"],
                     b["scale_factors"], image_ids=b["gids"])
        torch.cuda.synchronize()
        return {k: getattr(r, k).cpu().numpy().copy() for k in ("score_rows", "topk_idx", "boxes", "row_max", "dets",
                                                                  "det_flat", "n_det", "pair_off", "image_scores")}
    ref1, ref2 = run(fresh(), b1), run(fresh(), b2)
    sc = fresh()
    run(sc, b1)
    again1 = run(sc, b1)
    again2 = run(sc, b2)
    for k in ref1:
        assert np.array_equal(again1[k], ref1[k]), k
        assert np.array_equal(again2[k], ref2[k]), k


def test_per_image_shapes_scale_factors_and_pair_overflow():
    """Images of one batch with different img_shape (clip window) and scale_factor; and the loud
    failure when an image produces more pairs than pair_cap."""
    from aod_meh_hua_b200 import _lib
    spec, batch = make_batch("tiny_retina_coco", [0, 1, 2])
    batch["img_shapes"] = [(96, 128, 3), (80, 100, 3), (64, 128, 3)]
    batch["scale_factors"] = [(1.0, 1.0, 1.0, 1.0), (1.25, 1.2, 1.25, 1.2), (0.5, 0.75, 0.5, 0.75)]
    params = ScoringParams()
    out, rec = run_oracle(spec, batch, params)
    sc = Scorer(spec, params, max_batch=3, device="cuda:0")
    B = sc.bind(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
                batch["img_shapes"], batch["scale_factors"], image_ids=batch["gids"])
    sc.k1(); torch.cuda.synchronize()
    override, swapped = check_topk_order(spec, out, sc.result().topk_idx.cpu().numpy())
    if swapped:
        out, rec = run_oracle(spec, batch, params, topk_override=override)
    sc.nms(); sc.pairs()
    inj, off = injection_buffers(spec, rec, B, sc.device)
    sc.k2(inj, off); sc.hua()
    res = sc.result()
    np.testing.assert_allclose(res.boxes.cpu().numpy(), out["boxes"].numpy(), rtol=RTOL, atol=1e-4)
    for b in range(3):
        n = int(res.n_det[b])
        assert np.array_equal(res.det_flat[b, :n].cpu().numpy(), out["det_flat"][b].numpy())
    np.testing.assert_allclose(res.image_scores.cpu().numpy(), np.asarray(out["image_scores"], dtype=np.float32),
                               rtol=RTOL, atol=1e-5)
    # pair_cap too small -> MEHHUA_ST_PAIR_OVERFLOW -> exception, never a silently truncated score
    small = Scorer(spec, params, max_batch=3, device="cuda:0", pair_cap=8)
    with pytest.raises(_lib.MehhuaError):
        small.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
                    batch["img_shapes"], batch["scale_factors"], image_ids=batch["gids"])


def test_pool_topk_randomised_against_numpy():
    """K4 / block radix select on adversarial score distributions: heavy ties, negatives, NaNs,
    denormals, constant arrays, k from 1 to beyond the number of candidates."""
    from aod_meh_hua_b200.scoring import pool_topk
    rs = np.random.RandomState(123)
    cases = []
    for n in (1, 7, 1000, 4097, 70001):
        cases.append(rs.rand(n).astype(np.float32))
        cases.append(np.round(rs.rand(n) * 4).astype(np.float32))                    # 5 distinct values
        cases.append((rs.randn(n) * 1e-3).astype(np.float32))                        # negatives around zero
        cases.append(np.full(n, 0.25, dtype=np.float32))                             # constant
        x = rs.rand(n).astype(np.float32); x[rs.rand(n) < 0.1] = np.nan; cases.append(x)
        cases.append((rs.rand(n) * 1e-40).astype(np.float32))                        # denormals
    for x in cases:
        n = len(x)
        mask = rs.rand(n) < 0.7
        for k in sorted({1, min(n, 17), min(n, 4096), min(n, 5000), n}):
            got = pool_topk(torch.from_numpy(x).cuda(), k, torch.from_numpy(mask).cuda()).cpu().numpy()
            cand = np.nonzero(mask & ~np.isnan(x))[0]
            order = np.argsort(x[cand], kind="stable")
            want = cand[order[-k:]][::-1] if k <= len(cand) else cand[order][::-1]
            assert np.array_equal(got, want), (n, k, x[:5])


def test_pool_topk_million_image_pool():
    """cfg 5's pool: 10^6 scores, k = 2.5 % - the grid-wide form of K4 (radix select with grid-wide histograms,
    chunk sort, merge-path merges).  Exact zeros (object-less images) make the index digits decide; k is
    also taken beyond the non-zero scores and beyond the candidates."""
    from aod_meh_hua_b200.scoring import pool_topk
    rs = np.random.RandomState(77)
    n = 1_000_000
    x = (rs.gamma(2.0, 1.5, n)).astype(np.float32)
    x[rs.rand(n) < 0.4] = 0.0
    mask = rs.rand(n) < 0.9
    xs, ms = torch.from_numpy(x).cuda(), torch.from_numpy(mask).cuda()
    cand = np.nonzero(mask)[0]
    order = np.argsort(x[cand], kind="stable")
    for k in (1, 4096, 25_000, 100_001, 600_000, len(cand), n):
        got = pool_topk(xs, k, ms).cpu().numpy()
        want = cand[order[-k:]][::-1] if k <= len(cand) else cand[order][::-1]
        assert np.array_equal(got, want), k
    got = pool_topk(xs, 25_000).cpu().numpy()                   # no mask
    assert np.array_equal(got, np.argsort(x, kind="stable")[-25_000:][::-1])


def test_scores_do_not_depend_on_the_batch_they_were_computed_in():
    """K2 splits the samples of a pair over four warps when a launch has few pairs (<= 8192) and lets
    one warp walk them otherwise; both orders of work combine the same four partial sums in the same
    order, so an image's score is bit-identical whether it is scored in a batch of 2 or of 24."""
    spec, batch = make_batch("cfg1_retina_r50_512_voc", list(range(24)))
    params = ScoringParams(n_samples=100)
    big = Scorer(spec, params, max_batch=24, device="cuda:0")
    res = big.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
                    batch["img_shapes"], batch["scale_factors"], image_ids=batch["gids"])
    assert int(res.pair_off[:, -1].sum()) > 8192, "the large batch must take the unsplit path"
    want = res.image_scores.cpu().numpy().copy()
    unc_big = [res.pair_unc[b, : int(res.pair_off[b, -1])].cpu().numpy().copy() for b in range(24)]
    small = Scorer(spec, params, max_batch=2, device="cuda:0")
    for lo in range(0, 24, 2):
        sl = slice(lo, lo + 2)
        r = small.score([t[sl] for t in batch["cls_scores"]], [t[sl] for t in batch["bbox_preds"]],
                        [t[sl] for t in batch["L_scores"]], batch["anchors"], batch["img_shapes"][sl],
                        batch["scale_factors"][sl], image_ids=batch["gids"][sl])
        assert int(r.pair_off[:, -1].sum()) <= 8192
        for j in range(2):
            n = int(r.pair_off[j, -1])
            assert np.array_equal(r.pair_unc[j, :n].cpu().numpy(), unc_big[lo + j])
        assert np.array_equal(r.image_scores.cpu().numpy(), want[sl])


def test_scores_do_not_depend_on_the_batch_large_batch_few_pairs():
    """Batches above 64 images always take the per-pair form of K2, even with few pairs; the same
    images in batches of 7 take the split form.  Bit-identical scores either way."""
    spec, batch = make_batch("tiny_retina_coco", list(range(70)))
    params = ScoringParams(n_samples=500)
    big = Scorer(spec, params, max_batch=70, device="cuda:0")
    res = big.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
                    batch["img_shapes"], batch["scale_factors"], image_ids=batch["gids"])
    assert 0 < int(res.pair_off[:, -1].sum()) <= 8192
    want = res.image_scores.cpu().numpy().copy()
    small = Scorer(spec, params, max_batch=7, device="cuda:0")
    for lo in range(0, 70, 7):
        sl = slice(lo, lo + 7)
        r = small.score([t[sl] for t in batch["cls_scores"]], [t[sl] for t in batch["bbox_preds"]],
                        [t[sl] for t in batch["L_scores"]], batch["anchors"], batch["img_shapes"][sl],
                        batch["scale_factors"][sl], image_ids=batch["gids"][sl])
        assert np.array_equal(r.image_scores.cpu().numpy(), want[sl])


def test_entropy_all_without_lambda_matches_the_nol_head():
    """Entropy_ALL route of Lambda_L2Net_NoL / _ReLU (alpha = softmax row, no lambda'): the oracle form
    is pinned against the reference head's own output (tests/golden/all_nol_retina_coco.npz); here the
    kernels against that oracle with its samples injected."""
    from oracle import meh_hua_oracle as O
    from tests.helpers import Recorder
    kind = "scaleSum_classSum"
    spec, batch = make_batch("tiny_retina_coco", [0, 1, 2])
    params = ScoringParams(agg=kind, use_lambda=False)
    rec = Recorder(6)
    out = O.score_batch_all(batch, kind=kind, sampler=rec, **O.spec_kwargs(spec, params))
    sc = Scorer(spec, params, max_batch=3, device="cuda:0", mode="all")
    B = sc.bind(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
                batch["img_shapes"], batch["scale_factors"], image_ids=batch["gids"])
    sc.all_rows()
    inj, off = injection_buffers(spec, rec, B, sc.device)
    sc.k2(inj, off)
    sc.hua()
    torch.cuda.synchronize()
    res = sc.result()
    poff = res.pair_off.cpu().numpy()
    for b in range(B):
        for r in [r for r in out["flat"] if r["image"] == b]:
            a, e = poff[b, r["level"]], poff[b, r["level"] + 1]
            assert np.array_equal(res.topk_idx[b, a:e].cpu().numpy(), r["prior"])
            np.testing.assert_allclose(res.score_rows[b, a:e].cpu().numpy(), r["alpha"], rtol=2e-5, atol=1e-12)
            unc = res.pair_unc[b, a:e].cpu().numpy()
            np.testing.assert_allclose(unc[:, 1], r["ale"], rtol=RTOL, atol=2e-6)
            assert_epi_close(unc[:, 2], r["total"], r["ale"], r["epi"])
    np.testing.assert_allclose(res.image_scores.cpu().numpy(), np.asarray(out["image_scores"], dtype=np.float64),
                               rtol=RTOL, atol=1e-5)
