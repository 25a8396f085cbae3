"""Pool-level identity (north_star: "an identical selected set to the reference"; SURVEY 8.1 last row).

A 2048-image synthetic pool goes through the CUDA path (K1 -> K3a -> K3b -> K2 -> K3c, then K4 inside the
update_X_L drop-in) and through the CPU oracle.  The Monte-Carlo step cannot be compared draw for draw
(torch's Philox stream is not reproducible), so BOTH sides replace it by its T -> infinity
limit - `n_samples = 0` in the library, `analytic=True` in the oracle: total = H(alpha/alpha_0),
aleatoric = psi(alpha_0+1) - sum_c (alpha_c/alpha_0) psi(alpha_c+1).  Everything else - every
threshold decision, the top-k, NMS, the pair list, the group means, the aggregation, the selection
with its host-RNG draws - is the real path.

Integer decisions hang on fp32 values that the fused kernel and ATen round differently in the last
ulp, so an image may legitimately differ when one of its decisions sits within ~1e-6 (relative) of
its threshold.  The test therefore (a) demands rel. 2e-5 agreement of every image score, except for
images the oracle's own margin report flags as marginal, (b) reports the number of such images (the
"residual flip count" of SURVEY 7), and (c) demands the identical X_L_next / X_U_next arrays.
"""
import numpy as np
import pytest
import torch

from aod_meh_hua_b200 import pool as P
from aod_meh_hua_b200.scoring import Scorer
from aod_meh_hua_b200.specs import ScoringParams, get_spec
from aod_meh_hua_b200.synth import SyntheticPool
from oracle import meh_hua_oracle as O

pytestmark = pytest.mark.gpu

N_POOL = 2048
BATCH = 64
MARGINAL = 5e-6       # a decision closer than this (relative) to its threshold may flip between fp32 implementations


@pytest.fixture(scope="module")
def pool_scores():
    spec = get_spec("tiny_retina_coco")
    params = ScoringParams(n_samples=0)
    synth = SyntheticPool(spec, seed0=20, device="cpu")
    sc = Scorer(spec, params, max_batch=BATCH, device="cuda:0")
    kw = O.spec_kwargs(spec, params)
    kw.pop("T")
    got = np.zeros(N_POOL, dtype=np.float32)
    want = np.zeros(N_POOL, dtype=np.float32)
    batches = {}
    for i0 in range(0, N_POOL, BATCH):
        gids = list(range(i0, i0 + BATCH))
        batch = synth.batch(gids)
        res = sc.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
                       batch["img_shapes"], batch["scale_factors"], image_ids=gids)
        got[i0:i0 + BATCH] = res.image_scores.cpu().numpy()
        out = O.score_batch(batch, analytic=True, **kw)
        want[i0:i0 + BATCH] = np.asarray(out["image_scores"], dtype=np.float32)
        batches[i0] = (batch, out)
    return spec, params, got, want, batches


def test_every_image_score_matches_the_oracle(pool_scores):
    spec, params, got, want, batches = pool_scores
    assert (want > 0).sum() > N_POOL // 2
    bad = np.nonzero(~np.isclose(got, want, rtol=2e-5, atol=2e-6))[0]
    flips = []
    for g in bad:                     # a differing image must be a marginal one
        i0 = (g // BATCH) * BATCH
        batch, out = batches[i0]
        j = g - i0
        one = {k: ([t[j:j + 1] for t in v] if k in ("cls_scores", "bbox_preds", "L_scores") else v)
               for k, v in batch.items()}
        one["img_shapes"], one["scale_factors"], one["gids"] = [batch["img_shapes"][j]], [batch["scale_factors"][j]], [g]
        kw = O.spec_kwargs(spec, params)
        kw.pop("T")
        out1 = O.score_batch(one, analytic=True, **kw)
        m = O.decision_margins(one, out1, **O.spec_kwargs(spec, params))
        worst = min(v for k, v in m.items() if k not in ("topk_adjacent", "nms_adjacent"))
        assert worst < MARGINAL, (g, got[g], want[g], m)
        flips.append((int(g), float(got[g]), float(want[g]), worst))
    print(f"\npool of {N_POOL}: {len(flips)} image(s) differ through a marginal decision (< {MARGINAL:g} rel.): {flips}")
    assert len(flips) <= 2
    # images without an object score exactly 0 on both sides (zeros drive zeroRate in the selection)
    ok = np.ones(N_POOL, dtype=bool)
    ok[bad] = False
    assert np.array_equal(got[ok] == 0, want[ok] == 0)


@pytest.mark.parametrize("kwargs", [{}, {"zeroRate": 0.15, "useMaxConf": "False"}])
def test_selected_set_is_identical(pool_scores, kwargs):
    """update_X_L (utils/active_datasets.py:102-135): the drop-in on the GPU scores against the
    restatement on the oracle scores, same numpy RNG state: identical X_L_next and X_U_next."""
    spec, params, got, want, _ = pool_scores
    rs = np.random.RandomState(3)
    X_all = np.arange(N_POOL)
    X_L = np.sort(rs.choice(N_POOL, 200, replace=False))
    unl = np.setdiff1d(X_all, X_L)
    n_sel = 100
    srt = np.sort(want[unl])[::-1]
    n_top = n_sel - int(n_sel * kwargs.get("zeroRate", 0))
    # the test needs a selection boundary that is not itself a near-tie
    assert (srt[n_top - 1] - srt[n_top]) > 1e-4 * srt[n_top - 1]
    np.random.seed(11)
    want_L, want_U = O.update_X_L(want.copy(), X_all, X_L, n_sel, **kwargs)
    np.random.seed(11)
    got_L, got_U = P.update_X_L(torch.from_numpy(got).cuda(), X_all, X_L, n_sel, **kwargs)
    assert np.array_equal(got_L, want_L)
    assert np.array_equal(got_U, want_U)
    # and straight from the score vectors: the top-n_top id sets agree
    top_got = set(unl[np.argsort(got[unl], kind="stable")[-n_top:]].tolist())
    top_want = set(unl[np.argsort(want[unl], kind="stable")[-n_top:]].tolist())
    assert top_got == top_want
