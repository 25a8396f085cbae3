"""GPU parity of the mutual-information baselines (kernel KM, through the C ABI) against the oracle and
against the outputs of the reference's own ComputeMI / ComputeMCDropoutMI (tests/golden/mi_baselines.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import meh_hua_oracle as O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
# epi = total - ale is a difference of two entropies of order nCls * 0.2: fp32 rounding of either (1e-7 relative)
# is 1e-5 of an epistemic value a hundred times smaller.  The level values are means over hundreds of priors, where
# those errors average out; the bound below is north_star's 1e-5 on the score plus that cancellation floor.
RTOL, ATOL = 1e-5, 2e-7


def _cuda(members):
    return [[t.cuda() for t in m] for m in members]


@pytest.mark.parametrize("key,seed,members", [("ensemble", 301, 3), ("mcdropout", 302, 7)])
def test_mi_matches_the_reference_outputs(key, seed, members):
    from aod_meh_hua_b200.mi_baselines import ComputeMCDropoutMI, ComputeMI, mutual_information
    g = np.load(os.path.join(GOLD, "mi_baselines.npz"))
    x = O.mi_inputs(seed, members)
    want, want_levels = O.compute_mi(x, 20)
    got, got_levels = mutual_information(_cuda(x), 20)
    np.testing.assert_allclose(got_levels.cpu().numpy(), want_levels.numpy(), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(got.cpu().numpy(), g[key], rtol=RTOL, atol=ATOL)
    fn = ComputeMI if key == "ensemble" else ComputeMCDropoutMI
    out = fn(*_cuda(x), nCls=20)
    assert isinstance(out, list) and len(out) == 2 and all(isinstance(v, float) for v in out)
    np.testing.assert_allclose(out, g[key], rtol=RTOL, atol=ATOL)


def test_mi_detector_shapes_and_edge_cases():
    """RetinaNet VOC head shapes (A = 9, 20 classes, a 512 x 512 image's five levels), a plane that does not fill
    its last tile, identical members (MI = 0 up to rounding), and a saturated logit (NaN as in the reference)."""
    from aod_meh_hua_b200.mi_baselines import mutual_information
    rs = np.random.RandomState(9)
    shapes = [(64, 64), (32, 32), (16, 16), (8, 8), (4, 4)]
    x = [[torch.from_numpy((rs.standard_normal((2, 180, h, w)) * 1.5 - 3.0).astype(np.float32)) for (h, w) in shapes]
         for _ in range(3)]
    want, want_levels = O.compute_mi(x, 20)
    got, got_levels = mutual_information(_cuda(x), 20)
    np.testing.assert_allclose(got_levels.cpu().numpy(), want_levels.numpy(), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=RTOL, atol=ATOL)
    same = [x[0], x[0], x[0]]
    got, _ = mutual_information(_cuda(same), 20)
    assert float(got.abs().max()) < 1e-6
    sat = [[t.clone() for t in m] for m in x]
    sat[1][2][0, 5, 3, 3] = -200.0            # sigmoid underflows to 0: 0 * log 0 = NaN in the reference
    want, _ = O.compute_mi(sat, 20)
    got, _ = mutual_information(_cuda(sat), 20)
    assert np.isnan(want.numpy()[0]) and np.isnan(got.cpu().numpy()[0])
    np.testing.assert_allclose(got.cpu().numpy()[1], want.numpy()[1], rtol=RTOL, atol=ATOL)
    with pytest.raises(Exception):
        mutual_information([[t for t in m] for m in x], 20)      # CPU tensors: no fallback


def test_mi_pool_loops_mirror_the_reference_drivers():
    """Ensemble_MI / MCDropout_MI (apis/CalEnsembleUnc.py:137-164, CalMCDropoutUnc.py:137-165) with stub detectors
    that return the `justOut=True` classification maps."""
    from aod_meh_hua_b200.mi_baselines import Ensemble_MI, MCDropout_MI

    class Stub(torch.nn.Module):
        def __init__(self, seed):
            super().__init__()
            self.seed, self.calls, self.drop = seed, 0, torch.nn.Dropout2d(0.5)

        def forward(self, img=None, img_metas=None, **kw):
            assert kw["justOut"] and kw["isEval"] and not kw["return_loss"]
            self.calls += 1
            g = torch.Generator().manual_seed(self.seed + 13 * kw["batchIdx"] + (self.calls if self.drop.training else 0))
            return [torch.randn(len(img_metas[0]), 60, h, w, generator=g).cuda() for (h, w) in ((6, 7), (3, 4))]

    loader = [dict(img=[torch.zeros(2, 3, 8, 8)], img_metas=[[{}, {}]]), dict(img=[torch.zeros(1, 3, 8, 8)], img_metas=[[{}]])]
    ms = [Stub(1), Stub(2), Stub(3)]
    out = Ensemble_MI(*ms, loader)
    assert out.shape == (3,) and out.dtype == torch.float32
    want = []
    for i, d in enumerate(loader):
        outs = []
        for m in ms:
            g = torch.Generator().manual_seed(m.seed + 13 * i)
            outs.append([torch.randn(len(d["img_metas"][0]), 60, h, w, generator=g) for (h, w) in ((6, 7), (3, 4))])
        want.extend(O.compute_mi(outs, 20)[0].tolist())
    np.testing.assert_allclose(out.numpy(), want, rtol=RTOL, atol=ATOL)
    mc = MCDropout_MI(Stub(5), loader, n=4)
    assert mc.shape == (3,) and bool((mc > 0).all())
