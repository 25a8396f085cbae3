"""K2's free-running sampler against the Dirichlet closed forms (`mean_t x_c -> alpha_c/alpha_0`,
`E[ale] -> psi(alpha_0+1) - sum (alpha_c/alpha_0) psi(alpha_c+1)`), per class and per alpha regime,
including the tiny-alpha regime of real softmax rows (alpha ~ 1e-3 ... 1e-6), the two-pass
form for rows with alpha_0 < 1 and the analytic (T -> infinity) form.

Tolerances are statistical and self-calibrated: every pair is an independent replicate, so the standard
error of a mean over pairs is std / sqrt(P); the assertions allow 4.5 standard errors plus the
enumerated grid bias of the GS sampler (profiles/r2_gs_grid_bias.txt, <= 2e-5 for alpha >= 1e-3).
"""
import numpy as np
import pytest
import torch

from aod_meh_hua_b200.scoring import pair_uncertainty
from aod_meh_hua_b200.specs import ScoringParams
from oracle import meh_hua_oracle as O

pytestmark = pytest.mark.gpu

N_MINOR = 63


def _replicates(row, P, T, seed_ids=(1, 0), lib=None):
    """P independent pairs with the same alpha row -> (unc [P,3], avg [P,C]) as float64 numpy."""
    alpha = torch.tensor(row, dtype=torch.float32, device="cuda:0").repeat(P, 1).contiguous()
    params = ScoringParams(n_samples=T, use_lambda=False)
    unc, avg = pair_uncertainty(alpha, torch.ones(P, device="cuda:0"), torch.arange(P), torch.zeros(P, dtype=torch.long),
                                params, seed_ids=seed_ids, return_avg=True, lib=lib)
    return unc.cpu().numpy().astype(np.float64), avg.cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("big", [1.0, 30.0])
@pytest.mark.parametrize("minor", [0.3, 1e-2, 3e-3, 1e-3, 1e-4, 2.5e-5, 1e-5, 1e-6])
def test_per_class_means_single_dominant_row(minor, big):
    """One dominant class (alpha = 1 or 30: Marsaglia-Tsang, fixes the reference exponent) and 63 minor
    classes with the same alpha: mean_t x_c over T * P samples against alpha_c / alpha_0 per class group,
    and the aleatoric mean against its closed form.  VERDICT r1 weak #1: the 16-bit grid of round 1 read
    0.96 at 1e-4, 0.78 at 1e-5 and 0.011 at 1e-6 here."""
    row = [big] + [minor] * N_MINOR
    P, T = 2048, 100_000
    unc, avg = _replicates(row, P, T)
    a = np.asarray(row, dtype=np.float64)
    a0 = a.sum()
    # minor classes: one estimate per pair = mean over the 63 copies (they are exchangeable)
    est = avg[:, 1:].mean(axis=1)
    se = est.std(ddof=1) / np.sqrt(P)
    want = minor / a0
    assert abs(est.mean() - want) <= 4.5 * se + 2e-5 * want, (minor, big, est.mean() / want, se / want)
    # dominant class
    estb = avg[:, 0]
    seb = max(estb.std(ddof=1) / np.sqrt(P), 1e-9)
    assert abs(estb.mean() - big / a0) <= 4.5 * seb + 1e-6, (estb.mean(), big / a0, seb)
    # every sample is normalised: the class means of a pair sum to 1
    np.testing.assert_allclose(avg.sum(axis=1), 1.0, rtol=0, atol=2e-3)
    # aleatoric = E[-sum x ln x]
    h, e_ent, _ = O.dirichlet_expectations(a[None, :])
    se_a = unc[:, 1].std(ddof=1) / np.sqrt(P)
    assert abs(unc[:, 1].mean() - e_ent[0]) <= 4.5 * se_a + 2e-5 * max(e_ent[0], 1e-3) + 1e-6, (unc[:, 1].mean(), e_ent[0], se_a)
    # total = H(mean_t x): per pair it is the entropy of the pair's own class means (a concave function of
    # a noisy mean, so E[total] < H(alpha/alpha_0) by a Jensen gap that is large for tiny alpha - the
    # reference has the same gap); pooled over all pairs the class means give H(alpha/alpha_0)
    avg_c = np.maximum(avg, 1.17549435e-38)
    np.testing.assert_allclose(unc[:, 0], -(avg_c * np.log(avg_c)).sum(axis=1), rtol=2e-5, atol=2e-7)
    pooled = np.maximum(avg.mean(axis=0), 1e-300)
    h_pooled = -(pooled * np.log(pooled)).sum()
    assert abs(h_pooled - h[0]) <= (4.5 * se / want + 1e-4) * h[0] + 1e-7, (h_pooled, h[0], se / want)


def test_bias_table_written_for_profiles(tmp_path):
    """The table VERDICT r1 asked for (mean_t x_c / (alpha_c/alpha_0) per alpha with its Monte-Carlo
    standard error), printed so that a GPU run can be copied into profiles/."""
    lines = ["# alpha_minor  ratio = mean_t x_c / (alpha_c/alpha_0)   standard error   (dominant alpha = 1, 63 minor classes, T*P = 2.0e8 samples)"]
    for minor in [1e-2, 1e-3, 1e-4, 2.5e-5, 1e-5, 1e-6]:
        row = [1.0] + [minor] * N_MINOR
        P, T = 2048, 100_000
        _, avg = _replicates(row, P, T, seed_ids=(7, 3))
        est = avg[:, 1:].mean(axis=1)
        want = minor / (1.0 + N_MINOR * minor)
        ratio, se = est.mean() / want, est.std(ddof=1) / np.sqrt(P) / want
        lines.append(f"{minor:10.2e}   {ratio:10.5f}   {se:10.5f}")
        assert abs(ratio - 1.0) <= 4.5 * se + 2e-5
    print("\n" + "\n".join(lines))
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "k2_bias_table.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")


@pytest.mark.parametrize("row", [
    [0.02, 0.01, 0.005, 0.001, 0.0003, 0.3, 0.1, 0.05],            # alpha_0 = 0.486: two-pass form
    [0.004, 0.002, 0.001, 0.0005, 0.003, 0.0001, 0.00005, 0.0002],  # alpha_0 = 0.011: almost one-hot samples
    [0.9, 0.05, 0.02],                                              # alpha_0 = 0.97
])
def test_rows_with_alpha0_below_one_take_the_two_pass_form(row):
    """with every alpha < 1 and a small alpha_0 all draws of a sample can fall below
    2^-126; round 1 dropped such samples (biasing total / aleatoric low).  The two-pass form
    normalises each sample around its own largest log2 draw: class means and the mean entropy match
    the closed forms, and no sample is lost (class means sum to 1)."""
    P, T = 2048, 40_000
    unc, avg = _replicates(row, P, T, seed_ids=(2, 5))
    a = np.asarray(row, dtype=np.float64)
    a0 = a.sum()
    np.testing.assert_allclose(avg.sum(axis=1), 1.0, rtol=0, atol=2e-3)
    for c in range(len(row)):
        est = avg[:, c]
        se = est.std(ddof=1) / np.sqrt(P)
        assert abs(est.mean() - a[c] / a0) <= 4.5 * se + 2e-5 * a[c] / a0 + 1e-9, (c, est.mean(), a[c] / a0, se)
    h, e_ent, _ = O.dirichlet_expectations(a[None, :])
    se_a = unc[:, 1].std(ddof=1) / np.sqrt(P)
    assert abs(unc[:, 1].mean() - e_ent[0]) <= 4.5 * se_a + 1e-6, (unc[:, 1].mean(), e_ent[0], se_a)


def test_softmax_like_rows_with_tiny_alphas():
    """A peaked 80-class softmax row times lambda' (the regime SURVEY 7 describes for real heads: most
    classes at alpha ~ 1e-3 ... 1e-6): per-class means and aleatoric mean vs closed forms."""
    rs = np.random.RandomState(3)
    logits = rs.randn(80) * 2.0 - 8.0
    logits[17] = 3.0
    logits[41] = 0.5
    p = np.exp(logits - logits.max())
    p /= p.sum()
    row = (p * 18.0).astype(np.float32)
    assert (row < 1e-3).sum() > 40 and row.min() < 1e-5
    P, T = 1024, 50_000
    unc, avg = _replicates(row.tolist(), P, T, seed_ids=(9, 2))
    a = row.astype(np.float64)
    a0 = a.sum()
    est, se = avg.mean(axis=0), avg.std(axis=0, ddof=1) / np.sqrt(P)
    bad = np.abs(est - a / a0) > 4.5 * se + 2e-5 * a / a0 + 1e-12
    assert bad.sum() <= 1, (np.nonzero(bad)[0], est[bad], (a / a0)[bad], se[bad])     # 80 tests at 4.5 sigma
    # summed over the tiny classes (the mass round 1 lost): within 4.5 standard errors
    tiny = a < 1e-3
    est_t = avg[:, tiny].sum(axis=1)
    se_t = est_t.std(ddof=1) / np.sqrt(P)
    assert abs(est_t.mean() - a[tiny].sum() / a0) <= 4.5 * se_t
    h, e_ent, _ = O.dirichlet_expectations(a[None, :])
    se_a = unc[:, 1].std(ddof=1) / np.sqrt(P)
    assert abs(unc[:, 1].mean() - e_ent[0]) <= 4.5 * se_a + 1e-6


def test_analytic_form_matches_closed_forms():
    """n_samples = 0: total = H(alpha/alpha_0), aleatoric = psi(alpha_0+1) - sum m_c psi(alpha_c+1),
    class means = alpha/alpha_0 (double precision on the device)."""
    rs = np.random.RandomState(11)
    rows = np.abs(rs.randn(64, 20)).astype(np.float32) * np.float32(0.5) + np.float32(1e-4)
    rows[:, 3] += rs.rand(64).astype(np.float32) * 20
    rows[5] *= np.float32(1e-3)
    alpha = torch.from_numpy(rows).cuda()
    P = rows.shape[0]
    params = ScoringParams(n_samples=0, use_lambda=False)
    unc, avg = pair_uncertainty(alpha, torch.ones(P, device="cuda:0"), torch.arange(P), torch.zeros(P, dtype=torch.long),
                                params, return_avg=True)
    h, e_ent, epi = O.dirichlet_expectations(rows.astype(np.float64))
    np.testing.assert_allclose(unc[:, 0].cpu().numpy(), h, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(unc[:, 1].cpu().numpy(), e_ent, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(unc[:, 2].cpu().numpy(), epi, rtol=1e-5, atol=2e-7)
    np.testing.assert_allclose(avg.cpu().numpy(), rows / rows.sum(axis=1, keepdims=True), rtol=2e-6, atol=1e-12)
