"""Gaussian collapse (``thy_model_collapse``) and the device-built analytic Gaussian posterior
(``thy_model_gaussian_posterior``; GaussianAnalyticPosterior.scala:141-169) against the oracle.

The collapsed posterior must be the SAME function as the canonical one: log-posterior and gradient within 1e-10
of the oracle's canonical evaluation ("must show the 1e-10 parity still holds near the mode"),
at typical-set points (where cancellation is worst), at the mode itself and in the far field.
"""
import math

import numpy as np
import pytest

import thylacine_b200 as t
from thylacine_b200 import capi
from oracle import ref_oracle as o
from helpers import linear_gaussian_problem, typical_set_points, rel_err_logp, rel_err_grad

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _posterior(prob, **kw):
    return t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                   [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")], **kw)


@pytest.mark.parametrize("d,n,B", [(12, 200, 96), (100, 10_000, 512), (37, 37, 80), (200, 3000, 130)])
def test_collapsed_posterior_is_the_same_function(d, n, B):
    prob = linear_gaussian_problem(d, n, seed=100 + d)
    post = _posterior(prob)
    col = post.collapsed()
    assert col.domainDimension == d and col.layout == post.layout
    mean, _ = o.analytic_gaussian_posterior(prob["pm"], prob["pvar"], prob["A"], prob["y"], prob["var"])
    rng = np.random.default_rng(d)
    X = np.concatenate([typical_set_points(prob, B, seed=1), mean[None, :],            # typical set + the mode
                        rng.standard_normal((B, d)) * 10.0])                            # prior draws (far field)
    want_lp, want_g = o.linear_gaussian_batch(prob["A"], prob["y"], prob["var"], X, prior_mean=prob["pm"],
                                              prior_var=prob["pvar"])
    lp, g = col.logPdfGradBatch(X)
    assert rel_err_logp(lp, want_lp) < TOL
    # gradient tolerance relative to ||grad||_inf of the oracle per point; at the mode the canonical gradient itself
    # is pure rounding noise (A^T W r^ = 0 analytically), so the mode row is compared in absolute terms
    rows = np.arange(X.shape[0]) != B
    assert rel_err_grad(g[rows], want_g[rows]) < TOL
    scale = np.max(np.abs(prob["A"].T @ (np.abs(prob["y"]) / prob["var"])))
    assert np.max(np.abs(g[B] - want_g[B])) < TOL * scale
    # and against the canonical device path
    lp2, g2 = post.logPdfGradBatch(X)
    assert rel_err_logp(lp, lp2) < TOL and rel_err_grad(g[rows], g2[rows]) < TOL
    assert col.lastKernelPath in ("fused_dmma", "resident_dmma", "stream", "stream_dmma", "wide_dmma", "generic")
    col.close()
    post.close()


@pytest.mark.parametrize("d,B", [(100, 24_003), (34, 40_000)])
def test_resident_kernel_large_batch_with_staged_rows_matches_small_batches_and_the_oracle(d, B):
    """The resident kernel changes shape with the batch: up to ~4 700 chains a group of 8 chains is the work of a PAIR of
    warps, from ~19 000 chains on the chain rows of the next group are staged in shared memory by cp.async while the
    current one computes (kernels_fused_dmma.cu, k_resident_dmma).  One large device-resident batch (staged rows, ragged
    last group) against the same points in chunks of 2 000 (warp pairs, no staging): same function to summation order;
    and against the oracle's canonical evaluation at 1e-10."""
    import torch
    n = 1500
    prob = linear_gaussian_problem(d, n, seed=300 + d)
    col = _posterior(prob).collapsed()
    X = typical_set_points(prob, B, seed=2)
    Xd = torch.from_numpy(X).cuda()
    lp = torch.empty(B, dtype=torch.float64, device="cuda")
    g = torch.empty(B, d, dtype=torch.float64, device="cuda")
    col.logPdfGradBatchDevice(B, Xd.data_ptr(), lp.data_ptr(), g.data_ptr())
    col.synchronize()
    assert col.lastKernelPath == "resident_dmma"
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    lp_small = np.empty(B)
    g_small = np.empty((B, d))
    for b0 in range(0, B, 2000):
        lp_small[b0:b0 + 2000], g_small[b0:b0 + 2000] = col.logPdfGradBatch(X[b0:b0 + 2000])
    assert rel_err_logp(lp, lp_small) < 1e-13 and rel_err_grad(g, g_small) < 1e-12
    pick = np.r_[0:64, B - 64:B]
    want_lp, want_g = o.linear_gaussian_batch(prob["A"], prob["y"], prob["var"], X[pick], prior_mean=prob["pm"],
                                              prior_var=prob["pvar"])
    assert rel_err_logp(lp[pick], want_lp) < TOL and rel_err_grad(g[pick], want_g) < TOL
    lv = torch.empty(B, dtype=torch.float64, device="cuda")                        # value only: the general output step
    col.logPdfGradBatchDevice(B, Xd.data_ptr(), lv.data_ptr(), None)
    col.synchronize()
    assert np.array_equal(lv.cpu().numpy(), lp)


def test_collapse_multi_label_offset_two_likelihoods_and_uniform_prior():
    rng = np.random.default_rng(21)
    n1, n2 = 70, 45
    A1 = {"gamma": rng.standard_normal((n1, 2)), "alpha": rng.standard_normal((n1, 3))}      # skips beta
    A2 = {"beta": rng.standard_normal((n2, 5))}
    off = rng.standard_normal(n1)
    y1, y2 = rng.standard_normal(n1) + 50.0, rng.standard_normal(n2)                         # large |b| vs |G|
    ci1, ci2 = rng.uniform(0.1, 2, n1), rng.uniform(0.1, 2, n2)
    cov = np.array([[2.0, 0.3], [0.3, 1.0]])
    priors = [t.GaussianPrior.fromConfidenceIntervals("beta", rng.standard_normal(5), rng.uniform(1, 3, 5)),
              t.UniformPrior.fromBounds("alpha", [40, 40, 40], [-40, -40, -40]),
              t.GaussianPrior.fromCovarianceMatrix("gamma", [0.5, -0.5], cov),
              t.GaussianPrior.fromConfidenceIntervals("unused", [1.0, 2.0], [3.0, 4.0])]     # no likelihood reads it
    liks = [t.GaussianLikelihood(t.LinearForwardModel.of(transform=A1, vectorOffset=off), y1, ci1),
            t.GaussianLinearLikelihood.of(A2["beta"], y2, ci2, "beta")]
    post = t.UnnormalisedPosterior(priors, liks)
    col = post.collapsed()
    ref = o.Posterior([o.gaussian_prior_from_ci("beta", priors[0].a, priors[0].b),
                       o.uniform_prior_from_bounds("alpha", [40, 40, 40], [-40, -40, -40]),
                       o.gaussian_prior_from_cov("gamma", [0.5, -0.5], cov),
                       o.gaussian_prior_from_ci("unused", [1.0, 2.0], [3.0, 4.0])],
                      [o.gaussian_likelihood(o.LinearForwardModel(A1, off), y1, ci1),
                       o.gaussian_linear_likelihood(A2["beta"], y2, ci2, "beta")])
    X = rng.uniform(-3.9, 3.9, (41, 12))
    X[5, 1] = 45.0                                                             # outside the uniform box
    lp, g = col.logPdfGradBatch(X)
    for b in range(X.shape[0]):
        want = ref.log_pdf(X[b])
        if math.isinf(want):
            assert lp[b] == -math.inf
        else:
            assert lp[b] == pytest.approx(want, rel=TOL)
        assert rel_err_grad(g[b][None], ref.log_pdf_gradient(X[b])[None]) < TOL
    # a collapsed posterior collapses again onto itself
    col2 = col.collapsed()
    lp2, g2 = col2.logPdfGradBatch(X)
    ok = np.isfinite(lp)
    assert rel_err_logp(lp2[ok], lp[ok]) < TOL and rel_err_grad(g2, g) < TOL
    for p in (col2, col, post):
        p.close()


def test_collapse_rejects_what_is_not_quadratic():
    rng = np.random.default_rng(3)
    A, y, ci = rng.standard_normal((30, 4)), rng.standard_normal(30), np.full(30, 0.5)
    pri = [t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(4), np.ones(4))]
    cauchy = t.UnnormalisedPosterior(pri, [t.CauchyLikelihood.of(A, y, ci, "theta")])
    with pytest.raises(t.ThylacineError) as e:
        cauchy.collapsed()
    assert e.value.status == capi.THY_ERR_UNSUPPORTED
    fm = t.NonLinearForwardModel.of("tanh", {"theta": A}, differential=None, domainDimensions={"theta": 4}, rangeDimension=30)
    nonlinear = t.UnnormalisedPosterior(pri, [t.GaussianLikelihood(fm, y, ci)])
    with pytest.raises(t.ThylacineError) as e:
        nonlinear.collapsed()
    assert e.value.status == capi.THY_ERR_UNSUPPORTED
    wide = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(40), np.ones(40))],
                                   [t.GaussianLinearLikelihood.of(rng.standard_normal((30, 40)), y, ci, "theta")])
    with pytest.raises(t.ThylacineError) as e:                                 # n < d: the Gram matrix is singular
        wide.collapsed()
    assert e.value.status == capi.THY_ERR_UNSUPPORTED
    priors_only = t.UnnormalisedPosterior(pri)
    with pytest.raises(t.ThylacineError) as e:
        priors_only.collapsed()
    assert e.value.status == capi.THY_ERR_UNSUPPORTED
    for p in (cauchy, nonlinear, wide, priors_only):
        p.close()


@pytest.mark.parametrize("d,n", [(10, 1000), (100, 10_000), (300, 2000)])
def test_device_gaussian_posterior_matches_the_analytic_oracle(d, n):
    prob = linear_gaussian_problem(d, n, seed=7 + d)
    post = _posterior(prob)
    mean, cov, prec, lnz, mx = post.gaussianPosterior()
    want_mean, want_cov = o.analytic_gaussian_posterior(prob["pm"], prob["pvar"], prob["A"], prob["y"], prob["var"])
    sd = np.sqrt(np.diag(want_cov))
    assert np.max(np.abs(mean - want_mean) / sd) < 1e-8                         # in units of the posterior width
    assert np.max(np.abs(cov - want_cov)) / np.max(np.abs(want_cov)) < 1e-9
    want_prec = np.diag(1.0 / prob["pvar"]) + prob["A"].T @ (prob["A"] / prob["var"][:, None])
    assert np.max(np.abs(prec - want_prec)) / np.max(np.abs(want_prec)) < 1e-12
    want_lnz = o.analytic_gaussian_log_evidence(prob["pm"], prob["pvar"], prob["A"], prob["y"], prob["var"])
    assert lnz == pytest.approx(want_lnz, rel=1e-10)
    sign, logdet = np.linalg.slogdet(want_prec)
    assert mx == pytest.approx(0.5 * logdet - 0.5 * d * math.log(2 * math.pi), rel=1e-10, abs=1e-9)
    post.close()


def test_hmc_on_the_collapsed_posterior_samples_the_same_distribution():
    """Same engine, same seeds: the collapsed posterior has the same value and gradient to rounding, so early
    trajectories coincide and the sample moments match the analytic posterior."""
    d, n, B = 20, 2000, 2048
    prob = linear_gaussian_problem(d, n, seed=5)
    post = _posterior(prob)
    col = post.collapsed()
    mean, cov = o.analytic_gaussian_posterior(prob["pm"], prob["pvar"], prob["A"], prob["y"], prob["var"])
    eps = 0.25 * float(np.sqrt(np.min(np.linalg.eigvalsh(cov))))
    cfg = t.HmcmcConfig(stepsBetweenSamples=1, stepsInDynamicsSimulation=12, warmupStepCount=60, dynamicsSimulationStepSize=eps)
    seed = {"theta": mean}
    got = []
    for p in (post, col):
        hmc = t.HmcmcSampledPosterior(cfg, p, numberOfChains=B, rngSeed=9, seed=seed)
        s = hmc.sampleChains(4).reshape(-1, d)
        got.append(s)
        hmc.close()
    ess = B                                                                     # chains are independent
    for s in got:
        assert np.max(np.abs(s.mean(0) - mean) / np.sqrt(np.diag(cov))) < 5.0 / math.sqrt(ess)
        assert np.max(np.abs(s.var(0) / np.diag(cov) - 1.0)) < 5.0 * math.sqrt(2.0 / ess)
    col.close()
    post.close()


def test_full_size_config3_canonical_against_collapsed_and_oracle():
    """BASELINE config 3's model at full size (1000-d, 100k observations; A = 800 MB): two device algorithms that
    share no kernel configuration -- the canonical wide DMMA pass over all 10^5 observations and the collapsed
    1000-observation term -- must be the same function, and both match the oracle on a slice."""
    d, n, B = 1000, 100_000, 512
    prob = linear_gaussian_problem(d, n, seed=20260103)
    post = _posterior(prob)
    col = post.collapsed()
    mean, _, _, _, _ = post.gaussianPosterior()
    rng = np.random.default_rng(3)
    X = np.concatenate([mean[None, :] + 0.01 * rng.standard_normal((B // 2, d)), rng.standard_normal((B // 2, d))])
    lp, g = post.logPdfGradBatch(X)
    assert post.lastKernelPath == "wide_dmma"
    lpc, gc = col.logPdfGradBatch(X)
    assert rel_err_logp(lpc, lp) < TOL and rel_err_grad(gc, g) < TOL
    want_lp, want_g = o.linear_gaussian_batch(prob["A"], prob["y"], prob["var"], X[::64], prior_mean=prob["pm"],
                                              prior_var=prob["pvar"])
    assert rel_err_logp(lp[::64], want_lp) < TOL and rel_err_grad(g[::64], want_g) < TOL
    g0 = post.logPdfGradBatch(mean[None, :])[1]
    assert np.max(np.abs(g0)) < 1e-9 * np.max(np.abs(g))                       # stationary point of the device-built mean
    col.close()
    post.close()
