"""BASELINE.json's configs 1, 4 and 5 at their full sizes, checked through size-independent properties plus oracle
parity on a slice (configs 2 and 3 at full size: test_parity_gpu.py::test_full_size_config2_properties,
test_collapse_gpu.py::test_full_size_config3_canonical_against_collapsed_and_oracle)."""
import math

import numpy as np
import pytest

import thylacine_b200 as t
from oracle import ref_oracle as o
from helpers import linear_gaussian_problem, rel_err_grad

pytestmark = pytest.mark.gpu
TOL = 1e-10


def test_full_size_config5_nonlinear_cauchy_fd_16384_chains():
    """Config 5: tanh(A theta) with a Cauchy likelihood, d = 200, n = 256, 16 384 chains.  Three device paths that share
    no kernel -- analytic Jacobian on the wide DMMA pair, finite differences with the incremental perturbation, finite
    differences with every perturbed model through the full forward map -- agree to the finite-difference truncation
    error, all log-posteriors are identical, and the analytic path matches the oracle on a slice."""
    d, n, B, h = 200, 256, 16384, 1e-6
    rng = np.random.default_rng(20260105)
    A = rng.standard_normal((n, d)) / math.sqrt(d)
    theta = 0.3 * rng.standard_normal(d)
    y = np.tanh(A @ theta) + 0.05 * rng.standard_cauchy(n)
    ci, pm, pci = np.full(n, 0.1), np.zeros(d), np.full(d, 2.0)
    X = theta[None, :] + 0.05 * rng.standard_normal((B, d))
    priors = [t.GaussianPrior.fromConfidenceIntervals("theta", pm, pci)]
    out = {}
    for name, diff, inc, nb in (("analytic", None, False, B), ("fd_incremental", h, True, B), ("fd_literal", h, False, 1024)):
        fm = t.NonLinearForwardModel.of("tanh", {"theta": A}, differential=diff)
        post = t.UnnormalisedPosterior(priors, [t.CauchyLikelihood(fm, y, ci)], cauchyReferenceSign=False, fdIncremental=inc)
        out[name] = post.logPdfGradBatch(X[:nb])
        post.close()
    lp, g = out["analytic"]
    assert np.all(np.isfinite(lp)) and np.all(np.isfinite(g))
    assert np.max(np.abs(out["fd_incremental"][0] - lp) / np.abs(lp)) < 1e-13          # the value never sees the Jacobian
    assert np.max(np.abs(out["fd_literal"][0] - lp[:1024]) / np.abs(lp[:1024])) < 1e-13
    # forward difference: truncation h |f''| / 2 plus rounding eps |f| / h, relative to the gradient scale
    assert rel_err_grad(out["fd_incremental"][1], g) < 2e-5
    assert rel_err_grad(out["fd_literal"][1], g[:1024]) < 2e-5
    ofm = o.NonLinearForwardModel(lambda p: np.tanh(A @ p["theta"]), None, {"theta": d}, n,
                                  jacobian=lambda p: {"theta": (1 - np.tanh(A @ p["theta"]) ** 2)[:, None] * A})
    ref = o.Posterior([o.gaussian_prior_from_ci("theta", pm, pci)], [o.cauchy_likelihood(ofm, y, ci, False)])
    for b in range(0, B, 2048):
        assert lp[b] == pytest.approx(ref.log_pdf(X[b]), rel=TOL)
        assert rel_err_grad(g[b][None], ref.log_pdf_gradient(X[b])[None]) < TOL


def test_full_size_config4_slq_million_live_points_first_rounds():
    """Config 4's pool (50-d, 10^6 live points): the first rounds at full size.  Size-independent properties of the
    threshold step: every round retires exactly K points, the retired log-likelihoods are non-decreasing within and
    across rounds, no live point sits below the last threshold, and the pool size is conserved.  (The run to
    convergence against the analytic evidence is tools/run_multi.py --only c4full, profiles/r01_config4_slq_full_*.)"""
    d, P, K, rounds = 50, 1_000_000, 62_500, 6
    y = np.random.default_rng(20260104).standard_normal(d)
    post = t.UnnormalisedPosterior([t.UniformPrior.fromBounds("theta", np.full(d, 10.0), np.full(d, -10.0))],
                                   [t.GaussianLikelihood(t.LinearForwardModel.identityOf("theta", d), y, np.full(d, 2.0))])
    cfg = t.SlqConfig(poolSize=P, abscissaNumber=4, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                      sampleParallelism=K, maxIterationCount=2_000_000_000, minIterationCount=P)
    slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=8)
    slq.runRounds(rounds)
    ll = slq.retiredLogLikelihoods()
    assert ll.size == rounds * K
    assert np.all(np.diff(ll) >= 0)                                   # ascending within rounds and across thresholds
    x, live = slq.livePoints()
    # no live point below the last threshold; AT it only exact copies of a retired point (a walker whose 40 proposals were
    # all rejected is a copy of its survivor: ties are split by index order, test_slq_duplicate_live_points_...)
    assert x.shape == (P, d) and np.all(live >= ll[-1]) and np.sum(live == ll[-1]) <= 4
    assert np.all(np.abs(x) <= 10.0)                                  # walkers never leave the prior's support
    want = -0.5 * np.sum((x[::1000] - y) ** 2, axis=1) - 0.5 * d * math.log(2 * math.pi)
    assert np.max(np.abs(live[::1000] - want) / np.abs(want)) < 1e-12  # stored log-likelihoods belong to the stored points
    lz, _ = slq.abscissaResults()
    assert np.all(np.isfinite(lz))
    slq.close()
    post.close()


def test_full_size_config1_hmc_single_chain_recovers_the_analytic_posterior():
    """Config 1 (the reference's own CPU-runnable case): 10-d, 1k observations, ONE chain, 1000 samples with the
    recipe of SURVEY.md 8(d); sample mean against the analytic posterior with the ESS-aware 5-sigma bound."""
    d, n = 10, 1000
    prob = linear_gaussian_problem(d, n, seed=20260101)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                   [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")])
    mean, cov = o.analytic_gaussian_posterior(prob["pm"], prob["pvar"], prob["A"], prob["y"], prob["var"])
    cfg = t.HmcmcConfig(stepsBetweenSamples=1, stepsInDynamicsSimulation=16, warmupStepCount=200,
                        dynamicsSimulationStepSize=0.2 * 0.1 * math.sqrt(d / n))
    hmc = t.HmcmcSampledPosterior(cfg, post, numberOfChains=1, rngSeed=1, seed={"theta": prob["theta"]})
    s = hmc.sampleChains(1000)[0]
    assert post.lastKernelPath == "stream"                            # one chain: the memory-bound streaming pass
    sd = np.sqrt(np.diag(cov))
    # The recipe's trajectory length 16 * 0.002 = 0.032 is ~pi posterior widths (0.01): a near-half-period of the
    # harmonic motion, x -> -x + O(p).  The MEAN mixes within a few samples (sign flips), the spread of |x - mean| only
    # through the momentum refreshes, so the variance estimate of 1000 samples is loose: mean at 5 sigma / sqrt(ESS)
    # with ESS >= 1000 / 8, standard deviation within a factor of two.
    ess = 1000 / 8
    assert np.max(np.abs(s.mean(0) - mean) / sd) < 5.0 / math.sqrt(ess)
    assert np.all(s.std(0) / sd > 0.4) and np.all(s.std(0) / sd < 1.6)
    att, acc, _ = hmc.stats()
    assert acc[0] / att[0] > 0.6
    hmc.close()
    post.close()
