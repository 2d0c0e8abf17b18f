"""SLQ (nested sampling) on the device vs the reference's quadrature and vs analytic evidence.

Reference: /root/reference/src/main/scala/thylacine/model/integration/slq/
    SlqEngine.scala:120-197,249-277   live pool / threshold / convergence
    QuadratureAbscissa.scala:33-73    prior-mass simulation, trapezoid weights
    QuadratureIntegrator.scala:29-62  evidence and negative entropy per abscissa
The reference has no SLQ tests (README.md:114-118) and an unseeded RNG, so parity is defined by
  (1) REPLAY of the quadrature: the retired log-likelihoods the device produced, pushed through the oracle's
      trapezoid weights / log-sum-exp evidence with the very uniforms the device's counter RNG drew, and
  (2) the analytic evidence of Gaussian models, tolerance 5 sqrt(H / P) + the spread over abscissas.
"""
import math

import numpy as np
import pytest

import thylacine_b200 as t
from thylacine_b200 import ThylacineError
from oracle import ref_oracle as o
import philox_ref as ph

pytestmark = pytest.mark.gpu

RNG_SLQ_SHRINK = 10


def _identity_model(d, seed, uniform=True, half_width=10.0, prior_sigma=5.0):
    rng = np.random.default_rng(seed)
    y = rng.standard_normal(d)
    ci = np.full(d, 2.0)   # sigma = 1
    if uniform:
        prior = t.UniformPrior.fromBounds("theta", np.full(d, half_width), np.full(d, -half_width))
    else:
        prior = t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 2.0 * prior_sigma))
    lik = t.GaussianLikelihood(t.LinearForwardModel.identityOf("theta", d), y, ci)
    post = t.UnnormalisedPosterior([prior], [lik])
    if uniform:
        want = o.analytic_uniform_box_log_evidence(np.full(d, -half_width), np.full(d, half_width), np.eye(d), y, np.ones(d))
    else:
        want = o.analytic_gaussian_log_evidence(np.zeros(d), np.full(d, prior_sigma ** 2), np.eye(d), y, np.ones(d))
    return post, y, want


def _replay_abscissa(ll, P, K, live, seed, a):
    """QuadratureAbscissa + QuadratureIntegrator for abscissa a with the device's uniforms (host replay)."""
    N = ll.size
    lx = np.empty(N)
    cur = 0.0
    for i in range(N):
        lx[i] = cur
        j = (i % K) if i < live else (i - live)
        u0, _ = ph.uniform2(seed, a, 0, RNG_SLQ_SHRINK, i)
        cur += math.log(1.0 - u0) / (P - j)
    X = np.exp(lx)
    w = np.empty(N)
    w[0] = 0.5 * (X[0] - X[1])
    w[1:-1] = 0.5 * (X[:-2] - X[2:])
    w[-1] = 0.5 * (X[-2] - X[-1])
    lz = o.log_sum_exp_evidence(ll, w)
    keep = w > 0
    p = np.exp(ll[keep] + np.log(w[keep]) - lz)
    return lz, float(np.sum(p * ll[keep]) - lz)


def test_slq_quadrature_replays_reference_formulas():
    post, _, _ = _identity_model(3, seed=5)
    P, K, A, seed = 256, 16, 3, 99
    cfg = t.SlqConfig(poolSize=P, abscissaNumber=A, domainScalingIncrement=0.1, targetAcceptanceProbability=0.3,
                      sampleParallelism=K, maxIterationCount=1600, minIterationCount=100)
    slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=seed)
    slq.rebuildSampleSimulation()
    ll = slq.retiredLogLikelihoods()
    live = int(slq.stats.iterations)
    assert ll.size == live + P == slq.stats.total_retired
    assert live % K == 0 or live == cfg.maxIterationCount
    # thresholds only ever rise: the whole retirement sequence is sorted (size-independent property)
    assert np.all(np.diff(ll) >= 0)
    lz, hh = slq.abscissaResults()
    for a in range(A):
        want_lz, want_h = _replay_abscissa(ll, P, K, live, seed, a)
        assert abs(lz[a] - want_lz) < 1e-9 * max(1.0, abs(want_lz))
        assert abs(hh[a] - want_h) < 1e-8 * max(1.0, abs(want_h))
    assert abs(slq.logEvidence - lz.mean()) < 1e-12


def test_slq_exported_abscissas_and_integrate_follow_the_reference_quadrature():
    """thy_slq_retired_log_x replays QuadratureAbscissa (QuadratureAbscissa.scala:33-73) for one abscissa; it must equal
    the host replay with the device's uniforms, and thy_slq_integrate must equal QuadratureIntegrator.getIntegrationStats
    (QuadratureIntegrator.scala:64-83) restated in numpy on those abscissas (both the literal statistic and the plain
    quadrature)."""
    post, _, _ = _identity_model(3, seed=8)
    P, K, A, seed = 300, 20, 4, 5
    cfg = t.SlqConfig(poolSize=P, abscissaNumber=A, domainScalingIncrement=0.1, targetAcceptanceProbability=0.3,
                      sampleParallelism=K, maxIterationCount=1510, minIterationCount=100)   # last live round is partial
    slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=seed)
    slq.rebuildSampleSimulation()
    ll = slq.retiredLogLikelihoods()
    live, N = int(slq.stats.iterations), ll.size
    assert live == 1510 and N == live + P
    W = []
    for a in range(A):
        lx = slq.retiredLogX(a)
        want = np.empty(N)
        cur = 0.0
        for i in range(N):
            want[i] = cur
            j = (i % K) if i < live else (i - live)
            u0, _ = ph.uniform2(seed, a, 0, RNG_SLQ_SHRINK, i)
            cur += math.log(1.0 - u0) / (P - j)
        assert np.max(np.abs(lx - want)) < 1e-12 * max(1.0, np.max(np.abs(want)))
        X = np.exp(lx)
        w = np.empty(N)
        w[0] = 0.5 * (X[0] - X[1])
        w[1:-1] = 0.5 * (X[:-2] - X[2:])
        w[-1] = 0.5 * (X[-2] - X[-1])
        W.append(w)
    f = lambda v: v * v + 0.25
    for shift in (float(ll.max()), float("nan")):
        s_ref = 0.5 * (ll.max() + ll.min()) if math.isnan(shift) else shift
        e = np.exp(ll - s_ref)
        I = np.array([np.sum(w * f(e)) for w in W])
        Z = np.array([np.sum(w * e) for w in W])
        assert slq.integrate(f, logShift=shift) == pytest.approx(I.mean(), rel=1e-11)
        assert slq.integrate(f, logShift=shift, reference=True) == pytest.approx(np.mean(I / Z - np.log(Z)), rel=1e-11)
    slq.close()


@pytest.mark.parametrize("P,K,d", [(64, 1, 2), (1000, 500, 3), (4099, 37, 5), (2048, 1024, 2)])
def test_slq_radix_select_retires_exactly_the_k_lowest(P, K, d):
    """The threshold step (slq_select.cu) against numpy: after ONE round the retired values are the K smallest of the
    initial pool, in order, every survivor is above them, and the K replacements sit above the threshold."""
    post, _, _ = _identity_model(d, seed=P + K)
    cfg = t.SlqConfig(poolSize=P, abscissaNumber=1, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                      sampleParallelism=K, maxIterationCount=100 * P, minIterationCount=50 * P)
    slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=3, constrainedWalkSteps=5)
    _, ll0 = slq.livePoints()
    slq.runRounds(1)
    _, ll1 = slq.livePoints()
    ret = slq.retiredLogLikelihoods()
    want = np.sort(ll0)[:K]
    assert np.array_equal(ret, want)
    thr = want[-1]
    changed = ll1 != ll0
    assert changed.sum() <= K and np.all(ll0[changed] <= thr)
    assert np.all(ll1 > thr) or np.sum(ll1 <= thr) == 0
    slq.close()


@pytest.mark.parametrize("uniform", [True, False])
def test_slq_log_evidence_matches_analytic(uniform):
    d, P, K = 6, 8000, 400
    post, y, want = _identity_model(d, seed=11, uniform=uniform)
    cfg = t.SlqConfig(poolSize=P, abscissaNumber=8, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                      sampleParallelism=K, maxIterationCount=40 * P, minIterationCount=P)
    slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=2026, constrainedWalkSteps=30)
    slq.rebuildSampleSimulation()
    H = slq.stats.neg_entropy_mean
    assert 1.0 < H < 40.0
    tol = 5.0 * math.sqrt(H / P) + 3.0 * slq.stats.log_evidence_std / math.sqrt(8)
    assert abs(slq.logEvidence - want) < tol, (slq.logEvidence, want, tol)
    assert 0.05 < slq.stats.walk_acceptance < 0.95


def test_slq_linear_model_gaussian_prior_and_posterior_samples():
    rng = np.random.default_rng(3)
    d, n, P, K = 4, 30, 6000, 300
    A = rng.standard_normal((n, d))
    theta = rng.standard_normal(d)
    y = A @ theta + 0.5 * rng.standard_normal(n)
    ci = np.full(n, 1.0)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 6.0))],
                                   [t.GaussianLinearLikelihood.of(A, y, ci, "theta")])
    want = o.analytic_gaussian_log_evidence(np.zeros(d), np.full(d, 9.0), A, y, np.full(n, 0.25))
    mean, cov = o.analytic_gaussian_posterior(np.zeros(d), np.full(d, 9.0), A, y, np.full(n, 0.25))
    cfg = t.SlqConfig(poolSize=P, abscissaNumber=6, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                      sampleParallelism=K, maxIterationCount=60 * P, minIterationCount=P)
    slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=7, constrainedWalkSteps=30, keepRetiredPoints=True)
    slq.rebuildSampleSimulation()
    H = slq.stats.neg_entropy_mean
    tol = 5.0 * math.sqrt(H / P) + 3.0 * slq.stats.log_evidence_std / math.sqrt(6)
    assert abs(slq.logEvidence - want) < tol, (slq.logEvidence, want, tol)
    xs = np.array([s["theta"] for s in slq.sample(3000)])
    sd = np.sqrt(np.diag(cov))
    # posterior resampling of retired points (SamplingSimulation.scala:41-101): ESS is far below 3000 draws
    assert np.all(np.abs(xs.mean(axis=0) - mean) < 0.5 * sd)
    assert np.all(np.abs(xs.std(axis=0) / sd - 1.0) < 0.35)
    # integrate(identity) is the evidence scaled by exp(-max l)  (SlqEngine.scala:460-466)
    ll = slq.retiredLogLikelihoods()
    got = math.log(slq.integrate(lambda L: L)) + ll.max()
    assert abs(got - want) < tol + 0.05


@pytest.mark.parametrize("case", ["identity_uniform", "linear_gaussian", "linear_cauchy", "linear_gaussian_wide"])
def test_slq_fused_walk_kernel_matches_the_per_step_kernels(case):
    """The one-kernel replacement step (slq_walk.cu: spawn + M steps + commit, walker state in DMMA fragments) against
    the per-step kernels: same random numbers, same proposals; the log-likelihoods differ only in summation order, so
    after one round the two live pools coincide except where a 1e-16 difference flips a constraint test."""
    rng = np.random.default_rng(5)
    if case == "identity_uniform":
        d = 7
        post, y, _ = _identity_model(d, seed=3, uniform=True)
    else:
        d, n = (100, 120) if case == "linear_gaussian_wide" else (21, 45)   # wide: the 8-warp, 255-register instantiation
        A = rng.standard_normal((n, d)) / math.sqrt(d)
        y = A @ rng.standard_normal(d) + 0.3 * rng.standard_normal(n)
        lik = (t.CauchyLikelihood.of(A, y, np.full(n, 0.6), "theta") if case == "linear_cauchy"
               else t.GaussianLinearLikelihood.of(A, y, np.full(n, 0.6), "theta"))
        post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 4.0))], [lik])
    P, K = 4096, 1000                                                        # 1000 walkers: a ragged last group of 8
    cfg = t.SlqConfig(poolSize=P, abscissaNumber=2, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                      sampleParallelism=K, maxIterationCount=50 * P, minIterationCount=P)
    pools = []
    for fused in (True, False):
        slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=77, constrainedWalkSteps=12, fusedWalk=fused)
        slq.runRounds(1)
        x, ll = slq.livePoints()
        pools.append((x, ll, slq.retiredLogLikelihoods()))
        slq.close()
    (xa, la, ra), (xb, lb, rb) = pools
    assert np.array_equal(ra, rb)                                            # the first round retires prior draws
    same = np.all(xa == xb, axis=1)
    assert same.mean() > 0.995, same.mean()
    assert np.max(np.abs(la[same] - lb[same]) / np.maximum(1.0, np.abs(lb[same]))) < 1e-12


@pytest.mark.parametrize("batch", [1, 8])
def test_slq_reference_proposals_replay_the_cubical_cover(batch):
    """thy_slq_config.referenceProposals: the reference's own proposal mechanism (PointInCubeCollection.scala:60-111,
    PointInCube.scala:84-117, PointInInterval.scala:40-109), one-at-a-time pool update (SlqEngine.scala:216-247) and
    rebuild schedule (QuadratureDomainTelemetry.scala:48-49), replayed draw for draw by oracle/slq_cubical_cover.py with
    the device engine's counter-based uniforms: same retirement sequence, same final pool, same number of proposals."""
    from oracle import slq_cubical_cover as cc
    rng = np.random.default_rng(4)
    d, n, P = 3, 12, 40
    A = rng.standard_normal((n, d))
    y = A @ np.array([0.5, -0.3, 0.8]) + 0.3 * rng.standard_normal(n)
    pri = [t.UniformPrior.fromBounds("theta", np.full(d, 4.0), np.full(d, -4.0))]
    lik = [t.GaussianLinearLikelihood.of(A, y, np.full(n, 0.6), "theta")]
    post = t.UnnormalisedPosterior(pri, lik)
    ref = o.Posterior([o.uniform_prior_from_bounds("theta", np.full(d, 4.0), np.full(d, -4.0))],
                      [o.gaussian_linear_likelihood(A, y, np.full(n, 0.6), "theta")])
    seed, max_it = 321, 700
    cfg = t.SlqConfig(poolSize=P, abscissaNumber=3, domainScalingIncrement=0.1, targetAcceptanceProbability=0.3,
                      sampleParallelism=batch, maxIterationCount=max_it, minIterationCount=max_it)
    slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=seed, referenceProposals=True)
    x0, _ = slq.livePoints()                                              # the prior draws the run starts from
    slq.rebuildSampleSimulation()
    got = slq.retiredLogLikelihoods()
    live = int(slq.stats.iterations)
    RNG_SLQ_REFERENCE = 12
    uniform = lambda stream, k: ph.uniform2(seed, stream, 0, RNG_SLQ_REFERENCE, k)[0]
    want, pool_x, pool_lp, nprop = cc.run_reference_slq(x0, ref.log_pdf, uniform,
                                                        max_iterations=max_it, batch=batch)
    assert live == want.size == max_it and got.size == live + P
    assert np.max(np.abs(got[:live] - want) / np.maximum(1.0, np.abs(want))) < 1e-11
    assert np.all(np.diff(got) >= 0) or np.all(np.diff(got[:live]) >= 0)
    x1, lp1 = slq.livePoints()
    assert np.array_equal(x1, pool_x) and np.max(np.abs(lp1 - pool_lp) / np.maximum(1.0, np.abs(pool_lp))) < 1e-11
    assert int(slq.stats.rounds) == nprop                                 # proposals made (telemetry), rebuild discards included
    lz, _ = slq.abscissaResults()
    for a in range(3):                                                    # quadrature: one retirement per step
        want_lz, _ = _replay_abscissa(got, P, 1, live, seed, a)
        assert abs(lz[a] - want_lz) < 1e-9 * max(1.0, abs(want_lz))
    slq.close()
    post.close()


def test_slq_duplicate_live_points_do_not_break_the_threshold_step():
    """A walker that never moves is an exact copy of a survivor (the reference nudges duplicates by one ulp,
    SlqEngine.scala:107-118); with one walk step per replacement most replacements are copies, so log-likelihood ties
    straddle the threshold constantly.  Exactly K points must still be retired per round."""
    post, _, _ = _identity_model(8, seed=3)
    P, K = 4096, 1024
    cfg = t.SlqConfig(poolSize=P, abscissaNumber=2, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                      sampleParallelism=K, maxIterationCount=60 * K, minIterationCount=60 * K)
    slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=17, constrainedWalkSteps=1)
    slq.rebuildSampleSimulation()
    ll = slq.retiredLogLikelihoods()
    assert slq.stats.iterations == 60 * K and ll.size == 61 * K + (P - K)
    assert np.all(np.diff(ll) >= 0)
    assert np.sum(np.diff(ll) == 0) > 0, "the scenario is meant to produce ties"


@pytest.mark.parametrize("P,K,steps", [(4096, 1024, 1), (20000, 5000, 20), (3000, 37, 20)])
def test_slq_bucket_sort_of_the_retired_values_equals_the_library_sort(P, K, steps, monkeypatch):
    """A round's retired values are sorted by the bucket + rank kernels of slq_sort.cu (no library call in a round);
    THY_SLQ_CUB_SORT=1 routes them through cub's radix sort instead.  Both are exact sorts of the same multiset, so the
    whole retirement sequence -- heavy ties included (one walk step: most replacements are copies) -- is bit-identical."""
    post, _, _ = _identity_model(6, seed=5)
    cfg = t.SlqConfig(poolSize=P, abscissaNumber=2, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                      sampleParallelism=K, maxIterationCount=25 * K, minIterationCount=25 * K)
    got = []
    for cub in ("0", "1"):
        monkeypatch.setenv("THY_SLQ_CUB_SORT", cub)
        slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=23, constrainedWalkSteps=steps)
        slq.rebuildSampleSimulation()
        got.append((slq.retiredLogLikelihoods().copy(), slq.abscissaResults()[0].copy()))
        slq.close()
    assert got[0][0].size == got[1][0].size and np.array_equal(got[0][0], got[1][0])
    assert np.all(np.diff(got[0][0]) >= 0)
    assert np.array_equal(got[0][1], got[1][1])
    post.close()


def test_slq_argument_checks_and_call_order():
    post, _, _ = _identity_model(2, seed=1)
    bad = t.SlqConfig(poolSize=2, abscissaNumber=1, domainScalingIncrement=0.1, targetAcceptanceProbability=0.3,
                      sampleParallelism=1, maxIterationCount=10, minIterationCount=1)
    with pytest.raises(ThylacineError) as e:
        t.SlqIntegratedPosterior.of(bad, post)
    assert e.value.status == t.capi.THY_ERR_INVALID_ARGUMENT
    ok = t.SlqConfig(poolSize=64, abscissaNumber=2, domainScalingIncrement=0.1, targetAcceptanceProbability=0.3,
                     sampleParallelism=4, maxIterationCount=128, minIterationCount=10)
    slq = t.SlqIntegratedPosterior.of(ok, post, seedsSpec=[{"theta": [0.1, 0.2]}])
    with pytest.raises(ThylacineError) as e:
        slq.sample(1)
    assert e.value.status == t.capi.THY_ERR_STATE
    assert slq.runRounds(3) is False
    slq.rebuildSampleSimulation()
    with pytest.raises(ThylacineError) as e:
        slq.sample(1)          # retired points were not kept
    assert e.value.status == t.capi.THY_ERR_UNSUPPORTED
    priors_only = t.UnnormalisedPosterior([t.UniformPrior.fromBounds("theta", [1, 1], [0, 0])])
    with pytest.raises(ThylacineError):
        t.SlqIntegratedPosterior.of(ok, priors_only)


@pytest.mark.parametrize("d,n", [(7, 1000), (100, 5000), (33, 31)])
def test_sample_moments_match_numpy(d, n):
    import torch
    rng = np.random.default_rng(d)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.ones(d))])
    X = 1.0 + 0.1 * rng.standard_normal((n, d))
    X[::17, 0] = np.nan            # unfilled sample slots are skipped
    keep = ~np.isnan(X[:, 0])
    mom = t.SampleMoments(post)
    xd = torch.from_numpy(X).cuda()
    half = n // 2
    mom.add(xd[:half].contiguous())
    mom.add(xd[half:].contiguous())
    cnt, mean, cov = mom.finish()
    assert cnt == keep.sum()
    assert np.max(np.abs(mean - X[keep].mean(axis=0))) < 1e-13
    want = np.cov(X[keep].T)
    assert np.max(np.abs(cov - want)) < 1e-12 * max(1.0, np.max(np.abs(want))) + 1e-13
