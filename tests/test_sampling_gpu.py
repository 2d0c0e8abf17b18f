"""Prior draws: Posterior.samplePriors, Posterior.scala:64-65, with
GaussianPrior.scala:47-48 (mu + chol(Sigma) z), UniformDistribution.scala:88-93 (lo + u (hi - lo)),
CauchyDistribution.scala:84-87 (MVN(mu, Sigma / chi2_1)).  The reference's generators are unseeded, so parity is
(i) exact replay of the device's counter-based draws on the host and (ii) the distributions themselves."""
import math

import numpy as np
import pytest

import thylacine_b200 as t
import philox_ref as ph

pytestmark = pytest.mark.gpu


def _posterior():
    cov = np.array([[2.0, 0.6, 0.0], [0.6, 1.0, -0.3], [0.0, -0.3, 0.5]])
    priors = [t.GaussianPrior.fromConfidenceIntervals("a", [1.0, -2.0], [3.0, 0.5]),             # sigma = 1.5, 0.25
              t.UniformPrior.fromBounds("b", [5.0, 0.0], [1.0, -4.0]),
              t.CauchyPrior("c", [10.0], [2.0]),                                                 # scale sigma = 1
              t.GaussianPrior.fromCovarianceMatrix("d", [0.0, 1.0, 2.0], cov)]
    return t.UnnormalisedPosterior(priors), cov


def test_prior_draws_replay_on_the_host():
    post, cov = _posterior()
    seed, B = 987654321, 7
    X = post.samplePriors(B, seed)
    assert np.array_equal(X, post.samplePriors(B, seed))                       # counter-based: reproducible
    assert not np.array_equal(X, post.samplePriors(B, seed + 1))
    oa, ob, oc, od = (post.layout[k][0] for k in "abcd")
    L = np.linalg.cholesky(cov)
    for b in range(B):
        for i, (mu, sd) in enumerate([(1.0, 1.5), (-2.0, 0.25)]):
            assert X[b, oa + i] == pytest.approx(mu + sd * ph.normal(seed, b, oa + i, ph.RNG_PRIOR, 0), rel=1e-14, abs=1e-14)
        for i, (lo, hi) in enumerate([(1.0, 5.0), (-4.0, 0.0)]):
            u, _ = ph.uniform2(seed, b, ob + i, ph.RNG_PRIOR, 0)
            assert X[b, ob + i] == pytest.approx(lo + u * (hi - lo), rel=1e-15, abs=1e-15)
        z = np.array([ph.normal(seed, b, od + k, ph.RNG_PRIOR, 0) for k in range(3)])
        assert np.allclose(X[b, od:od + 3], np.array([0.0, 1.0, 2.0]) + L @ z, rtol=1e-13, atol=1e-13)
    post.close()


def test_prior_draws_have_the_priors_distributions():
    post, cov = _posterior()
    B = 400_000
    X = post.samplePriors(B, 2026)
    oa, ob, oc, od = (post.layout[k][0] for k in "abcd")
    se = 5.0 / math.sqrt(B)
    assert abs(X[:, oa].mean() - 1.0) < 1.5 * se and abs(X[:, oa].std() / 1.5 - 1.0) < se
    assert abs(X[:, oa + 1].mean() + 2.0) < 0.25 * se and abs(X[:, oa + 1].std() / 0.25 - 1.0) < se
    for i, (lo, hi) in enumerate([(1.0, 5.0), (-4.0, 0.0)]):
        v = X[:, ob + i]
        assert v.min() >= lo and v.max() < hi                                   # half-open box
        assert abs(v.mean() - 0.5 * (lo + hi)) < (hi - lo) / math.sqrt(12) * se
        assert abs(v.var() / ((hi - lo) ** 2 / 12) - 1.0) < 3 * se
    c = X[:, oc]                                                                # Cauchy(10, 1): quartiles at 10 -+ 1
    q1, q2, q3 = np.quantile(c, [0.25, 0.5, 0.75])
    assert abs(q2 - 10.0) < 0.02 and abs(q1 - 9.0) < 0.03 and abs(q3 - 11.0) < 0.03
    assert np.mean(np.abs(c - 10.0) > 100.0) == pytest.approx(2 / math.pi * math.atan(1 / 100.0), rel=0.15)   # heavy tails
    dsmp = X[:, od:od + 3]
    assert np.max(np.abs(dsmp.mean(0) - np.array([0.0, 1.0, 2.0]))) < 1.5 * se
    assert np.max(np.abs(np.cov(dsmp.T) - cov)) < 4 * se * 2.0
    post.close()
