"""The reference's optimiser and analytic-posterior specs, through the drop-in surface.

    T/model/components/posterior/HookeAndJeevesOptimisedPosteriorSpec.scala:30-38   argmax within 1e-5
    T/model/components/posterior/CoordinateSlideOptimisedPosteriorSpec.scala:30-38  argmax within 1e-5
    T/model/components/posterior/ConjugateGradientOptimisedPosteriorSpec.scala:31-39 argmax within 1e-1
    T/model/components/posterior/MdsOptimisedPosteriorSpec.scala:30-38              (tolerance 1e5: vacuous; 1e-5 here)
    T/model/components/posterior/GaussianAnalyticPosteriorSpec.scala:34-50          mean within 0.1, covariance within 0.01
Fixture: T/model/components/ComponentFixture.scala:34-211 (uniform priors fooniform / barniform, the two linear
likelihoods re-labelled onto them; configs :146-200).
"""
import json
from pathlib import Path

import numpy as np
import pytest

import thylacine_b200 as t
from oracle import ref_oracle as o

pytestmark = pytest.mark.gpu
G = json.loads((Path(__file__).parent / "golden" / "reference_specs.json").read_text())


def _uniform_posterior():
    fu, bu = G["foo_uniform_prior"], G["bar_uniform_prior"]
    fl, bl = G["foo_likelihood"], G["bar_likelihood"]
    return t.UnnormalisedPosterior(
        [t.UniformPrior.fromBounds(fu["label"], fu["max"], fu["min"]), t.UniformPrior.fromBounds(bu["label"], bu["max"], bu["min"])],
        [t.GaussianLinearLikelihood.of(fl["A"], fl["y"], fl["ci"], fu["label"]),
         t.GaussianLinearLikelihood.of(bl["A"], bl["y"], bl["ci"], bu["label"])])


def _max_diff(arg):
    return max(np.max(np.abs(np.asarray(arg["fooniform"]) - [1, 2])), np.max(np.abs(np.asarray(arg["barniform"]) - [5])))


def test_HookeAndJeevesOptimisedPosteriorSpec():
    opt = t.HookeAndJeevesOptimisedPosterior(t.HookeAndJeevesConfig(1e-7, 100), _uniform_posterior(), rngSeed=1)
    lp, arg = opt.findMaximumLogPdf({})
    assert _max_diff(arg) < 1e-5


def test_CoordinateSlideOptimisedPosteriorSpec():
    ups = []
    opt = t.CoordinateSlideOptimisedPosterior(t.CoordinateSlideConfig(1e-7, 1e-10, 2.0, 100), _uniform_posterior(),
                                              iterationUpdateCallback=ups.append, rngSeed=2)
    lp, arg = opt.findMaximumLogPdf({})
    assert _max_diff(arg) < 1e-5
    assert ups and ups[-1].prefix == "Coordinate Slide" and ups[-1].maxLogPdf == pytest.approx(lp, abs=1e-6)


def test_MdsOptimisedPosteriorSpec():
    done = []
    opt = t.MdsOptimisedPosterior(t.MdsConfig(1e-15, 2.0, 0.5, 100), _uniform_posterior(), isConvergedCallback=lambda: done.append(1),
                                  rngSeed=3)
    lp, arg = opt.findMaximumLogPdf({})
    assert _max_diff(arg) < 1e-5 and done == [1]
    # the starting simplex is regular with unit edges (ModelParameterSimplex.scala:93-120)
    v = t.MdsOptimisedPosterior.unitRegularCenteredOn(np.array([0.3, -1.0, 2.0]))
    dist = [np.linalg.norm(v[i] - v[j]) for i in range(4) for j in range(i)]
    assert np.allclose(dist, dist[0]) and np.allclose(v.mean(axis=0), [0.3, -1.0, 2.0])


def test_ConjugateGradientOptimisedPosteriorSpec():
    # the reference runs minimumNumberOfIterations = 10000 with threshold 1e-20 and accepts 1e-1 (the fixture's two
    # likelihoods differ by 10^6 in precision, so conjugate gradients crawl); 60 iterations already meet that bar
    opt = t.ConjugateGradientOptimisedPosterior(t.ConjugateGradientConfig(1e-20, 1e-10, 2.0, 60), _uniform_posterior(), rngSeed=4)
    lp, arg = opt.findMaximumLogPdf({})
    assert _max_diff(arg) < 1e-1
    pdf, arg2 = opt.findMaximumPdf({"fooniform": [0.5, 0.5], "barniform": [4.0]})          # explicit starting point
    assert pdf == pytest.approx(np.exp(opt.posterior.logPdfAt(arg2)), rel=1e-9)


def test_optimisers_agree_with_the_analytic_maximum_on_a_larger_model():
    rng = np.random.default_rng(8)
    d, n = 6, 40
    A = rng.standard_normal((n, d))
    y = A @ rng.standard_normal(d) + 0.1 * rng.standard_normal(n)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("p", np.zeros(d), np.full(d, 6.0))],
                                   [t.GaussianLinearLikelihood.of(A, y, np.full(n, 0.2), "p")])
    mean, _ = o.analytic_gaussian_posterior(np.zeros(d), np.full(d, 9.0), A, y, np.full(n, 0.01))
    for opt in (t.CoordinateSlideOptimisedPosterior(t.CoordinateSlideConfig(1e-7, 1e-9, 2.0, 64), post, rngSeed=5),
                t.MdsOptimisedPosterior(t.MdsConfig(1e-9, 2.0, 0.5, 64), post, rngSeed=6),
                t.ConjugateGradientOptimisedPosterior(t.ConjugateGradientConfig(1e-14, 1e-10, 2.0, 12), post, rngSeed=7)):
        _, arg = opt.findMaximumLogPdf({})
        assert np.max(np.abs(arg["p"] - mean)) < 1e-5, type(opt).__name__


def test_GaussianAnalyticPosteriorSpec():
    fp, bp, fl, bl = G["foo_prior"], G["bar_prior"], G["foo_likelihood"], G["bar_likelihood"]
    post = t.GaussianAnalyticPosterior(
        [t.GaussianPrior.fromConfidenceIntervals(fp["label"], fp["mean"], fp["ci"]),
         t.GaussianPrior.fromConfidenceIntervals(bp["label"], bp["mean"], bp["ci"])],
        [t.GaussianLinearLikelihood.of(fl["A"], fl["y"], fl["ci"], fp["label"]),
         t.GaussianLinearLikelihood.of(bl["A"], bl["y"], bl["ci"], bp["label"])])
    post.init()
    assert max(np.max(np.abs(post.mean["foo"] - [1, 2])), np.max(np.abs(post.mean["bar"] - [5]))) < 0.1
    assert np.max(np.abs(post.covarianceStridedVector)) < 0.01 and post.covarianceStridedVector.size == 9
    # and exactly the oracle's analytic posterior (label-sorted order: bar, foo)
    A = np.zeros((4, 3))
    A[:2, 1:] = fl["A"]
    A[2:, :1] = bl["A"]
    var = np.concatenate([o.ci_to_variance(fl["ci"]), o.ci_to_variance(bl["ci"])])
    pm = np.array(bp["mean"] + fp["mean"], dtype=float)
    pv = np.concatenate([o.ci_to_variance(bp["ci"]), o.ci_to_variance(fp["ci"])])
    mean, cov = o.analytic_gaussian_posterior(pm, pv, A, np.array(fl["y"] + bl["y"], dtype=float), var)
    assert np.max(np.abs(post.meanVector - mean)) < 1e-9 and np.max(np.abs(post.covariance - cov)) < 1e-9 * np.max(np.abs(cov)) + 1e-18
    xs = post.sampleArray(20000, rngSeed=1)
    assert np.max(np.abs(xs.mean(axis=0) - mean)) < 5 * np.sqrt(np.max(np.diag(cov)) / 20000)
    assert post.maxLogPdf == pytest.approx(post.logPdfAt(post.mean))
    with pytest.raises(AssertionError):
        t.GaussianAnalyticPosterior([t.UniformPrior.fromBounds("u", [1], [0])], [])
