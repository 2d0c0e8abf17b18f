"""The optimisers' HOST logic without a GPU.

The four optimisers (thylacine_b200/optimisers.py) are host loops over the batched evaluation surface
(``logPdfBatch`` / ``logPdfGradBatch`` / ``samplePriors``).  Here that surface is played by a test double built on the
oracle, so the loops themselves -- golden-section line search, Hooke & Jeeves / coordinate-slide scans, the MDS simplex
moves, conjugate-gradient updates -- run against the reference's own specs on the CPU:
    T/model/components/posterior/{HookeAndJeeves,CoordinateSlide,Mds,ConjugateGradient}OptimisedPosteriorSpec.scala
(the same specs run on the device in tests/test_optimisers_gpu.py).  The product never takes this path: the double lives
in tests/ only.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import ref_oracle as o
from thylacine_b200 import optimisers as opt

G = json.loads((Path(__file__).parent / "golden" / "reference_specs.json").read_text())


class _OraclePosterior:
    """Duck-types the slice of UnnormalisedPosterior the optimisers use."""

    def __init__(self, ref: o.Posterior, seed=0):
        self.ref = ref
        self.domainDimension = ref.domain_dimension
        self.layout = dict(ref.offsets)
        self._rng = np.random.default_rng(seed)
        self.evaluations = 0

    def flatten(self, params):
        return self.ref.flatten({k: np.asarray(v, dtype=np.float64) for k, v in params.items()})

    def unflatten(self, x):
        return self.ref.unflatten(x)

    def logPdfBatch(self, X):
        X = np.asarray(X, dtype=np.float64).reshape(-1, self.domainDimension)
        self.evaluations += X.shape[0]
        return np.array([self.ref.log_pdf(x) for x in X])

    def logPdfGradBatch(self, X):
        X = np.asarray(X, dtype=np.float64).reshape(-1, self.domainDimension)
        return self.logPdfBatch(X), np.array([self.ref.log_pdf_gradient(x) for x in X])

    def logPdfAt(self, params):
        return float(self.logPdfBatch(self.flatten(params)[None, :])[0])

    def samplePriors(self, B=1, seed=0):
        rng = np.random.default_rng(seed)
        return np.array([self.ref.sample_priors(rng) for _ in range(B)])


def _uniform_posterior():
    fu, bu, fl, bl = G["foo_uniform_prior"], G["bar_uniform_prior"], G["foo_likelihood"], G["bar_likelihood"]
    return _OraclePosterior(o.Posterior(
        [o.uniform_prior_from_bounds(fu["label"], fu["max"], fu["min"]), o.uniform_prior_from_bounds(bu["label"], bu["max"], bu["min"])],
        [o.gaussian_linear_likelihood(fl["A"], fl["y"], fl["ci"], fu["label"]),
         o.gaussian_linear_likelihood(bl["A"], bl["y"], bl["ci"], bu["label"])]))


def _max_diff(arg):
    return max(np.max(np.abs(np.asarray(arg["fooniform"]) - [1, 2])), np.max(np.abs(np.asarray(arg["barniform"]) - [5])))


def test_hooke_and_jeeves_host_loop():
    post = _uniform_posterior()
    lp, arg = opt.HookeAndJeevesOptimisedPosterior(opt.HookeAndJeevesConfig(1e-7, 100), post, rngSeed=1).findMaximumLogPdf({})
    assert _max_diff(arg) < 1e-5 and np.isfinite(lp) and post.evaluations > 0


def test_coordinate_slide_host_loop_and_telemetry():
    ups = []
    post = _uniform_posterior()
    lp, arg = opt.CoordinateSlideOptimisedPosterior(opt.CoordinateSlideConfig(1e-7, 1e-10, 2.0, 100), post,
                                                    iterationUpdateCallback=ups.append, rngSeed=2).findMaximumLogPdf({})
    assert _max_diff(arg) < 1e-5
    assert ups and ups[-1].prefix == "Coordinate Slide" and ups[-1].maxLogPdf == pytest.approx(lp, abs=1e-6)


def test_mds_host_loop_evaluates_whole_simplices():
    done = []
    post = _uniform_posterior()
    lp, arg = opt.MdsOptimisedPosterior(opt.MdsConfig(1e-15, 2.0, 0.5, 100), post, isConvergedCallback=lambda: done.append(1),
                                        rngSeed=3).findMaximumLogPdf({})
    assert _max_diff(arg) < 1e-5 and done == [1]
    v = opt.MdsOptimisedPosterior.unitRegularCenteredOn(np.array([0.3, -1.0, 2.0]))          # ModelParameterSimplex.scala:93-120
    dist = [np.linalg.norm(v[i] - v[j]) for i in range(4) for j in range(i)]
    assert np.allclose(dist, dist[0]) and np.allclose(v.mean(axis=0), [0.3, -1.0, 2.0])


def test_conjugate_gradient_host_loop():
    post = _uniform_posterior()
    cg = opt.ConjugateGradientOptimisedPosterior(opt.ConjugateGradientConfig(1e-20, 1e-10, 2.0, 60), post, rngSeed=4)
    lp, arg = cg.findMaximumLogPdf({})
    assert _max_diff(arg) < 1e-1                                              # the reference's own tolerance (spec :31-39)
    pdf, arg2 = cg.findMaximumPdf({"fooniform": [0.5, 0.5], "barniform": [4.0]})
    assert pdf == pytest.approx(np.exp(post.logPdfAt(arg2)), rel=1e-9)
