"""``GaussianAnalyticPosterior`` as a thin client of the batched device gradient.

Reference: /root/reference/src/main/scala/thylacine/model/components/posterior/GaussianAnalyticPosterior.scala
    :141-169  Lambda = Sigma_p^-1 + A^T Sigma_d^-1 A,  mean = Lambda^-1 (Sigma_p^-1 mu_p + A^T Sigma_d^-1 y)
    :72-116   maxLogPdf, entropy, mean, covarianceStridedVector, logPdfAt, sample, init
The precision Lambda and the information vector are assembled by the C library (``thy_model_gaussian_posterior``,
csrc/collapse.cu): the likelihood part A^T Sigma_d^-1 A comes from the device's own forward + adjoint kernels applied to
the homogeneous problem at the unit vectors (no cancellation against the information vector), the prior part and the
d x d Cholesky factorisation are host-side one-offs inside the library.  The reference concatenates prior blocks in Set
iteration order while its transform columns use sorted order (latent bug); here everything is in
label-sorted order.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np

from . import _capi
from .api import UnnormalisedPosterior, _LikelihoodSpec, _PriorSpec


class GaussianAnalyticPosterior:
    def __init__(self, priors: Sequence[_PriorSpec], likelihoods: Sequence[_LikelihoodSpec]):
        for p in priors:
            assert p.kind in ("gaussian_ci", "gaussian_cov"), "GaussianAnalyticPosterior takes GaussianPriors (:39)"
        for lk in likelihoods:
            fm = lk.forwardModel
            assert lk.distribution == _capi.DIST_GAUSSIAN and fm.link == _capi.LINK_IDENTITY and \
                fm.jacobian == _capi.JAC_CONSTANT, "GaussianAnalyticPosterior takes GaussianLinearLikelihoods (:40)"
        self.posterior = UnnormalisedPosterior(priors, likelihoods)
        self._ready = False

    def init(self) -> None:
        """Builds the posterior distribution (the reference's ``init`` forces its lazy ``rawDistribution``, :116-119)."""
        self._mean, self._cov, lam, self._log_evidence, self._max_logpdf = self.posterior.gaussianPosterior()
        self._chol = np.linalg.cholesky(lam)
        self._precision = lam
        self._ready = True

    def _need(self):
        if not self._ready:
            self.init()

    @property
    def mean(self) -> Dict[str, np.ndarray]:
        self._need()
        return {k: np.array(v) for k, v in self.posterior.unflatten(self._mean).items()}

    @property
    def meanVector(self) -> np.ndarray:
        self._need()
        return self._mean.copy()

    @property
    def covariance(self) -> np.ndarray:
        self._need()
        return self._cov.copy()

    @property
    def covarianceStridedVector(self) -> np.ndarray:
        return self.covariance.ravel()

    @property
    def entropy(self) -> float:
        """Differential entropy of the posterior Gaussian (Breeze ``MultivariateGaussian.entropy``, :74)."""
        self._need()
        d = self._mean.size
        logdet_cov = -2.0 * float(np.sum(np.log(np.diag(self._chol))))
        return 0.5 * (d * (1.0 + math.log(2.0 * math.pi)) + logdet_cov)

    def logPdfAt(self, params) -> float:
        """Normalised log-density of the analytic posterior (:102-105)."""
        self._need()
        x = self.posterior.flatten(params) - self._mean
        d = x.size
        quad = float(x @ self._precision @ x)
        return -0.5 * quad + float(np.sum(np.log(np.diag(self._chol)))) - 0.5 * d * math.log(2.0 * math.pi)

    @property
    def maxLogPdf(self) -> float:
        self._need()
        return self._max_logpdf

    @property
    def logEvidence(self) -> float:
        """ln of the integral of the unnormalised posterior (not in the reference; the yardstick of the SLQ tests)."""
        self._need()
        return self._log_evidence

    def sampleArray(self, numberOfSamples: int, rngSeed: int = 0) -> np.ndarray:
        self._need()
        z = np.random.default_rng(rngSeed).standard_normal((numberOfSamples, self._mean.size))
        return self._mean[None, :] + np.linalg.solve(self._chol.T, z.T).T        # cov = L^-T L^-1

    def sample(self, numberOfSamples: int, rngSeed: int = 0) -> List[Dict[str, np.ndarray]]:
        return [self.posterior.unflatten(x) for x in self.sampleArray(numberOfSamples, rngSeed)]
