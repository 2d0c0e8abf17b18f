"""CPU restatement of the Thylacine posterior-evaluation hot path.  TEST INFRASTRUCTURE ONLY.

This file is the parity oracle for the CUDA path in ``thylacine_b200``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
it; the product path never does (it fails loudly when the CUDA library is missing).

The reference (gvonness/thylacine, Scala 2.13) cannot be compiled or run here (no JVM, no sbt, four
un-vendored dependencies), so this restates its arithmetic in numpy, calling the same LAPACK
routines Breeze 2.1.0 would (``numpy.linalg.solve`` = dgesv = Breeze ``\\``; ``numpy.linalg.inv`` =
dgetrf+dgetri = Breeze ``inv``; ``numpy.linalg.cholesky`` = dpotrf).  Third-party arithmetic that is
NOT in /root/reference and is restated from its published algorithm:

* ``org.scalanlp:breeze_2.13:2.1.0``  MultivariateGaussian.logPdf:
  ``-(Sigma \\ (x-mu)) . (x-mu) / 2 - (n/2 * log(2 pi) + sum(log(diag(cholesky(Sigma)))))``
* ``org.apache.commons:commons-math3:3.6.1``  Gamma.gamma  (here: math.gamma)
* ``ch.obermuhlner:big-math:2.3.2``  BigDecimal exp/log at 34 digits (here: mpmath at 34 digits)

Pinning: the three gradient vectors the reference's own tests hold (GaussianPriorSpec.scala:57,
GaussianLinearLikelihoodSpec.scala:36-38,45-47, GaussianLikelihoodSpec.scala:37-39,46-48) and the
(CI/2)^2 convention (GaussianPriorSpec.scala:34,46) are asserted bit-for-bit in
``tests/test_oracle_golden.py``.  No reference test pins any logPdf VALUE, the Cauchy distribution,
or any sampler, so for those rows parity is "unpinned" beyond the formula restated here (it is
cross-checked against a 50-digit mpmath evaluation and analytic Gaussian results instead).

All ``path:line`` citations are relative to /root/reference/src/main/scala/thylacine/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------------
# RecordedData  (model/core/RecordedData.scala:48-72)
# --------------------------------------------------------------------------------------------
def ci_to_variance(ci: Sequence[float]) -> np.ndarray:
    """sigma_i^2 = Math.pow(ci_i / 2, 2); asserts ci > 0 (RecordedData.scala:57,63)."""
    ci = np.asarray(ci, dtype=np.float64)
    assert np.all(ci > 0), "confidence intervals must be > 0 (RecordedData.scala:57)"
    return np.array([math.pow(c / 2.0, 2) for c in ci], dtype=np.float64)


# --------------------------------------------------------------------------------------------
# Distributions  (model/distributions/*.scala)
# --------------------------------------------------------------------------------------------
class GaussianDistribution:
    """GaussianDistribution.scala:52-68 -- literal: dense covariance, LU solve per logPdf."""

    def __init__(self, mean, covariance):
        self.mean = np.asarray(mean, dtype=np.float64)
        cov = np.asarray(covariance, dtype=np.float64)
        if cov.ndim == 1:  # diagonal given as a vector -> densify like MatrixContainer.rawMatrix
            cov = np.diag(cov)
        assert cov.shape == (self.mean.size, self.mean.size)
        self.covariance = cov
        self.domain_dimension = self.mean.size
        self._root = None
        self._inv = None

    @classmethod
    def from_ci(cls, values, ci):
        return cls(values, ci_to_variance(ci))

    def _log_normalizer(self) -> float:
        # Breeze MultivariateGaussian.logNormalizer: n/2 log(2 pi) + sum(log(diag(cholesky(cov))))
        if self._root is None:
            self._root = np.linalg.cholesky(self.covariance)
        return self.mean.size / 2.0 * LOG_2PI + float(np.sum(np.log(np.diag(self._root))))

    def log_pdf(self, x) -> float:
        centered = np.asarray(x, dtype=np.float64) - self.mean
        slv = np.linalg.solve(self.covariance, centered)  # Breeze `covariance \ centered`
        return -float(slv @ centered) / 2.0 - self._log_normalizer()

    def log_pdf_gradient(self, x) -> np.ndarray:
        # rawInverseCovariance * (mean - input)  (GaussianDistribution.scala:55-56,63-68)
        if self._inv is None:
            self._inv = np.linalg.inv(self.covariance)
        return self._inv @ (self.mean - np.asarray(x, dtype=np.float64))


class CauchyDistribution:
    """CauchyDistribution.scala:51-79 (multivariate Student-t, nu = 1).

    ``reference_faithful=True`` reproduces the reference literally, including
    Q3 (gradient returned with the wrong sign, :70-79) and Q4 (Gamma.gamma / det in the
    normaliser, :51-53, which overflow/underflow for large n).  ``False`` gives the correct sign
    and an lgamma / log-det normaliser (identical value wherever the reference is finite).
    """

    def __init__(self, mean, covariance, reference_faithful: bool = True):
        self.mean = np.asarray(mean, dtype=np.float64)
        cov = np.asarray(covariance, dtype=np.float64)
        if cov.ndim == 1:
            cov = np.diag(cov)
        self.covariance = cov
        self.domain_dimension = self.mean.size
        self.reference_faithful = reference_faithful
        self._inv = np.linalg.inv(cov)
        n = self.domain_dimension
        if reference_faithful:
            with np.errstate(all="ignore"):
                try:
                    g = math.gamma((1 + n) / 2.0)
                except OverflowError:
                    g = math.inf
                det = float(np.linalg.det(cov))
                self.log_multiplier = (
                    (math.log(g) if g > 0 and math.isfinite(g) else math.inf)
                    - math.log(math.gamma(0.5))
                    - n / 2.0 * math.log(math.pi)
                    - (math.log(det) if det > 0 else -math.inf) / 2.0
                )
        else:
            sign, logdet = np.linalg.slogdet(cov)
            self.log_multiplier = (
                math.lgamma((1 + n) / 2.0) - math.lgamma(0.5) - n / 2.0 * math.log(math.pi) - logdet / 2.0
            )

    @classmethod
    def from_ci(cls, values, ci, reference_faithful: bool = True):
        return cls(values, ci_to_variance(ci), reference_faithful)

    def _q(self, x):
        diff = np.asarray(x, dtype=np.float64) - self.mean
        return diff, float(diff @ (self._inv @ diff))

    def log_pdf(self, x) -> float:
        _, q = self._q(x)
        return self.log_multiplier - (1.0 + self.domain_dimension) / 2.0 * math.log(1 + q)

    def log_pdf_gradient(self, x) -> np.ndarray:
        diff, q = self._q(x)
        mult = (1 + self.domain_dimension) / (1 + q)
        v = mult * (self._inv @ diff)
        return v if self.reference_faithful else -v


class UniformDistribution:
    """UniformDistribution.scala:50-81 -- half-open box [lo, hi); -sum log(hi-lo) or -inf; zero gradient."""

    def __init__(self, upper, lower):
        self.upper = np.asarray(upper, dtype=np.float64)
        self.lower = np.asarray(lower, dtype=np.float64)
        assert self.upper.shape == self.lower.shape
        assert not np.any(self.upper <= self.lower), "upper bound must exceed lower bound (:36-39)"
        self.domain_dimension = self.upper.size
        self.neg_log_volume = -float(sum(math.log(u - l) for l, u in zip(self.lower, self.upper)))

    def inside_bounds(self, x) -> bool:
        x = np.asarray(x, dtype=np.float64)
        return bool(np.all((x >= self.lower) & (x < self.upper)))

    def log_pdf(self, x) -> float:
        return self.neg_log_volume if self.inside_bounds(x) else -math.inf

    def log_pdf_gradient(self, x) -> np.ndarray:
        return np.zeros(self.domain_dimension)


# --------------------------------------------------------------------------------------------
# Forward models  (model/components/forwardmodel/*.scala)
# --------------------------------------------------------------------------------------------
class LinearForwardModel:
    """LinearForwardModel.scala:76-114.  ``blocks``: label -> (n x d_label) row-major matrix.

    Blocks are column-merged in ascending label order (:87-101); eval = A . flatten(theta) (+offset)
    (:103-104,111-114); the Jacobian is the constant per-label block (:65-71).
    """

    def __init__(self, blocks: Dict[str, np.ndarray], offset=None):
        self.blocks = {k: np.atleast_2d(np.asarray(v, dtype=np.float64)) for k, v in blocks.items()}
        rows = {b.shape[0] for b in self.blocks.values()}
        assert len(rows) == 1, "all blocks must share the range dimension (:47)"
        self.range_dimension = rows.pop()
        self.labels = sorted(self.blocks)
        self.matrix = np.concatenate([self.blocks[l] for l in self.labels], axis=1)
        self.domain_dimension = self.matrix.shape[1]
        self.offset = None if offset is None else np.asarray(offset, dtype=np.float64)
        if self.offset is not None:
            assert self.offset.size == self.range_dimension

    @classmethod
    def identity_of(cls, label: str, dimension: int):
        return cls({label: np.eye(dimension)})

    def eval_at(self, params: Dict[str, np.ndarray]) -> np.ndarray:
        theta = np.concatenate([np.asarray(params[l], dtype=np.float64) for l in self.labels])
        f = self.matrix @ theta
        return f if self.offset is None else self.offset + f

    def jacobian_at(self, params) -> Dict[str, np.ndarray]:
        return self.blocks


class NonLinearForwardModel:
    """NonLinearForwardModel.scala:44-85 with the finite-difference Jacobian of
    core/computation/FiniteDifferenceJacobian.scala:30-55: forward difference, d_total+1 evaluations,
    every component of every label present in the input is nudged (Q9), difference multiplied by
    the reciprocal ``1 / differential`` (:45)."""

    def __init__(self, evaluation: Callable[[Dict[str, np.ndarray]], np.ndarray], differential: Optional[float],
                 domain_dimensions: Dict[str, int], range_dimension: int,
                 jacobian: Optional[Callable[[Dict[str, np.ndarray]], Dict[str, np.ndarray]]] = None):
        self.evaluation = evaluation
        self.differential = differential
        self.domain_dimensions = dict(domain_dimensions)
        self.domain_dimension = sum(domain_dimensions.values())
        self.range_dimension = range_dimension
        self.jacobian = jacobian

    def eval_at(self, params):
        return np.asarray(self.evaluation(params), dtype=np.float64)

    def jacobian_at(self, params) -> Dict[str, np.ndarray]:
        if self.jacobian is not None:
            return {k: np.asarray(v, dtype=np.float64) for k, v in self.jacobian(params).items()}
        h = self.differential
        inv_h = 1 / h
        current = self.eval_at(params)
        out = {}
        for label, vec in params.items():
            vec = np.asarray(vec, dtype=np.float64)
            cols = []
            for j in range(vec.size):
                nudged = dict(params)
                v2 = vec.copy()
                v2[j] = v2[j] + h  # VectorContainer.rawNudgeComponent (values/VectorContainer.scala:111-113)
                nudged[label] = v2
                cols.append((self.eval_at(nudged) - current) * inv_h)
            out[label] = np.stack(cols, axis=1)
        return out


# --------------------------------------------------------------------------------------------
# Priors, likelihoods, posterior  (model/components/{prior,likelihood,posterior}/*.scala)
# --------------------------------------------------------------------------------------------
@dataclass
class Prior:
    """Prior.scala:52-66: evaluate the wrapped distribution on this prior's own labelled block."""
    label: str
    distribution: object

    @property
    def dimension(self) -> int:
        return self.distribution.domain_dimension

    def log_pdf(self, params) -> float:
        return self.distribution.log_pdf(params[self.label])

    def log_pdf_gradient(self, params) -> Dict[str, np.ndarray]:
        return {self.label: self.distribution.log_pdf_gradient(params[self.label])}


def gaussian_prior_from_ci(label, values, ci) -> Prior:                      # GaussianPrior.scala:62-75
    return Prior(label, GaussianDistribution.from_ci(values, ci))


def gaussian_prior_from_cov(label, values, cov) -> Prior:                    # GaussianPrior.scala:78-93
    return Prior(label, GaussianDistribution(values, np.asarray(cov, dtype=np.float64)))


def uniform_prior_from_bounds(label, max_bounds, min_bounds) -> Prior:       # UniformPrior.scala:68-77
    return Prior(label, UniformDistribution(max_bounds, min_bounds))


def cauchy_prior(label, values, ci, reference_faithful=True) -> Prior:       # CauchyPrior.scala:50-63
    return Prior(label, CauchyDistribution.from_ci(values, ci, reference_faithful))


@dataclass
class Likelihood:
    """Likelihood.scala:43-65: logPdf = D.logPdf(f(theta)); gradient per label = (D.grad(f)^T . J_label)^T."""
    forward_model: object
    distribution: object

    def log_pdf(self, params) -> float:
        return self.distribution.log_pdf(self.forward_model.eval_at(params))

    def log_pdf_gradient(self, params) -> Dict[str, np.ndarray]:
        mapped = self.forward_model.eval_at(params)
        jac = self.forward_model.jacobian_at(params)
        meas_grad = self.distribution.log_pdf_gradient(mapped)
        return {label: (meas_grad @ j) for label, j in jac.items()}


def gaussian_linear_likelihood(coefficients, measurements, uncertainties, prior_label) -> Likelihood:
    """GaussianLinearLikelihood.of (likelihood/GaussianLinearLikelihood.scala:62-83)."""
    fm = LinearForwardModel({prior_label: coefficients})
    assert fm.range_dimension == len(measurements)                            # :43-45
    return Likelihood(fm, GaussianDistribution.from_ci(measurements, uncertainties))


def gaussian_likelihood(forward_model, measurements, uncertainties) -> Likelihood:    # GaussianLikelihood.scala:55-67
    assert forward_model.range_dimension == len(measurements)
    return Likelihood(forward_model, GaussianDistribution.from_ci(measurements, uncertainties))


def cauchy_likelihood(forward_model, measurements, uncertainties, reference_faithful=True) -> Likelihood:
    """CauchyLikelihood.apply (likelihood/CauchyLikelihood.scala:59-72)."""
    assert forward_model.range_dimension == len(measurements)
    return Likelihood(forward_model, CauchyDistribution.from_ci(measurements, uncertainties, reference_faithful))


class Posterior:
    """Posterior.scala:29-74.  Flatten order = prior labels sorted ascending (:40-48)."""

    def __init__(self, priors: Sequence[Prior], likelihoods: Sequence[Likelihood]):
        labels = [p.label for p in priors]
        assert len(set(labels)) == len(labels), "duplicate prior labels (UnnormalisedPosterior.scala:35-47)"
        self.priors = list(priors)
        self.likelihoods = list(likelihoods)
        self.ordered = sorted(((p.label, p.dimension) for p in priors), key=lambda t: t[0])
        self.domain_dimension = sum(d for _, d in self.ordered)
        self.offsets = {}
        off = 0
        for label, d in self.ordered:
            self.offsets[label] = (off, d)
            off += d

    # ModelParameterContext.scala:42-80
    def unflatten(self, x) -> Dict[str, np.ndarray]:
        x = np.asarray(x, dtype=np.float64)
        return {l: x[o:o + d] for l, (o, d) in self.offsets.items()}

    def flatten(self, params: Dict[str, np.ndarray]) -> np.ndarray:
        out = np.zeros(self.domain_dimension)
        for l, (o, d) in self.offsets.items():
            if l in params:
                out[o:o + d] = params[l]
        return out

    def log_pdf(self, x) -> float:                                            # Posterior.scala:67-73
        params = self.unflatten(x)
        vals = [p.log_pdf(params) for p in self.priors] + [l.log_pdf(params) for l in self.likelihoods]
        total = 0.0
        for v in vals:  # left-to-right .sum
            total += v
        return total

    def log_pdf_gradient(self, x) -> np.ndarray:                              # Posterior.scala:50-62
        params = self.unflatten(x)
        out = np.zeros(self.domain_dimension)
        for term in list(self.priors) + list(self.likelihoods):
            for label, g in term.log_pdf_gradient(params).items():
                if label in self.offsets:
                    o, d = self.offsets[label]
                    out[o:o + d] += g
        return out

    def sample_priors(self, rng: np.random.Generator) -> np.ndarray:          # Posterior.scala:64-65
        params = {}
        for p in self.priors:
            dist = p.distribution
            if isinstance(dist, GaussianDistribution):                        # mu + chol(Sigma) z
                root = np.linalg.cholesky(dist.covariance)
                params[p.label] = dist.mean + root @ rng.standard_normal(dist.mean.size)
            elif isinstance(dist, UniformDistribution):                       # UniformDistribution.scala:88-93
                params[p.label] = dist.lower + rng.random(dist.lower.size) * (dist.upper - dist.lower)
            elif isinstance(dist, CauchyDistribution):                        # CauchyDistribution.scala:84-87
                chi = rng.chisquare(1)
                root = np.linalg.cholesky(dist.covariance / chi)
                params[p.label] = dist.mean + root @ rng.standard_normal(dist.mean.size)
            else:
                raise TypeError(dist)
        return self.flatten(params)


# --------------------------------------------------------------------------------------------
# Structure-aware batched evaluation (diagonal covariances): the SAME arithmetic as above once the
# dense n x n solve is replaced by a division by the diagonal.  Used for parity at sizes the literal
# dense form cannot reach (n = 10^4 needs an 800 MB covariance and 0.7 TFLOP per logPdf) and as the
# "port" CPU baseline.  Checked against the literal classes at small sizes in tests/.
# --------------------------------------------------------------------------------------------
def linear_gaussian_batch(A, y, var, X, offset=None, prior_mean=None, prior_var=None, want_grad=True):
    """X: (B, d) row-major chain-major.  Returns (logp[B], grad[B, d] or None).

    logL = -1/2 sum (y - A x)^2 / var - (n/2 log 2pi + 1/2 sum log var); grad = A^T ((y - A x)/var).
    Optional diagonal Gaussian prior over the whole flat vector.
    """
    A = np.asarray(A, dtype=np.float64)
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    y = np.asarray(y, dtype=np.float64)
    var = np.asarray(var, dtype=np.float64)
    n = y.size
    F = X @ A.T
    if offset is not None:
        F = F + np.asarray(offset)[None, :]
    R = y[None, :] - F
    W = R / var[None, :]
    logp = -0.5 * np.einsum("bn,bn->b", R, W) - (n / 2.0 * LOG_2PI + 0.5 * float(np.sum(np.log(var))))
    grad = (W @ A) if want_grad else None
    if prior_mean is not None:
        pm = np.asarray(prior_mean, dtype=np.float64)
        pv = np.asarray(prior_var, dtype=np.float64)
        D = pm[None, :] - X
        logp = logp + (-0.5 * np.einsum("bd,bd->b", D, D / pv[None, :])
                       - (pm.size / 2.0 * LOG_2PI + 0.5 * float(np.sum(np.log(pv)))))
        if want_grad:
            grad = grad + D / pv[None, :]
    return logp, grad


def mp_linear_gaussian(A, y, var, x, prior_mean=None, prior_var=None, dps=50):
    """50-digit mpmath arbiter for one point (logp, grad) -- used when fp64 orderings disagree."""
    import mpmath as mp
    with mp.workdps(dps):
        A_ = mp.matrix(np.asarray(A).tolist())
        x_ = mp.matrix([mp.mpf(float(v)) for v in x])
        f = A_ * x_
        n = len(y)
        r = [mp.mpf(float(y[i])) - f[i] for i in range(n)]
        w = [r[i] / mp.mpf(float(var[i])) for i in range(n)]
        logp = -sum(r[i] * w[i] for i in range(n)) / 2 - (mp.mpf(n) / 2 * mp.log(2 * mp.pi)
                                                              + sum(mp.log(mp.mpf(float(v))) for v in var) / 2)
        d = len(x)
        grad = [sum(A_[i, j] * w[i] for i in range(n)) for j in range(d)]
        if prior_mean is not None:
            for j in range(d):
                dj = mp.mpf(float(prior_mean[j])) - x_[j]
                pv = mp.mpf(float(prior_var[j]))
                logp -= dj * dj / pv / 2
                grad[j] += dj / pv
            logp -= mp.mpf(d) / 2 * mp.log(2 * mp.pi) + sum(mp.log(mp.mpf(float(v))) for v in prior_var) / 2
        return float(logp), np.array([float(g) for g in grad])


# --------------------------------------------------------------------------------------------
# Analytic Gaussian posterior + evidence  (GaussianAnalyticPosterior.scala:141-169; evidence is NOT
# in the reference -- standard conjugate result, required by north_star as the statistical oracle)
# --------------------------------------------------------------------------------------------
def analytic_gaussian_posterior(prior_mean, prior_cov, A, y, var, offset=None):
    """Lambda = Sp^-1 + A^T Sd^-1 A;  mu = Lambda^-1 (Sp^-1 mu_p + A^T Sd^-1 (y - offset))."""
    prior_mean = np.asarray(prior_mean, dtype=np.float64)
    prior_cov = np.asarray(prior_cov, dtype=np.float64)
    if prior_cov.ndim == 1:
        prior_cov = np.diag(prior_cov)
    A = np.asarray(A, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if offset is not None:
        y = y - np.asarray(offset)
    var = np.asarray(var, dtype=np.float64)
    sp_inv = np.linalg.inv(prior_cov)
    lam = sp_inv + A.T @ (A / var[:, None])
    cov = np.linalg.inv(lam)
    mean = np.linalg.solve(lam, sp_inv @ prior_mean + A.T @ (y / var))
    return mean, cov


def analytic_gaussian_log_evidence(prior_mean, prior_cov, A, y, var) -> float:
    """ln Z = ln N(y; A mu_p, Sd + A Sp A^T)."""
    prior_cov = np.asarray(prior_cov, dtype=np.float64)
    if prior_cov.ndim == 1:
        prior_cov = np.diag(prior_cov)
    A = np.asarray(A, dtype=np.float64)
    S = np.diag(np.asarray(var, dtype=np.float64)) + A @ prior_cov @ A.T
    r = np.asarray(y, dtype=np.float64) - A @ np.asarray(prior_mean, dtype=np.float64)
    sign, logdet = np.linalg.slogdet(S)
    return float(-0.5 * r @ np.linalg.solve(S, r) - 0.5 * logdet - r.size / 2.0 * LOG_2PI)


def analytic_uniform_box_log_evidence(lower, upper, A, y, var) -> float:
    """Uniform box prior of volume V, box >> posterior: ln Z ~ -ln V + ln int N(y; A x, Sd) dx
    = -ln V - 1/2 r^T Sd^-1 r|_xhat - 1/2 sum ln(2 pi var) + d/2 ln(2 pi) - 1/2 ln det(A^T Sd^-1 A)."""
    A = np.asarray(A, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    var = np.asarray(var, dtype=np.float64)
    lam = A.T @ (A / var[:, None])
    xhat = np.linalg.solve(lam, A.T @ (y / var))
    r = y - A @ xhat
    sign, logdet = np.linalg.slogdet(lam)
    lnV = float(np.sum(np.log(np.asarray(upper, dtype=np.float64) - np.asarray(lower, dtype=np.float64))))
    d = A.shape[1]
    return float(-lnV - 0.5 * np.sum(r * r / var) - 0.5 * np.sum(np.log(2 * np.pi * var))
                 + d / 2.0 * LOG_2PI - 0.5 * logdet)


# --------------------------------------------------------------------------------------------
# HMC  (model/sampling/hmcmc/HmcmcEngine.scala:55-214)
# --------------------------------------------------------------------------------------------
@dataclass
class HmcmcConfig:                                                            # config/HmcmcConfig.scala:20-25
    stepsBetweenSamples: int
    stepsInDynamicsSimulation: int
    warmupStepCount: int
    dynamicsSimulationStepSize: float


class HmcmcEngine:
    """Single-chain HMC exactly as HmcmcEngine.scala structures it.

    * leapfrog: p -= eps/2 g; x += eps p; g = -grad logp(x); p -= eps/2 g, L times (:55-80);
    * H = p.p/2 + E (:82-83); accept iff dH < 0 or U < exp(-dH) (:128);
    * ``iterationCount <= maxIterations`` => maxIterations+1 trajectories per block (:168, Q5);
    * samples form a Set: a fully rejected block adds nothing and sampling continues (:149-152);
    * burn-in (warmupStepCount+1 trajectories) re-runs on every ``sample`` call (:180-197,205-207).

    The reference RNG is unseeded; here the caller supplies the standard normals / uniforms through
    ``rng`` (np.random.Generator), or explicit hooks ``momentum_fn(traj_index, d)`` and
    ``uniform_fn(traj_index)`` so a device run with a counter-based RNG can be replayed bit-for-bit.
    """

    def __init__(self, posterior, config: HmcmcConfig, rng=None, seed_point=None,
                 momentum_fn=None, uniform_fn=None, logp_fn=None, grad_fn=None):
        self.posterior = posterior
        self.cfg = config
        self.rng = rng if rng is not None else np.random.default_rng(0)
        self.seed_point = seed_point
        self.momentum_fn = momentum_fn
        self.uniform_fn = uniform_fn
        self.logp = logp_fn or posterior.log_pdf
        self.grad = grad_fn or posterior.log_pdf_gradient
        self.jump_attempts = 0
        self.jump_acceptances = 0
        self.traj = 0
        self.last_dH = None

    def _leapfrog(self, x, p, g):
        eps = self.cfg.dynamicsSimulationStepSize
        for _ in range(self.cfg.stepsInDynamicsSimulation):
            p = p + g * (-eps / 2)
            x = x + p * eps
            g = -self.grad(x)
            p = p + g * (-eps / 2)
        return x, p

    def _block(self, x, E, g, max_iterations):
        """One pass of iterationCount = 0..maxIterations (inclusive) of runDynamicSimulationFrom."""
        d = x.size
        for _ in range(max_iterations + 1):
            if E is None:
                E = -self.logp(x)
            if g is None:
                g = -self.grad(x)
            p = self.momentum_fn(self.traj, d) if self.momentum_fn else self.rng.standard_normal(d)
            H = float(p @ p) / 2.0 + E
            x_new, p_new = self._leapfrog(x, p, g)
            E_new = -self.logp(x_new)
            H_new = float(p_new @ p_new) / 2.0 + E_new
            dH = H_new - H
            self.last_dH = dH
            u = self.uniform_fn(self.traj) if self.uniform_fn else self.rng.random()
            self.traj += 1
            self.jump_attempts += 1
            if dH < 0 or u < math.exp(-dH):
                x, E, g = x_new, E_new, -self.grad(x_new)
                self.jump_acceptances += 1
        return x, E, g

    def sample(self, n: int) -> List[np.ndarray]:
        x = self.seed_point if self.seed_point is not None else self.posterior.sample_priors(self.rng)
        x = np.asarray(x, dtype=np.float64)
        x, E, g = self._block(x, None, None, self.cfg.warmupStepCount)      # burn-in
        samples: List[np.ndarray] = []
        seen = set()
        guard = 0
        while len(samples) < n:
            x, E, g = self._block(x, E, g, self.cfg.stepsBetweenSamples)
            key = x.tobytes()
            if key not in seen:                                               # Set semantics
                seen.add(key)
                samples.append(x.copy())
            guard += 1
            if guard > 1000 * n + 1000:
                raise RuntimeError("HMC never accepts; check step size")
        return samples


# --------------------------------------------------------------------------------------------
# Leapfrog-MCMC  (model/sampling/leapfrog/LeapfrogMcmcEngine.scala:88-162,224-261; util/MathOps.scala:76-91)
# --------------------------------------------------------------------------------------------
def vector_cdf_staircase(values) -> List[Tuple[float, float]]:
    """MathOps.vectorCdfStaircase (:76-91): sequential running sum, each divided by the total."""
    acc = [0.0]
    for v in values:
        acc.append(acc[-1] + float(v))
    total = acc[-1]
    norm = [a / total for a in acc]
    return list(zip(norm[:-1], norm[1:]))


def staircase_index(staircase, u: float) -> int:
    """indexWhere(lo <= u && hi > u)  (LeapfrogMcmcEngine.scala:124-125); -1 if none."""
    for k, (lo, hi) in enumerate(staircase):
        if lo <= u and hi > u:
            return k
    return -1


def euclidean(a, b) -> float:
    return float(math.sqrt(np.sum((np.asarray(a) - np.asarray(b)) ** 2)))


@dataclass
class LeapfrogMcmcConfig:                                                     # config/LeapfrogMcmcConfig.scala:20-24
    stepsBetweenSamples: int
    warmupStepCount: int
    samplePoolSize: int


class LeapfrogMcmcEngine:
    """The sequential ensemble sampler exactly as the reference runs it: P x P distance table, one
    proposal at a time (generateProposal, :116-162)."""

    def __init__(self, posterior, config: LeapfrogMcmcConfig, distance=euclidean, rng=None, seed_point=None):
        self.posterior = posterior
        self.cfg = config
        self.distance = distance
        self.rng = rng if rng is not None else np.random.default_rng(0)
        P = config.samplePoolSize
        pool = [posterior.sample_priors(self.rng) for _ in range(P)]          # populateSamples :237-253
        if seed_point is not None:
            pool[0] = np.asarray(seed_point, dtype=np.float64)
        self.pool = pool
        self.dist = [[distance(a, b) for b in pool] for a in pool]           # initialisePool :224-235
        self.logp = [posterior.log_pdf(s) for s in pool]
        self.tally = [0] * P
        self.proposals = 0
        self.acceptances = 0
        while not all(t >= config.warmupStepCount for t in self.tally):       # burnInRecursion :172-182
            self.generate_proposal()
        self.tally = [0] * P

    def generate_proposal(self) -> Tuple[bool, int]:
        P = self.cfg.samplePoolSize
        i = int(self.rng.integers(P))
        row = self.dist[i]
        stair = vector_cdf_staircase(row)
        u = float(self.rng.random())
        j = staircase_index(stair, u)
        q_fwd = stair[j][1] - stair[j][0]
        new_point = 2.0 * self.pool[j] - self.pool[i]                         # getLeapfrogVector :88-89
        new_d = [self.distance(new_point, self.pool[k]) if k != i else 0.0 for k in range(P)]   # :91-104
        stair2 = vector_cdf_staircase(new_d)
        q_rev = stair2[j][1] - stair2[j][0]
        new_logp = self.posterior.log_pdf(new_point)
        with np.errstate(over="ignore"):
            a = math.exp(min(new_logp - self.logp[i], 700.0)) * q_rev / q_fwd
        self.proposals += 1
        if a >= 1 or float(self.rng.random()) < a:
            for k in range(P):
                self.dist[k][i] = self.distance(self.pool[k], new_point)     # :141-146
            self.dist[i] = new_d
            self.logp[i] = new_logp
            self.pool[i] = new_point
            self.acceptances += 1
        self.tally[i] += 1
        return self.tally[i] >= self.cfg.stepsBetweenSamples, i

    def sample(self, n: int) -> List[np.ndarray]:
        out = []
        for _ in range(n):
            while True:
                ready, idx = self.generate_proposal()
                if ready:
                    break
            out.append(self.pool[idx].copy())
            self.tally[idx] = 0
        return out


# --------------------------------------------------------------------------------------------
# SLQ quadrature pieces  (model/integration/slq/QuadratureAbscissa.scala:33-73,
# QuadratureIntegrator.scala:29-83).  BigDecimal(34 digits) -> mpmath at 34 digits.
# --------------------------------------------------------------------------------------------
class QuadratureAbscissa:
    """Prior-mass simulation: the pool starts as {1.0} + (P-1) uniforms (:64-73); each step moves
    the pool max to the abscissa list and replaces it by U(0, max) (:33-37).  ``abscissa`` here is
    kept in chronological order (largest X first); the reference prepends, so its vector is the
    reverse.  Trapezoid weights (:39-49): interior 1/2 (X_{i-1} - X_{i+1}), ends half the adjacent gap."""

    def __init__(self, pool_size: int, rng: np.random.Generator):
        self.rng = rng
        self.samples = [1.0] + list(rng.random(pool_size - 1))
        self.abscissa: List[float] = []

    def extend_by_one(self):
        m = max(self.samples)
        self.samples.remove(m)
        self.samples.append(float(self.rng.random()) * m)
        self.abscissa.append(m)

    def trapezoid_weights(self) -> List[float]:
        xs = self.abscissa
        k = len(xs)
        assert k >= 2, "the reference's take(2).reduce needs two abscissas"
        w = [0.5 * (xs[0] - xs[1])]
        for i in range(1, k - 1):
            w.append(0.5 * (xs[i - 1] - xs[i + 1]))
        w.append(0.5 * (xs[k - 2] - xs[k - 1]))
        return w


def quadrature_integrator_stats(logpdfs: Sequence[float], weights: Sequence[float], dps: int = 34):
    """QuadratureIntegrator.evidenceStats / negativeEntropyStats for ONE abscissa (:29-62), literal:
    shift s = (max + min)/2 is subtracted and never restored (Q6).  Returns (Z_a, H_a) as floats."""
    import mpmath as mp
    with mp.workdps(dps):
        s = (max(logpdfs) + min(logpdfs)) / 2.0
        z = mp.mpf(0)
        e = mp.mpf(0)
        for l, w in zip(logpdfs, weights):
            t = mp.mpf(float(l) - s)
            ex = mp.exp(t)
            z += ex * mp.mpf(float(w))
            e += ex * t * mp.mpf(float(w))
        return float(z), float(e / z - mp.log(z))


def log_sum_exp_evidence(logps: Sequence[float], weights: Sequence[float]) -> float:
    """ln sum_i w_i exp(l_i) in log space (the build's own calibrated estimator)."""
    l = np.asarray(logps, dtype=np.float64)
    w = np.asarray(weights, dtype=np.float64)
    keep = w > 0
    a = l[keep] + np.log(w[keep])
    m = np.max(a)
    return float(m + math.log(np.sum(np.exp(a - m))))


# --------------------------------------------------------------------------------------------
# CartesianSurface  (visualisation/CartesianSurface.scala:30-200, SimplexChain.scala:19-45,
# Simplex.scala:21-53, GraphPoint.scala:20-47, util/MathOps.scala:26-55) -- literal restatement
# --------------------------------------------------------------------------------------------
def simplex_interpolation_points(p0, p1, start: float, ds: float):
    """Simplex.getInterpolationPoints (Simplex.scala:36-47): points at start, start+ds, ... < length, PREPENDED
    (so the list comes back reversed), and the remainder start - length."""
    length = math.sqrt((p0[0] - p1[0]) ** 2 + (p0[1] - p1[1]) ** 2)        # GraphPoint.distanceTo
    prev: List[Tuple[float, float]] = []
    while not (start >= length):
        t = start / length
        if t <= 0:
            pt = p0
        elif t >= 1:
            pt = p1
        else:
            pt = (p0[0] + t * (p1[0] - p0[0]), p0[1] + t * (p1[1] - p0[1]))
        prev = [pt] + prev
        start = start + ds
    return prev, start - length


def simplex_chain_points(abscissa, values, ds: float) -> List[Tuple[float, float]]:
    """SimplexChain(points).linearInterpolationUsing(ds).getPoints (SimplexChain.scala:19-45): the interpolated points
    are wrapped in a new chain of simplexes, whose getPoints is every simplex START -- the last point is dropped, and a
    chain of fewer than two points is empty."""
    pts = list(zip([float(a) for a in abscissa], [float(v) for v in values]))
    if len(pts) < 2:
        return []
    carry, out = 0.0, []
    for p0, p1 in zip(pts[:-1], pts[1:]):
        seg, carry = simplex_interpolation_points(p0, p1, carry, ds)
        out = out + seg
    if len(out) < 2:
        return []
    return out[:-1]


def trapezoidal_quadrature(abscissa, values) -> float:
    """MathOps.trapezoidalQuadrature (MathOps.scala:26-55)."""
    pairs = list(zip([float(a) for a in abscissa], [float(v) for v in values]))
    if not (len(pairs) > 1 and len({p[0] for p in pairs}) == len(pairs)):
        raise RuntimeError("Malformed abscissa for trapezoidal quadrature")
    pairs.sort(key=lambda p: p[0])
    return sum(0.5 * (a[1] + b[1]) * (b[0] - a[0]) for a, b in zip(pairs[:-1], pairs[1:]))


class CartesianSurface:
    """CartesianSurface.of / addSamples / getResults, one grid point and one curve point at a time."""

    def __init__(self, x_abscissa, y_abscissa):
        self.x = [float(v) for v in x_abscissa]
        self.y = [float(v) for v in y_abscissa]
        self.x_scale = max(self.x) - min(self.x)                             # :40-41
        self.y_scale = max(self.y) - min(self.y)
        self.values = {(gx, gy): 0.0 for gx in self.x for gy in self.y}      # zeroValuesSpec :85-88

    def add_samples(self, abscissa, samples, ds: float, kernel_variance: float) -> None:
        for curve in samples:                                                # :139-143, one addValues per sample
            pts = simplex_chain_points(abscissa, curve, ds)
            for g in self.values:                                            # concatenateMapping :110-127
                acc = self.values[g]
                for p in pts:
                    sd = ((p[0] - g[0]) / self.x_scale) ** 2 + ((p[1] - g[1]) / self.y_scale) ** 2
                    acc = acc + math.exp(-0.5 * sd / kernel_variance)
                self.values[g] = acc

    def get_results(self) -> Dict[Tuple[float, float], float]:
        out = {}
        for gx in self.x:                                                    # normaliseColumns :53-83
            col = [self.values[(gx, gy)] for gy in self.y]
            integration = trapezoidal_quadrature(self.y, col)
            for gy, v in zip(self.y, col):
                out[(gx, gy)] = v / integration if integration > 0 else v
        return out

    def results_array(self, normalise: bool = True) -> np.ndarray:
        src = self.get_results() if normalise else self.values
        return np.array([[src[(gx, gy)] for gy in self.y] for gx in self.x])
