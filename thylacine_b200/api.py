"""Host-side mirror of the reference's Scala surface for the posterior-evaluation hot path.

Same names, argument meaning and error behaviour as the reference (camelCase kept on purpose so the
parity tests read like the reference's specs); every numeric call goes through the C ABI in
``include/thylacine_b200.h``.  Paths below are relative to
/root/reference/src/main/scala/thylacine/model/components/.

    GaussianPrior.fromConfidenceIntervals / fromCovarianceMatrix   prior/GaussianPrior.scala:62-93
    UniformPrior.fromBounds                                        prior/UniformPrior.scala:68-77
    CauchyPrior                                                    prior/CauchyPrior.scala:50-63
    LinearForwardModel.of / identityOf                             forwardmodel/LinearForwardModel.scala:76-160
    NonLinearForwardModel.of                                       forwardmodel/NonLinearForwardModel.scala:44-85
    GaussianLinearLikelihood.of                                    likelihood/GaussianLinearLikelihood.scala:62-83
    GaussianLikelihood / CauchyLikelihood / UniformLikelihood      likelihood/*.scala
    UnnormalisedPosterior                                          posterior/UnnormalisedPosterior.scala:28-47
    logPdfAt / logPdfGradientAt                                    posterior/Posterior.scala:50-73

A JVM closure cannot run on a GPU: ``NonLinearForwardModel.of`` takes a registered device family
``link(A . theta + offset)`` instead of ``evaluation: Map => Vector``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Mapping, Optional, Sequence

import numpy as np

from . import _capi
from ._capi import ThylacineError, check, lib

_f64 = np.float64


def _arr(v, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(v, dtype=_f64))
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# ---------------------------------------------------------------------------------------------- priors
@dataclass
class _PriorSpec:
    label: str
    kind: str
    a: np.ndarray
    b: np.ndarray

    @property
    def dimension(self) -> int:
        return int(self.a.size)


class GaussianPrior:
    @staticmethod
    def fromConfidenceIntervals(label: str, values, confidenceIntervals) -> _PriorSpec:
        v, ci = _arr(values), _arr(confidenceIntervals)
        assert v.size == ci.size, "values and confidenceIntervals differ in size (GaussianPrior.scala:67)"
        return _PriorSpec(label, "gaussian_ci", v, ci)

    @staticmethod
    def fromCovarianceMatrix(label: str, values, covarianceMatrix) -> _PriorSpec:
        v = _arr(values)
        cov = _arr(covarianceMatrix)
        assert cov.shape == (v.size, v.size), "covariance must be square and match values (GaussianPrior.scala:85)"
        return _PriorSpec(label, "gaussian_cov", v, cov)


class UniformPrior:
    @staticmethod
    def fromBounds(label: str, maxBounds, minBounds) -> _PriorSpec:
        hi, lo = _arr(maxBounds), _arr(minBounds)
        assert hi.size == lo.size
        return _PriorSpec(label, "uniform", hi, lo)


def CauchyPrior(label: str, values, confidenceIntervals) -> _PriorSpec:
    v, ci = _arr(values), _arr(confidenceIntervals)
    assert v.size == ci.size
    return _PriorSpec(label, "cauchy", v, ci)


# --------------------------------------------------------------------------------------- forward models
LINKS = {"identity": _capi.LINK_IDENTITY, "tanh": _capi.LINK_TANH, "exp": _capi.LINK_EXP, "square": _capi.LINK_SQUARE,
         "softplus": _capi.LINK_SOFTPLUS, "sigmoid": _capi.LINK_SIGMOID}


@dataclass
class _ForwardModel:
    blocks: Dict[str, np.ndarray]        # label -> (n x d_label)
    offset: Optional[np.ndarray]
    link: int
    jacobian: int
    differential: float

    @property
    def rangeDimension(self) -> int:
        return next(iter(self.blocks.values())).shape[0]

    @property
    def domainDimension(self) -> int:
        return sum(b.shape[1] for b in self.blocks.values())


class LinearForwardModel:
    @staticmethod
    def of(label=None, values=None, evalCacheDepth=None, transform: Optional[Mapping[str, Sequence]] = None,
           vectorOffset=None) -> _ForwardModel:
        """``of(label, values, evalCacheDepth)`` or ``of(transform={label: matrix}, vectorOffset=...)``.
        evalCacheDepth is accepted and ignored (a per-point memo is meaningless for a batched device path)."""
        if transform is None:
            transform = {label: values}
        blocks = {k: np.atleast_2d(_arr(v)) for k, v in transform.items()}
        rows = {b.shape[0] for b in blocks.values()}
        assert len(rows) == 1, "all blocks must share the range dimension (LinearForwardModel.scala:47)"
        off = None if vectorOffset is None else _arr(vectorOffset)
        if off is not None:
            assert off.size == rows.copy().pop(), "vectorOffset dimension mismatch (LinearForwardModel.scala:48-50)"
        return _ForwardModel(blocks, off, _capi.LINK_IDENTITY, _capi.JAC_CONSTANT, 0.0)

    @staticmethod
    def identityOf(label: str, dimension: int, evalCacheDepth=None) -> _ForwardModel:
        return LinearForwardModel.of(label, np.eye(dimension))


@dataclass(frozen=True)
class UserLink:
    """A forward-model family the device registry does not hold, as CUDA C source in the variable ``z``
    (``thy_model_register_link``: compiled with NVRTC for sm_100a when the posterior is built).  ``derivative`` is the
    link's derivative (``z`` and ``f = link(z)`` in scope); omitted => pass ``differential`` to ``NonLinearForwardModel.of``
    and the link's own forward difference is used.  Stands in for the reference's ``evaluation`` / ``jacobian`` closures
    (NonLinearForwardModel.scala:44-85), which cannot run on a GPU."""
    evaluation: str
    derivative: Optional[str] = None


class NonLinearForwardModel:
    @staticmethod
    def of(link, coefficients: Mapping[str, Sequence], differential: Optional[float] = None, vectorOffset=None,
           domainDimensions: Optional[Mapping[str, int]] = None, rangeDimension: Optional[int] = None,
           evalCacheDepth=None, jacobianCacheDepth=None) -> _ForwardModel:
        """Device-registered family f(theta) = link(A . theta + offset).  ``differential`` given => the
        finite-difference Jacobian of FiniteDifferenceJacobian.scala:30-55; omitted => analytic Jacobian
        (the reference's second ``of`` overload, NonLinearForwardModel.scala:64-85)."""
        blocks = {k: np.atleast_2d(_arr(v)) for k, v in coefficients.items()}
        rows = {b.shape[0] for b in blocks.values()}
        assert len(rows) == 1
        if domainDimensions is not None:
            assert {k: b.shape[1] for k, b in blocks.items()} == dict(domainDimensions)
        if rangeDimension is not None:
            assert rows.copy().pop() == rangeDimension
        off = None if vectorOffset is None else _arr(vectorOffset)
        code = link if isinstance(link, UserLink) else LINKS[link]   # a UserLink is resolved to an id by the posterior
        if differential is None:
            return _ForwardModel(blocks, off, code, _capi.JAC_ANALYTIC, 0.0)
        return _ForwardModel(blocks, off, code, _capi.JAC_FINITE_DIFFERENCE, float(differential))


# ------------------------------------------------------------------------------------------ likelihoods
@dataclass
class _LikelihoodSpec:
    forwardModel: _ForwardModel
    distribution: int
    measurements: np.ndarray    # Uniform: upper bounds
    uncertainties: np.ndarray   # Uniform: lower bounds


class GaussianLinearLikelihood:
    @staticmethod
    def of(coefficients, measurements, uncertainties, priorLabel: str, evalCacheDepth=None) -> _LikelihoodSpec:
        fm = LinearForwardModel.of(priorLabel, coefficients)
        y, ci = _arr(measurements), _arr(uncertainties)
        assert fm.rangeDimension == y.size, "range dimension mismatch (GaussianLinearLikelihood.scala:43-45)"
        return _LikelihoodSpec(fm, _capi.DIST_GAUSSIAN, y, ci)


def GaussianLikelihood(forwardModel: _ForwardModel, measurements, uncertainties) -> _LikelihoodSpec:
    y, ci = _arr(measurements), _arr(uncertainties)
    assert forwardModel.rangeDimension == y.size
    return _LikelihoodSpec(forwardModel, _capi.DIST_GAUSSIAN, y, ci)


class CauchyLikelihood:
    def __new__(cls, forwardModel: _ForwardModel, measurements, uncertainties) -> _LikelihoodSpec:
        y, ci = _arr(measurements), _arr(uncertainties)
        assert forwardModel.rangeDimension == y.size
        return _LikelihoodSpec(forwardModel, _capi.DIST_CAUCHY, y, ci)

    @staticmethod
    def of(coefficients, measurements, uncertainties, priorLabel: str, evalCacheDepth=None) -> _LikelihoodSpec:
        return CauchyLikelihood(LinearForwardModel.of(priorLabel, coefficients), measurements, uncertainties)


def UniformLikelihood(forwardModel: _ForwardModel, upperBounds, lowerBounds) -> _LikelihoodSpec:
    hi, lo = _arr(upperBounds), _arr(lowerBounds)
    assert hi.size == lo.size
    return _LikelihoodSpec(forwardModel, _capi.DIST_UNIFORM, hi, lo)


# -------------------------------------------------------------------------------------------- posterior
class UnnormalisedPosterior:
    """``UnnormalisedPosterior(priors, likelihoods)``; the model is uploaded to the current CUDA device.

    ``logPdfAt`` / ``logPdfGradientAt`` take the reference's ``Map[String, Vector[Double]]``;
    the ``*Batch`` methods take flat row-major (B, d) arrays in label-sorted order.
    """

    def __init__(self, priors: Sequence[_PriorSpec], likelihoods: Sequence[_LikelihoodSpec] = (),
                 cauchyReferenceSign: bool = True, fdIncremental: bool = False, kernelPath: int = _capi.KERNEL_AUTO,
                 fusePrior: int = 1):
        L = lib()
        self._lib = L
        h = C.c_void_p()
        check(L.thy_model_create(C.byref(h)))
        self._h = h
        self._user_links: Dict[UserLink, int] = {}
        try:
            for p in priors:
                lab = p.label.encode()
                if p.kind == "gaussian_ci":
                    check(L.thy_model_add_gaussian_prior_ci(h, lab, p.dimension, _ptr(p.a), _ptr(p.b)))
                elif p.kind == "gaussian_cov":
                    check(L.thy_model_add_gaussian_prior_cov(h, lab, p.dimension, _ptr(p.a), _ptr(p.b)))
                elif p.kind == "uniform":
                    check(L.thy_model_add_uniform_prior(h, lab, p.dimension, _ptr(p.a), _ptr(p.b)))
                elif p.kind == "cauchy":
                    check(L.thy_model_add_cauchy_prior(h, lab, p.dimension, _ptr(p.a), _ptr(p.b)))
                else:
                    raise ValueError(p.kind)
            for lk in likelihoods:
                self._add_likelihood(lk)
            check(L.thy_model_set_option(h, _capi.OPT_CAUCHY_REFERENCE_SIGN, 1 if cauchyReferenceSign else 0))
            check(L.thy_model_set_option(h, _capi.OPT_FD_INCREMENTAL, int(fdIncremental)))
            check(L.thy_model_set_option(h, _capi.OPT_KERNEL_PATH, int(kernelPath)))
            check(L.thy_model_set_option(h, _capi.OPT_FUSE_PRIOR, int(fusePrior)))
            check(L.thy_model_finalize(h))
            d = C.c_int()
            check(L.thy_model_dimension(h, C.byref(d)))
            self.domainDimension = d.value
            self.layout: Dict[str, tuple] = {}
            for p in priors:
                o, w = C.c_int(), C.c_int()
                check(L.thy_model_label_layout(h, p.label.encode(), C.byref(o), C.byref(w)))
                self.layout[p.label] = (o.value, w.value)
        except Exception:
            L.thy_model_destroy(h)
            self._h = None
            raise

    def _add_likelihood(self, lk: _LikelihoodSpec) -> None:
        fm = lk.forwardModel
        labels = list(fm.blocks)
        nb = len(labels)
        lab_arr = (C.c_char_p * nb)(*[s.encode() for s in labels])
        dims = (C.c_int * nb)(*[fm.blocks[s].shape[1] for s in labels])
        mats = [np.ascontiguousarray(fm.blocks[s]) for s in labels]
        blk = (C.POINTER(C.c_double) * nb)(*[_ptr(m) for m in mats])
        link = fm.link
        if isinstance(link, UserLink):
            if link not in self._user_links:
                out = C.c_int()
                check(self._lib.thy_model_register_link(self._h, link.evaluation.encode(),
                                                        link.derivative.encode() if link.derivative else None, C.byref(out)))
                self._user_links[link] = out.value
            link = self._user_links[link]
        desc = _capi.LikelihoodDesc()
        desc.distribution = lk.distribution
        desc.link = link
        desc.jacobian = fm.jacobian
        desc.n = fm.rangeDimension
        desc.nblocks = nb
        desc.labels = lab_arr
        desc.dims = dims
        desc.blocks = blk
        desc.offset = _ptr(fm.offset) if fm.offset is not None else None
        desc.measurements = _ptr(lk.measurements)
        desc.uncertainties = _ptr(lk.uncertainties)
        desc.differential = fm.differential
        check(self._lib.thy_model_add_likelihood(self._h, C.byref(desc)))

    def collapsed(self) -> "UnnormalisedPosterior":
        """The same posterior with every GaussianLinearLikelihood term folded into ONE d-observation linear-Gaussian
        term (``thy_model_collapse``; SURVEY.md 8(d) reduced-flop form): identical log-density and gradient at every
        point for 4 d^2 instead of 4 n d flops per evaluation.  Priors (of any kind) and the label layout are kept, so
        the samplers and the SLQ engine take the collapsed posterior unchanged."""
        h = C.c_void_p()
        check(self._lib.thy_model_collapse(self._h, C.byref(h)))
        other = object.__new__(UnnormalisedPosterior)
        other._lib, other._h = self._lib, h
        other.domainDimension, other.layout = self.domainDimension, dict(self.layout)
        return other

    def gaussianPosterior(self):
        """(mean, covariance, precision, log_evidence, max_logpdf) of the analytic Gaussian posterior
        (``thy_model_gaussian_posterior``; GaussianAnalyticPosterior.scala:141-169)."""
        d = self.domainDimension
        mean, cov, prec = np.empty(d), np.empty((d, d)), np.empty((d, d))
        lz, mx = C.c_double(), C.c_double()
        check(self._lib.thy_model_gaussian_posterior(self._h, _ptr(mean), _ptr(cov), _ptr(prec), C.byref(lz), C.byref(mx)))
        return mean, cov, prec, lz.value, mx.value

    # -- lifetime
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.thy_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # -- flat <-> labelled (ModelParameterContext.scala:42-80)
    def flatten(self, params: Mapping[str, Sequence[float]]) -> np.ndarray:
        x = np.zeros(self.domainDimension)
        for label, v in params.items():
            if label not in self.layout:
                raise ThylacineError(_capi.THY_ERR_UNKNOWN_LABEL, f"unknown label '{label}'")
            o, w = self.layout[label]
            x[o:o + w] = v
        return x

    def unflatten(self, x) -> Dict[str, np.ndarray]:
        x = np.asarray(x)
        return {label: x[..., o:o + w] for label, (o, w) in self.layout.items()}

    # -- reference surface
    def logPdfAt(self, params: Mapping[str, Sequence[float]]) -> float:
        return float(self.logPdfBatch(self.flatten(params)[None, :])[0])

    def logPdfGradientAt(self, params: Mapping[str, Sequence[float]]) -> Dict[str, np.ndarray]:
        _, g = self.logPdfGradBatch(self.flatten(params)[None, :])
        return self.unflatten(g[0])

    # -- batched surface (host buffers)
    def logPdfBatch(self, X) -> np.ndarray:
        X = _arr(X).reshape(-1, self.domainDimension)
        out = np.empty(X.shape[0])
        check(self._lib.thy_logpdf_batch(self._h, X.shape[0], _ptr(X), _ptr(out)))
        return out

    def logPdfGradBatch(self, X):
        X = _arr(X).reshape(-1, self.domainDimension)
        out = np.empty(X.shape[0])
        grad = np.empty_like(X)
        check(self._lib.thy_logpdf_grad_batch(self._h, X.shape[0], _ptr(X), _ptr(out), _ptr(grad)))
        return out, grad

    def logPdfGradBatchAsync(self, X: np.ndarray, out_logp: np.ndarray, out_grad: Optional[np.ndarray]) -> int:
        """Pipelined host-buffer evaluation (``thy_logpdf_grad_batch_async``): queues copies and kernels and returns a
        ticket; ``X`` / ``out_*`` (C-contiguous fp64, ideally pinned) belong to the library until ``wait(ticket)``."""
        import ctypes as C
        B = X.shape[0]
        assert X.flags.c_contiguous and X.dtype == np.float64 and X.shape[1] == self.domainDimension
        assert out_logp.flags.c_contiguous and out_logp.size >= B
        assert out_grad is None or (out_grad.flags.c_contiguous and out_grad.size >= B * self.domainDimension)
        ticket = C.c_int64()
        check(self._lib.thy_logpdf_grad_batch_async(self._h, B, _ptr(X), _ptr(out_logp),
                                                    _ptr(out_grad) if out_grad is not None else None, C.byref(ticket)))
        return int(ticket.value)

    def wait(self, ticket: int = 0) -> None:
        """Block until the async call ``ticket`` (0: every outstanding call) has delivered its outputs."""
        check(self._lib.thy_model_wait(self._h, ticket))

    # -- batched surface (device pointers: torch CUDA fp64 tensors or raw ints)
    def logPdfGradBatchDevice(self, B: int, x_ptr: int, logp_ptr: int, grad_ptr: Optional[int]) -> None:
        check(self._lib.thy_logpdf_grad_batch_dev(self._h, B, x_ptr, logp_ptr, grad_ptr))

    def samplePriors(self, B: int = 1, seed: int = 0) -> np.ndarray:
        """Posterior.samplePriors (Posterior.scala:64-65), B independent draws -> (B, d)."""
        out = np.empty((B, self.domainDimension))
        check(self._lib.thy_sample_priors(self._h, B, seed, _ptr(out)))
        return out

    def setStream(self, cuda_stream: int) -> None:
        check(self._lib.thy_model_set_stream(self._h, cuda_stream))

    def synchronize(self) -> None:
        check(self._lib.thy_model_synchronize(self._h))

    @property
    def lastKernelPath(self) -> str:
        return self._lib.thy_model_last_kernel_path(self._h).decode()


def kernelLaunchCount() -> int:
    return int(lib().thy_kernel_launch_count())
