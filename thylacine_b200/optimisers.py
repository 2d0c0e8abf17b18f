"""Point optimisers as thin batched clients of the device posterior ("next" row 1).

Host-side mirror of the reference's four ``*OptimisedPosterior`` case classes; every log-pdf / gradient goes through
``thy_logpdf_batch`` / ``thy_logpdf_grad_batch``.  Where the reference evaluates several points with ``parTraverse``
(prior draws that set the scale, the two nudges of a coordinate, all vertices of a simplex, the two probes of a line)
they are evaluated here as ONE batch.  Paths relative to /root/reference/src/main/scala/thylacine/:

    model/optimization/ModelParameterOptimizer.scala:28-43          findMaximumLogPdf / findMaximumPdf
    model/optimization/hookeandjeeves/HookeAndJeevesEngine.scala    :46-152
    model/optimization/hookeandjeeves/CoordinateSlideEngine.scala   :35-58
    model/optimization/mds/MdsEngine.scala:32-145, mds/ModelParameterSimplex.scala:25-123
    model/optimization/gradientdescent/ConjugateGradientEngine.scala:31-119
    model/optimization/line/GoldenSectionSearch.scala:30-211, line/LineProbe.scala:27-75
    config/{HookeAndJeeves,CoordinateSlide,Mds,ConjugateGradient}Config.scala

The reference shuffles with the unseeded ``scala.util.Random``; here the order comes from ``numpy.random.default_rng(rngSeed)``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, Mapping, Optional, Sequence, Tuple

import numpy as np

from .api import UnnormalisedPosterior

INVERSE_PHI = (math.sqrt(5.0) - 1.0) / 2.0           # GoldenSectionSearch.scala:208
INVERSE_PHI_SQUARED = (3.0 - math.sqrt(5.0)) / 2.0   # :209


@dataclass
class OptimisationTelemetryUpdate:
    """core/telemetry/OptimisationTelemetryUpdate.scala"""
    maxLogPdf: float
    currentScale: float
    prefix: str


@dataclass
class HookeAndJeevesConfig:
    convergenceThreshold: float
    numberOfPriorSamplesToSetScale: Optional[int] = None


@dataclass
class CoordinateSlideConfig:
    convergenceThreshold: float
    goldenSectionTolerance: float
    lineProbeExpansionFactor: float
    numberOfPriorSamplesToSetScale: Optional[int] = None


@dataclass
class MdsConfig:
    convergenceThreshold: float
    expansionMultiplier: float
    contractionMultiplier: float
    numberOfPriorSamplesToSetStartingPoint: Optional[int] = None


@dataclass
class ConjugateGradientConfig:
    convergenceThreshold: float
    goldenSectionTolerance: float
    lineProbeExpansionFactor: float
    minimumNumberOfIterations: int


Evaluation = Tuple[float, np.ndarray]


class _Optimiser:
    telemetryPrefix = ""

    def __init__(self, posterior: UnnormalisedPosterior, iterationUpdateCallback=None, isConvergedCallback=None,
                 rngSeed: int = 0):
        self.posterior = posterior
        self._update_cb: Optional[Callable[[OptimisationTelemetryUpdate], None]] = iterationUpdateCallback
        self._converged_cb: Optional[Callable[[], None]] = isConvergedCallback
        self._rng = np.random.default_rng(rngSeed)
        self._prior_seed = rngSeed
        self.evaluations = 0

    # -- batched evaluation helpers
    def _logp(self, points: np.ndarray) -> np.ndarray:
        points = np.atleast_2d(points)
        self.evaluations += points.shape[0]
        return self.posterior.logPdfBatch(points)

    def _logp1(self, x: np.ndarray) -> float:
        return float(self._logp(x[None, :])[0])

    def _best_prior_sample(self, n: int):
        """n prior draws evaluated in one batch -> (per-dimension spread, best (logPdf, point))."""
        xs = self.posterior.samplePriors(n, seed=self._prior_seed)
        lp = self._logp(xs)
        k = int(np.argmax(lp))
        return float(np.max(xs.max(axis=0) - xs.min(axis=0))), (float(lp[k]), xs[k].copy())

    def _telemetry(self, max_logpdf: float, scale: float) -> None:
        if self._update_cb is not None:
            self._update_cb(OptimisationTelemetryUpdate(float(max_logpdf), float(scale), self.telemetryPrefix))

    def _converged(self) -> None:
        if self._converged_cb is not None:
            self._converged_cb()

    # -- reference surface (ModelParameterOptimizer.scala:33-42)
    def findMaximumLogPdf(self, startPt: Mapping[str, Sequence[float]]):
        start = self.posterior.flatten(startPt) if startPt else None
        lp, x = self._calculate_maximum(start)
        return lp, {k: np.array(v) for k, v in self.posterior.unflatten(x).items()}

    def findMaximumPdf(self, startPt: Mapping[str, Sequence[float]]):
        lp, x = self.findMaximumLogPdf(startPt)
        return math.exp(lp), x

    def _calculate_maximum(self, start: Optional[np.ndarray]) -> Evaluation:
        raise NotImplementedError


# ------------------------------------------------------------------------------------------- line searches
class _GoldenSection:
    """GoldenSectionSearch.scala:30-211 + LineProbe.scala:27-75 over a batched log-pdf."""

    def __init__(self, logp_batch: Callable[[np.ndarray], np.ndarray], tolerance: float, expansion: float,
                 rng: np.random.Generator):
        self._logp = logp_batch
        self.tolerance = tolerance
        self.expansion = expansion
        self._rng = rng

    def search_colinear_triple(self, start: Evaluation, mid: Evaluation, end: Evaluation) -> Evaluation:   # :35-53
        if start[0] == end[0]:
            best_new = (start, end)[int(self._rng.integers(2))]
        else:
            best_new = start if start[0] > end[0] else end
        if best_new[0] > mid[0]:
            return self.search_along_line_joining(mid, best_new)
        return self._golden(start, end)

    def search_direction_along(self, start: Evaluation, direction: np.ndarray) -> Evaluation:            # :55-101
        unit = direction / np.linalg.norm(direction)
        step = self.tolerance / 10.0
        while True:
            fwd, rev = start[1] + step * unit, start[1] - step * unit
            lp = self._logp(np.stack([fwd, rev]))
            if lp[0] == lp[1] and start[0] == lp[0]:
                step *= self.expansion      # flat to working precision: widen the probe
                continue
            return self.search_colinear_triple((float(lp[1]), rev), start, (float(lp[0]), fwd))

    def search_along_line_joining(self, a: Evaluation, b: Evaluation) -> Evaluation:                       # :103-119
        lo, _, hi = self._probe(a, b)
        return self._golden(lo, hi)

    def _probe(self, p1: Evaluation, p2: Evaluation, ordered: bool = False):                               # LineProbe.scala:32-74
        lower, upper = (p1, p2) if (ordered or p1[0] < p2[0]) else (p2, p1)
        while True:
            x = self.expansion * (upper[1] - lower[1]) + upper[1]
            ev = (float(self._logp(x[None, :])[0]), x)
            if ev[0] >= upper[0]:
                lower, upper = upper, ev
                continue
            return lower, upper, ev

    def _golden(self, first: Evaluation, second: Evaluation) -> Evaluation:                                # :136-204
        h = second[1] - first[1]
        c_ev: Optional[Evaluation] = None
        d_ev: Optional[Evaluation] = None
        while True:
            if np.linalg.norm(h) <= self.tolerance:
                return first if first[0] >= second[0] else second
            new_h = h * INVERSE_PHI
            xc, xd = first[1] + h * INVERSE_PHI_SQUARED, first[1] + new_h
            need = [x for x, ev in ((xc, c_ev), (xd, d_ev)) if ev is None]
            vals = list(self._logp(np.stack(need))) if need else []
            if c_ev is None:
                c_ev = (float(vals.pop(0)), xc)
            if d_ev is None:
                d_ev = (float(vals.pop(0)), xd)
            if c_ev[0] > d_ev[0]:
                second, d_ev, c_ev = d_ev, c_ev, None
            else:
                first, c_ev, d_ev = c_ev, d_ev, None
            h = new_h


# --------------------------------------------------------------------------------------- Hooke and Jeeves
class HookeAndJeevesOptimisedPosterior(_Optimiser):
    """``HookeAndJeevesOptimisedPosterior(hookeAndJeevesConfig, posterior, iterationUpdateCallback, isConvergedCallback)``
    (components/posterior/HookeAndJeevesOptimisedPosterior.scala:28-65; engine HookeAndJeevesEngine.scala:46-152)."""
    telemetryPrefix = "Hooke and Jeeves"

    def __init__(self, hookeAndJeevesConfig, posterior, iterationUpdateCallback=None, isConvergedCallback=None, rngSeed=0):
        super().__init__(posterior, iterationUpdateCallback, isConvergedCallback, rngSeed)
        self.config = hookeAndJeevesConfig
        self.convergenceThreshold = hookeAndJeevesConfig.convergenceThreshold
        self.numberOfSamplesToSetScale = hookeAndJeevesConfig.numberOfPriorSamplesToSetScale or 100

    def _nudge_pair(self, index: int, amount: float, x: np.ndarray):
        pts = np.stack([x, x])
        pts[0, index] -= amount
        pts[1, index] += amount
        lp = self._logp(pts)
        return (float(lp[0]), pts[0]), (float(lp[1]), pts[1])

    def _dimension_scan(self, amount: float, best: Evaluation) -> Evaluation:                              # :85-101
        for index in self._rng.permutation(best[1].size):
            minus, plus = self._nudge_pair(int(index), amount, best[1])
            cands = [minus, plus] if self._rng.integers(2) else [plus, minus]
            for cand in cands:                          # maxBy keeps the first maximum: the incumbent wins ties
                if cand[0] > best[0]:
                    best = cand
        return best

    def _calculate_maximum(self, start: Optional[np.ndarray]) -> Evaluation:                               # :103-151
        scale, best = self._best_prior_sample(self.numberOfSamplesToSetScale)
        if start is not None:
            best = (self._logp1(start), start.copy())
        while True:
            scan = self._dimension_scan(scale, best)
            if scan[0] > best[0]:
                self._telemetry(scan[0], scale)
            elif scale > self.convergenceThreshold:
                scale /= 2.0
                self._telemetry(scan[0], scale)
            else:
                self._converged()
                return best
            best = scan


class CoordinateSlideOptimisedPosterior(HookeAndJeevesOptimisedPosterior):
    """Hooke and Jeeves with a golden-section line search along each coordinate (CoordinateSlideEngine.scala:35-58)."""
    telemetryPrefix = "Coordinate Slide"

    def __init__(self, coordinateSlideConfig, posterior, iterationUpdateCallback=None, isConvergedCallback=None, rngSeed=0):
        _Optimiser.__init__(self, posterior, iterationUpdateCallback, isConvergedCallback, rngSeed)
        self.config = coordinateSlideConfig
        self.convergenceThreshold = coordinateSlideConfig.convergenceThreshold
        self.numberOfSamplesToSetScale = coordinateSlideConfig.numberOfPriorSamplesToSetScale or 100
        self._line = _GoldenSection(self._logp, coordinateSlideConfig.goldenSectionTolerance,
                                    coordinateSlideConfig.lineProbeExpansionFactor, self._rng)

    def _dimension_scan(self, amount: float, best: Evaluation) -> Evaluation:
        for index in self._rng.permutation(best[1].size):
            minus, plus = self._nudge_pair(int(index), amount, best[1])
            best = self._line.search_colinear_triple(minus, best, plus)
        return best


# ------------------------------------------------------------------------------------------------------- MDS
class MdsOptimisedPosterior(_Optimiser):
    """Multi-directional search on a regular simplex (MdsEngine.scala:32-145, ModelParameterSimplex.scala:25-123).
    All d vertices other than the pivot are evaluated as one batch per reflection / expansion / contraction."""
    telemetryPrefix = "MDS"

    def __init__(self, mdsConfig, posterior, iterationUpdateCallback=None, isConvergedCallback=None, rngSeed=0):
        super().__init__(posterior, iterationUpdateCallback, isConvergedCallback, rngSeed)
        self.config = mdsConfig

    @staticmethod
    def unitRegularCenteredOn(x: np.ndarray) -> np.ndarray:                                              # Simplex :93-120
        n = x.size
        base = -n ** -1.5 * (math.sqrt(n + 1.0) + 1.0)
        nudge = math.sqrt(1.0 + 1.0 / n)
        verts = np.full((n + 1, n), base)
        verts[np.arange(n), np.arange(n)] += nudge
        verts[n, :] = n ** -0.5
        return verts + x[None, :]

    def _others(self, verts: np.ndarray, pivot: int) -> Tuple[int, float]:
        idx = [i for i in range(verts.shape[0]) if i != pivot]
        lp = self._logp(verts[idx])
        k = int(np.argmax(lp))
        return idx[k], float(lp[k])

    def _calculate_maximum(self, start: Optional[np.ndarray]) -> Evaluation:
        cfg = self.config
        if start is None:
            _, (_, start) = self._best_prior_sample(cfg.numberOfPriorSamplesToSetStartingPoint or 100)
        verts = self.unitRegularCenteredOn(start)
        lp = self._logp(verts)
        best = (int(np.argmax(lp)), float(np.max(lp)))
        while True:
            pivot = verts[best[0]].copy()
            reflected = 2.0 * pivot[None, :] - verts                                                   # reflectAbout :66-69
            new_best = self._others(reflected, best[0])
            if new_best[1] > best[1]:
                expanded = (1.0 - cfg.expansionMultiplier) * pivot[None, :] + cfg.expansionMultiplier * reflected
                exp_best = self._others(expanded, best[0])
                best, verts = (exp_best, expanded) if exp_best[1] > new_best[1] else (new_best, reflected)
            else:
                contracted = (1.0 + cfg.contractionMultiplier) * pivot[None, :] - cfg.contractionMultiplier * reflected
                con_best = self._others(contracted, best[0])
                best, verts = (con_best if con_best[1] > best[1] else best), contracted
            # the simplex stays regular under these maps, so one adjacent edge measures its size (:81-88)
            measure = float(np.linalg.norm(verts[0] - verts[1]))
            self._telemetry(best[1], measure)
            if measure <= cfg.convergenceThreshold:
                self._converged()
                return best[1], verts[best[0]].copy()


# ---------------------------------------------------------------------------------------- conjugate gradient
class ConjugateGradientOptimisedPosterior(_Optimiser):
    """Polak-Ribiere(+) non-linear conjugate gradient with golden-section line searches
    (ConjugateGradientEngine.scala:31-119).  ``maxIterations`` bounds the loop (the reference relies on
    ``minimumNumberOfIterations`` and its threshold alone)."""
    telemetryPrefix = "Conjugate Gradient"

    def __init__(self, conjugateGradientConfig, posterior, iterationUpdateCallback=None, isConvergedCallback=None,
                 rngSeed=0, maxIterations: int = 1_000_000):
        super().__init__(posterior, iterationUpdateCallback, isConvergedCallback, rngSeed)
        self.config = conjugateGradientConfig
        self.maxIterations = maxIterations
        self._line = _GoldenSection(self._logp, conjugateGradientConfig.goldenSectionTolerance,
                                    conjugateGradientConfig.lineProbeExpansionFactor, self._rng)

    def _calculate_maximum(self, start: Optional[np.ndarray]) -> Evaluation:
        if start is None:
            start = np.zeros(self.posterior.domainDimension)   # findMaximumLogPdf(Map()): missing entries read as 0
        lp, g = self.posterior.logPdfGradBatch(start[None, :])
        self.evaluations += 1
        current: Evaluation = (float(lp[0]), start.copy())
        prev_grad = -g[0]
        prev_dir = prev_grad.copy()
        iteration = 1
        while True:
            _, g = self.posterior.logPdfGradBatch(current[1][None, :])
            self.evaluations += 1
            neg_grad = -g[0]
            mag2 = float(neg_grad @ neg_grad)
            if mag2 == 0.0:                       # exactly stationary: the reference would divide by zero here
                self._converged()
                return current
            beta = max(float((neg_grad - prev_grad) @ neg_grad) / mag2, 0.0)                              # :57-64
            direction = neg_grad + beta * prev_dir
            nxt = self._line.search_direction_along(current, direction)
            diff = nxt[0] - current[0]
            self._telemetry(nxt[0], diff)
            if (diff > self.config.convergenceThreshold or iteration < self.config.minimumNumberOfIterations) \
                    and iteration < self.maxIterations:
                current, prev_grad, prev_dir, iteration = nxt, neg_grad, direction, iteration + 1
                continue
            self._converged()
            return nxt
