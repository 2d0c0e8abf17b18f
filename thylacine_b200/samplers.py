"""Samplers and integrators built on the batched posterior -- host-side mirror of the reference's
``HmcmcSampledPosterior`` / ``LeapfrogMcmcSampledPosterior`` / ``SlqIntegratedPosterior`` case classes.

Paths relative to /root/reference/src/main/scala/thylacine/:
    config/HmcmcConfig.scala:20-25, model/components/posterior/HmcmcSampledPosterior.scala:63-75,
    model/sampling/hmcmc/HmcmcEngine.scala:55-214, model/sampling/ModelParameterSampler.scala:27-38
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Dict, List, Mapping, Optional, Sequence

import numpy as np

from . import _capi
from ._capi import HmcmcConfigC, check, lib, register
from .api import UnnormalisedPosterior, _arr, _ptr

_vp, _dp = C.c_void_p, C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)

register("thy_hmc_create", C.c_int, [_vp, C.POINTER(HmcmcConfigC), C.c_int64, C.c_uint64, _dp, C.POINTER(_vp)])
register("thy_hmc_destroy", C.c_int, [_vp])
register("thy_hmc_set_use_graph", C.c_int, [_vp, C.c_int])
register("thy_hmc_burn_in", C.c_int, [_vp])
register("thy_hmc_run_trajectories", C.c_int, [_vp, C.c_int])
register("thy_hmc_sample", C.c_int, [_vp, C.c_int, _dp, C.c_int])
register("thy_hmc_sample_dev", C.c_int, [_vp, C.c_int, _vp, C.c_int])
register("thy_hmc_stats", C.c_int, [_vp, _i64p, _i64p, _dp])
register("thy_hmc_get_state", C.c_int, [_vp, _dp, _dp])
register("thy_hmc_set_state", C.c_int, [_vp, _dp, C.c_uint64])
register("thy_hmc_get_trajectory_counter", C.c_int, [_vp, C.POINTER(C.c_uint64)])
register("thy_hmc_set_fused_steps", C.c_int, [_vp, C.c_int])
register("thy_lfm_create", C.c_int, [_vp, C.c_void_p, C.c_int, C.c_uint64, _dp, C.POINTER(_vp)])
register("thy_lfm_destroy", C.c_int, [_vp])
register("thy_lfm_run_sweeps", C.c_int, [_vp, C.c_int])
register("thy_lfm_sample", C.c_int, [_vp, C.c_int, _dp])
register("thy_lfm_stats", C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)])
register("thy_lfm_get_pool", C.c_int, [_vp, _dp, _dp])
register("thy_lfm_set_pool", C.c_int, [_vp, _dp])


class _LeapfrogMcmcConfigC(C.Structure):
    _fields_ = [("stepsBetweenSamples", C.c_int), ("warmupStepCount", C.c_int), ("samplePoolSize", C.c_int)]


@dataclass
class HmcmcConfig:
    """config/HmcmcConfig.scala:20-25"""
    stepsBetweenSamples: int
    stepsInDynamicsSimulation: int
    warmupStepCount: int
    dynamicsSimulationStepSize: float


@dataclass
class HmcmcTelemetryUpdate:
    """core/telemetry/HmcmcTelemetryUpdate.scala:20-25"""
    samplesRemaining: int
    jumpAttempts: int
    jumpAcceptances: int
    hamiltonianDifferential: Optional[float]


class HmcmcSampledPosterior:
    """``HmcmcSampledPosterior(hmcmcConfig, posterior, telemetryUpdateCallback, seed)``.

    ``numberOfChains`` independent chains run in lock-step on the device (the reference runs one);
    ``seed`` is the reference's starting point (a ``Map[String, Vector[Double]]``, shared by all chains,
    or a (numberOfChains, d) array); omitted => prior draws (HmcmcEngine.scala:183-188).
    ``sample(n)`` reproduces the reference: burn-in re-runs on every call (:205-207), stepsBetweenSamples+1
    trajectories per sample (:168), fully rejected blocks add no sample (:149-152).
    """

    def __init__(self, hmcmcConfig: HmcmcConfig, posterior: UnnormalisedPosterior,
                 telemetryUpdateCallback: Optional[Callable[[HmcmcTelemetryUpdate], None]] = None,
                 seed=None, numberOfChains: int = 1, rngSeed: int = 0, useGraph: bool = True, fusedSteps=None):
        self.posterior = posterior
        self.config = hmcmcConfig
        self.numberOfChains = int(numberOfChains)
        self.telemetryUpdateCallback = telemetryUpdateCallback
        self._lib = lib()
        cfg = HmcmcConfigC(hmcmcConfig.stepsBetweenSamples, hmcmcConfig.stepsInDynamicsSimulation,
                           hmcmcConfig.warmupStepCount, float(hmcmcConfig.dynamicsSimulationStepSize))
        seed_arr = None
        if seed is not None:
            if isinstance(seed, Mapping):
                flat = posterior.flatten(seed)
                seed_arr = np.ascontiguousarray(np.tile(flat[None, :], (self.numberOfChains, 1)))
            else:
                seed_arr = _arr(seed).reshape(self.numberOfChains, posterior.domainDimension)
        h = C.c_void_p()
        check(self._lib.thy_hmc_create(posterior.handle, C.byref(cfg), self.numberOfChains, rngSeed,
                                       _ptr(seed_arr) if seed_arr is not None else None, C.byref(h)))
        self._h = h
        if not useGraph:
            check(self._lib.thy_hmc_set_use_graph(h, 0))
        if fusedSteps is not None:   # None: the library's cost model; True / False: always / never one launch per step
            check(self._lib.thy_hmc_set_fused_steps(h, 2 if fusedSteps else 0))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.thy_hmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # -- reference surface
    def sample(self, numberOfSamples: int) -> List[Dict[str, np.ndarray]]:
        """Samples of chain 0 as labelled collections (the reference returns a Set of these)."""
        chains = self.sampleChains(numberOfSamples)
        return [self.posterior.unflatten(x) for x in chains[0]]

    # -- batched surface
    def sampleChains(self, numberOfSamples: int, rerunBurnIn: bool = True) -> np.ndarray:
        out = np.empty((self.numberOfChains, numberOfSamples, self.posterior.domainDimension))
        check(self._lib.thy_hmc_sample(self._h, numberOfSamples, _ptr(out), 1 if rerunBurnIn else 0))
        if self.telemetryUpdateCallback is not None:
            att, acc, dh = self.stats()
            self.telemetryUpdateCallback(HmcmcTelemetryUpdate(0, int(att[0]), int(acc[0]), float(dh[0])))
        return out

    def burnIn(self) -> None:
        check(self._lib.thy_hmc_burn_in(self._h))

    def setState(self, x, trajectoryCounter: int) -> None:
        """Resume from a checkpoint: positions (numberOfChains, d) + the trajectory counter that keys every draw."""
        x = np.ascontiguousarray(_arr(x).reshape(self.numberOfChains, self.posterior.domainDimension))
        check(self._lib.thy_hmc_set_state(self._h, _ptr(x), int(trajectoryCounter)))

    def trajectoryCounter(self) -> int:
        t = C.c_uint64()
        check(self._lib.thy_hmc_get_trajectory_counter(self._h, C.byref(t)))
        return int(t.value)

    def runTrajectories(self, n: int) -> None:
        check(self._lib.thy_hmc_run_trajectories(self._h, n))

    def stats(self):
        att = np.empty(self.numberOfChains, dtype=np.int64)
        acc = np.empty(self.numberOfChains, dtype=np.int64)
        dh = np.empty(self.numberOfChains)
        check(self._lib.thy_hmc_stats(self._h, att.ctypes.data_as(_i64p), acc.ctypes.data_as(_i64p), _ptr(dh)))
        return att, acc, dh

    def state(self):
        x = np.empty((self.numberOfChains, self.posterior.domainDimension))
        lp = np.empty(self.numberOfChains)
        check(self._lib.thy_hmc_get_state(self._h, _ptr(x), _ptr(lp)))
        return x, lp


@dataclass
class LeapfrogMcmcConfig:
    """config/LeapfrogMcmcConfig.scala:20-24"""
    stepsBetweenSamples: int
    warmupStepCount: int
    samplePoolSize: int


DISTANCES = {"euclidean": 0, "squared_euclidean": 1, "manhattan": 2}


class LeapfrogMcmcSampledPosterior:
    """``LeapfrogMcmcSampledPosterior.of(leapfrogMcmcConfig, distanceCalculation, posterior, ..., seed)``
    (components/posterior/LeapfrogMcmcSampledPosterior.scala:72-106).  ``distanceCalculation`` must name a device
    distance ("euclidean", "squared_euclidean", "manhattan"): a JVM/Python closure cannot run on the GPU."""

    def __init__(self, *a, **k):
        raise TypeError("use LeapfrogMcmcSampledPosterior.of(...), as in the reference")

    @classmethod
    def of(cls, leapfrogMcmcConfig: LeapfrogMcmcConfig, distanceCalculation, posterior: UnnormalisedPosterior,
           sampleRequestSetCallback=None, sampleRequestUpdateCallback=None, seed=None, rngSeed: int = 0):
        if callable(distanceCalculation):
            raise TypeError("distanceCalculation must be the name of a device-registered distance; closures cannot run on the GPU")
        self = object.__new__(cls)
        self.posterior = posterior
        self.config = leapfrogMcmcConfig
        self._lib = lib()
        self._set_cb, self._update_cb = sampleRequestSetCallback, sampleRequestUpdateCallback
        cfg = _LeapfrogMcmcConfigC(leapfrogMcmcConfig.stepsBetweenSamples, leapfrogMcmcConfig.warmupStepCount,
                                   leapfrogMcmcConfig.samplePoolSize)
        seed_arr = None
        if seed is not None:
            seed_arr = posterior.flatten(seed) if isinstance(seed, Mapping) else _arr(seed).reshape(posterior.domainDimension)
        h = C.c_void_p()
        check(self._lib.thy_lfm_create(posterior.handle, C.byref(cfg), DISTANCES[distanceCalculation], rngSeed,
                                       _ptr(seed_arr) if seed_arr is not None else None, C.byref(h)))
        self._h = h
        return self

    def close(self):
        if getattr(self, "_h", None):
            self._lib.thy_lfm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sample(self, numberOfSamples: int) -> List[Dict[str, np.ndarray]]:
        return [self.posterior.unflatten(x) for x in self.sampleArray(numberOfSamples)]

    def sampleArray(self, numberOfSamples: int) -> np.ndarray:
        if self._set_cb is not None:
            self._set_cb(numberOfSamples)
        out = np.empty((numberOfSamples, self.posterior.domainDimension))
        check(self._lib.thy_lfm_sample(self._h, numberOfSamples, _ptr(out)))
        if self._update_cb is not None:
            self._update_cb(0)
        return out

    def runSweeps(self, n: int) -> None:
        check(self._lib.thy_lfm_run_sweeps(self._h, n))

    def stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        check(self._lib.thy_lfm_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def setPool(self, x) -> None:
        """Replace the pool (P, d): warm start near the mode or resume from ``pool()``."""
        x = np.ascontiguousarray(_arr(x).reshape(self.config.samplePoolSize, self.posterior.domainDimension))
        check(self._lib.thy_lfm_set_pool(self._h, _ptr(x)))

    def pool(self):
        P = self.config.samplePoolSize
        x = np.empty((P, self.posterior.domainDimension))
        lp = np.empty(P)
        check(self._lib.thy_lfm_get_pool(self._h, _ptr(x), _ptr(lp)))
        return x, lp


# ----------------------------------------------------------------------------------------------- moments
register("thy_moments_reset_dev", C.c_int, [_vp, _vp])
register("thy_moments_accumulate_dev", C.c_int, [_vp, C.c_int64, _vp, _vp])
register("thy_moments_finish", C.c_int, [_vp, _vp, _vp, _dp, _dp, _dp])


class SampleMoments:
    """Device accumulator of (count, sum x, sum x x^T) over sample blocks; ``finish(comm)`` sums over the ranks
    (NCCL all-reduce) and returns (count, mean, covariance).  Not in the reference."""

    def __init__(self, posterior: UnnormalisedPosterior):
        import torch
        self.posterior = posterior
        d = posterior.domainDimension
        self._acc = torch.zeros(1 + d + d * d, dtype=torch.float64, device="cuda")
        self._lib = lib()
        check(self._lib.thy_moments_reset_dev(posterior.handle, self._acc.data_ptr()))

    def add(self, x_dev) -> None:
        """x_dev: CUDA fp64 tensor (..., d), contiguous; rows starting with NaN are skipped."""
        assert x_dev.is_cuda and x_dev.is_contiguous() and x_dev.shape[-1] == self.posterior.domainDimension
        n = x_dev.numel() // self.posterior.domainDimension
        check(self._lib.thy_moments_accumulate_dev(self.posterior.handle, n, x_dev.data_ptr(), self._acc.data_ptr()))

    def finish(self, comm=None):
        d = self.posterior.domainDimension
        cnt = C.c_double()
        mean, cov = np.empty(d), np.empty((d, d))
        check(self._lib.thy_moments_finish(self.posterior.handle, comm.handle if comm is not None else None,
                                           self._acc.data_ptr(), C.byref(cnt), _ptr(mean), _ptr(cov)))
        return cnt.value, mean, cov


# --------------------------------------------------------------------------------------------------- SLQ
class _SlqConfigC(C.Structure):
    _fields_ = [("poolSize", C.c_int), ("abscissaNumber", C.c_int), ("domainScalingIncrement", C.c_double),
                ("targetAcceptanceProbability", C.c_double), ("sampleParallelism", C.c_int),
                ("maxIterationCount", C.c_int), ("minIterationCount", C.c_int), ("constrainedWalkSteps", C.c_int),
                ("keepRetiredPoints", C.c_int), ("disableFusedWalk", C.c_int), ("referenceProposals", C.c_int)]


class _SlqStatsC(C.Structure):
    _fields_ = [("iterations", C.c_int64), ("total_retired", C.c_int64), ("rounds", C.c_int64),
                ("likelihood_evaluations", C.c_int64), ("log_evidence_mean", C.c_double),
                ("log_evidence_std", C.c_double), ("neg_entropy_mean", C.c_double), ("min_logpdf", C.c_double),
                ("walk_scale", C.c_double), ("walk_acceptance", C.c_double)]


register("thy_slq_create", C.c_int, [_vp, C.POINTER(_SlqConfigC), C.c_uint64, _dp, C.c_int, _vp, C.POINTER(_vp)])
register("thy_slq_destroy", C.c_int, [_vp])
register("thy_slq_run", C.c_int, [_vp, C.POINTER(_SlqStatsC)])
register("thy_slq_run_rounds", C.c_int, [_vp, C.c_int, C.POINTER(C.c_int)])
register("thy_slq_retired", C.c_int, [_vp, C.c_int64, _dp, _i64p])
register("thy_slq_sample", C.c_int, [_vp, C.c_int, C.c_uint64, _dp])
register("thy_slq_abscissa_results", C.c_int, [_vp, _dp, _dp])
register("thy_slq_live_points", C.c_int, [_vp, _i64p, _dp, _dp])
register("thy_slq_retired_log_x", C.c_int, [_vp, C.c_int, C.c_int64, _dp, _i64p])
SLQ_INTEGRAND = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
register("thy_slq_integrate", C.c_int, [_vp, SLQ_INTEGRAND, C.c_void_p, C.c_double, _dp, _dp, _dp])
register("thy_slq_load_balance", C.c_int, [_vp, _dp, _dp])
register("thy_slq_set_trace", C.c_int, [_vp, C.c_int])
register("thy_slq_phase_times", C.c_int, [_vp, _dp, _i64p])
SLQ_PHASES = ("all_gather_keys", "radix_select_flags_values", "walk_adapt", "wait_quadrature", "round_control",
              "pool_spread_side_stream", "graph_replay_bit0_peer_memory_bit1")


@dataclass
class SlqConfig:
    """config/SlqConfig.scala:20-28"""
    poolSize: int
    abscissaNumber: int
    domainScalingIncrement: float
    targetAcceptanceProbability: float
    sampleParallelism: int
    maxIterationCount: int
    minIterationCount: int


@dataclass
class SlqTelemetryUpdate:
    """core/telemetry/SlqTelemetryUpdate.scala:20-29 (domainCubeCount has no analogue: there is no cubical cover)"""
    negEntropyAvg: float
    logPdf: float
    samplePoolMinimumLogPdf: float
    domainVolumeScaling: float
    acceptancesSinceDomainRebuild: int
    samplePoolSize: int
    domainCubeCount: int
    iterationCount: int


class SlqIntegratedPosterior:
    """``SlqIntegratedPosterior.of(slqConfig, posterior, slqTelemetryUpdateCallback, domainRebuildStartCallback,
    domainRebuildFinishCallback, seedsSpec)`` (components/posterior/SlqIntegratedPosterior.scala:87-125) followed by
    ``rebuildSampleSimulation`` (:76-80), ``logEvidence`` / ``integrate`` and ``sample`` (SlqEngine.scala:460-469).

    Re-designed for 10^6 live points (DESIGN.md "SLQ"): K = sampleParallelism points are retired per round; their
    replacements are constrained random walks from surviving live points; the threshold is selected on the device
    and, across GPUs, with one NCCL all-gather per round.  ``comm`` shards the live points across ranks."""

    def __init__(self, *a, **k):
        raise TypeError("use SlqIntegratedPosterior.of(...), as in the reference")

    @classmethod
    def of(cls, slqConfig: SlqConfig, posterior: UnnormalisedPosterior, slqTelemetryUpdateCallback=None,
           domainRebuildStartCallback=None, domainRebuildFinishCallback=None, seedsSpec=None, rngSeed: int = 0,
           comm=None, constrainedWalkSteps: int = 0, keepRetiredPoints: bool = False, fusedWalk: bool = True,
           referenceProposals: bool = False):
        self = object.__new__(cls)
        self.posterior, self.config, self.comm = posterior, slqConfig, comm
        self._cb = slqTelemetryUpdateCallback
        self._lib = lib()
        cfg = _SlqConfigC(slqConfig.poolSize, slqConfig.abscissaNumber, float(slqConfig.domainScalingIncrement),
                          float(slqConfig.targetAcceptanceProbability), slqConfig.sampleParallelism,
                          slqConfig.maxIterationCount, slqConfig.minIterationCount, int(constrainedWalkSteps),
                          1 if keepRetiredPoints else 0, 0 if fusedWalk else 1, 1 if referenceProposals else 0)
        seeds = None
        n_seeds = 0
        if seedsSpec:
            rows = [posterior.flatten(s) if isinstance(s, Mapping) else _arr(s) for s in seedsSpec]
            seeds = np.ascontiguousarray(np.stack(rows))
            n_seeds = seeds.shape[0]
        h = C.c_void_p()
        check(self._lib.thy_slq_create(posterior.handle, C.byref(cfg), rngSeed, _ptr(seeds) if seeds is not None else None,
                                       n_seeds, comm.handle if comm is not None else None, C.byref(h)))
        self._h = h
        self.stats: Optional[_SlqStatsC] = None
        return self

    def close(self):
        if getattr(self, "_h", None):
            self._lib.thy_slq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rebuildSampleSimulation(self) -> None:
        st = _SlqStatsC()
        check(self._lib.thy_slq_run(self._h, C.byref(st)))
        self.stats = st
        if self._cb is not None:
            self._cb(SlqTelemetryUpdate(st.neg_entropy_mean, st.min_logpdf, st.min_logpdf, st.walk_scale, 0,
                                        self.config.poolSize, 0, int(st.iterations)))

    def runRounds(self, n: int) -> bool:
        conv = C.c_int()
        check(self._lib.thy_slq_run_rounds(self._h, n, C.byref(conv)))
        return bool(conv.value)

    @property
    def logEvidence(self) -> float:
        """Mean over the abscissas of ln Z (calibrated: prior-measure nested sampling)."""
        assert self.stats is not None, "call rebuildSampleSimulation first"
        return self.stats.log_evidence_mean

    def abscissaResults(self):
        lz, hh = np.empty(self.config.abscissaNumber), np.empty(self.config.abscissaNumber)
        check(self._lib.thy_slq_abscissa_results(self._h, _ptr(lz), _ptr(hh)))
        return lz, hh

    def livePoints(self):
        """(points, log-likelihoods) of the live pool held by this rank (SlqEngine.scala:120-143 ``samples``)."""
        n = C.c_int64()
        check(self._lib.thy_slq_live_points(self._h, C.byref(n), None, None))
        x, ll = np.empty((n.value, self.posterior.domainDimension)), np.empty(n.value)
        check(self._lib.thy_slq_live_points(self._h, C.byref(n), _ptr(x), _ptr(ll)))
        return x, ll

    def retiredLogLikelihoods(self) -> np.ndarray:
        n = C.c_int64()
        check(self._lib.thy_slq_retired(self._h, 0, None, C.byref(n)))
        out = np.empty(n.value)
        check(self._lib.thy_slq_retired(self._h, n.value, _ptr(out), C.byref(n)))
        return out

    def retiredLogX(self, abscissa: int) -> np.ndarray:
        """ln X_i of one simulated abscissa for the retired points (QuadratureAbscissa.scala:33-73), replayed on the
        device from the counter RNG."""
        n = C.c_int64()
        check(self._lib.thy_slq_retired_log_x(self._h, abscissa, 0, None, C.byref(n)))
        out = np.empty(n.value)
        check(self._lib.thy_slq_retired_log_x(self._h, abscissa, n.value, _ptr(out), C.byref(n)))
        return out

    def integrate(self, integrand: Callable[[float], float], logShift: Optional[float] = None, reference: bool = False) -> float:
        """``integrate`` (SlqEngine.scala:460-466): mean over the simulated abscissas of sum_i w_ai integrand(exp(l_i - s))
        through the C ABI's ``thy_slq_integrate`` (host callback, as the reference's integrand is a JVM closure).
        ``logShift`` None = max l (fp64-safe); NaN = the reference's (max l + min l) / 2.  ``reference=True`` returns the
        statistic the reference's ``getIntegrationStats`` literally computes, mean_a (I_a / Z_a - ln Z_a)."""
        if logShift is None:
            logShift = float(self.retiredLogLikelihoods().max())
        cb = SLQ_INTEGRAND(lambda v, _u: float(integrand(v)))
        stat, integ = C.c_double(), C.c_double()
        check(self._lib.thy_slq_integrate(self._h, cb, None, float(logShift), C.byref(stat), C.byref(integ), None))
        return stat.value if reference else integ.value

    def loadBalance(self):
        """(mean, max) number of points this rank replaced per round."""
        a, b = C.c_double(), C.c_double()
        check(self._lib.thy_slq_load_balance(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def setTrace(self, enable: bool) -> None:
        check(self._lib.thy_slq_set_trace(self._h, 1 if enable else 0))

    def phaseTimes(self) -> Dict[str, float]:
        """Mean microseconds per traced round of each phase (``setTrace(True)`` first)."""
        us = np.zeros(7)
        n = C.c_int64()
        check(self._lib.thy_slq_phase_times(self._h, _ptr(us), C.byref(n)))
        out = {k: float(v) for k, v in zip(SLQ_PHASES, us)}
        out["rounds"] = int(n.value)
        return out

    def sample(self, numberOfSamples: int, rngSeed: int = 1) -> List[Dict[str, np.ndarray]]:
        out = np.empty((numberOfSamples, self.posterior.domainDimension))
        check(self._lib.thy_slq_sample(self._h, numberOfSamples, rngSeed, _ptr(out)))
        return [self.posterior.unflatten(x) for x in out]
