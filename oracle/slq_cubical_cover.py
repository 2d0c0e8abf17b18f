"""TEST INFRASTRUCTURE ONLY: restatement of the reference's SLQ proposal geometry and sampling loop (the cubical cover).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package; the product never does.

Reference: /root/reference/src/main/scala/thylacine/model/integration/slq/
    PointInInterval.scala:40-109        interval around a point, symmetrize, findDisjointBoundary, getSample
    PointInCube.scala:30-117            product of intervals; "volume" = SUM of side lengths (:30-33, sic); makeDisjoint fold
    PointInCubeCollection.scala:60-111  initialise -> makeDisjoint -> symmetrize; volume-weighted cube choice (MathOps.cdfStaircase)
    QuadratureDomainTelemetry.scala:27-110  acceptance / rejection counters, rebuild trigger, exhaustion test (rescaling disabled)
    SlqEngine.scala:107-301,309-345     sampleAndAnalyze / updateStateWithSampleCalculation / recordMinimumLogPdf / rebuildDomain
The reference holds no test for any of this and draws from unseeded generators, so the restatement is
unpinned by golden vectors; it exists so that the device engine's reference-faithful mode (thy_slq_config.referenceProposals)
can be REPLAYED draw for draw: every random number comes from the `uniform(stream, index)` callable the caller injects.
Plain Python loops: small pools only.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Tuple

import numpy as np


# ---------------------------------------------------------------------------------------------- geometry
def initialise_cover(points: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """PointInCubeCollection.initialise (:60-85): bounds symmetric about each point, half-width = the largest distance to
    any other point along that dimension.  points (P, d) -> (lower, upper), each (P, d)."""
    diff = np.abs(points[:, None, :] - points[None, :, :])       # (P, P, d)
    half = diff.max(axis=1)
    return points - half, points + half


def _intersecting(lo1, up1, lo2, up2) -> bool:
    # PointInInterval.isIntersectingWith (:63-67) in every dimension (PointInCube.isIntersectingWith :66-69)
    a = (up1 > lo2) & (lo1 < up2)
    b = (up2 > lo1) & (lo2 < up1)
    return bool(np.all(a | b))


def _find_disjoint_boundary(p1, lo1, up1, p2, lo2, up2):
    """PointInInterval.findDisjointBoundary (:82-109) for ONE dimension; returns the two new (lower, upper) pairs."""
    one_larger = p1 > p2
    (lp, llo, lup), (sp, slo, sup) = ((p1, lo1, up1), (p2, lo2, up2)) if one_larger else ((p2, lo2, up2), (p1, lo1, up1))
    avg = (p1 + p2) / 2.0
    if llo <= avg and sup > avg:
        boundary = avg
    elif llo <= avg:
        boundary = sup
    elif sup > avg:
        boundary = llo
    else:
        raise RuntimeError("Can't find boundary between cubes that are already disjoint!")
    if one_larger:
        return (boundary, up1), (lo2, boundary)
    return (lo1, boundary), (boundary, up2)


def make_disjoint(points: np.ndarray, lower: np.ndarray, upper: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """PointInCube.makeDisjoint(cubes) (:99-115): fold the cubes in one at a time; the new cube is made disjoint from every
    cube already accepted, in the order the accumulator holds them -- the accumulator is rebuilt as
    Vector(new, current) ++ tail at every step, so after folding cube c in the order is [c', the earlier cubes reversed...].
    The fold is followed literally, order included, because later splits depend on earlier ones."""
    P = len(points)
    lower, upper = lower.copy(), upper.copy()
    acc: List[int] = []                                       # indices, in the accumulator's order
    for c in range(P):
        work = [c]
        for cur in acc:                                       # disjointCubes.foldLeft(Vector(newCube))
            head = work[0]
            if _intersecting(lower[head], upper[head], lower[cur], upper[cur]):
                sep = (points[head] - points[cur]) ** 2       # dimensionOfLargestSeparation (:71-80): first maximum
                k = int(np.argmax(sep))
                (l1, u1), (l2, u2) = _find_disjoint_boundary(points[head, k], lower[head, k], upper[head, k],
                                                             points[cur, k], lower[cur, k], upper[cur, k])
                lower[head, k], upper[head, k] = l1, u1
                lower[cur, k], upper[cur, k] = l2, u2
            work = [head, cur] + work[1:]
        acc = work
    return lower, upper


def symmetrize(points, lower, upper):
    """PointInInterval.symmetrize (:51-61): shrink the longer side of every interval to the shorter one."""
    l1, l2 = upper - points, points - lower
    upper = np.where(l1 > l2, points + l2, upper)
    lower = np.where(l1 < l2, points - l1, lower)
    return lower, upper


def ready_for_sampling(points: np.ndarray):
    lower, upper = initialise_cover(points)
    lower, upper = make_disjoint(points, lower, upper)
    return symmetrize(points, lower, upper)


def cube_volumes(lower, upper) -> np.ndarray:
    """PointInCube.cubeVolume (:30-33): the SUM of the interval lengths (not the product) -- reproduced as written."""
    # summed left to right in Python floats (numpy's pairwise reduction would round differently from the device engine's loop)
    out = np.empty(len(lower))
    for i in range(len(lower)):
        v = 0.0
        for k in range(lower.shape[1]):
            v += abs(float(upper[i, k]) - float(lower[i, k]))
        out[i] = v
    return out


def choose_cube(volumes: np.ndarray, u: float) -> int:
    """MathOps.cdfStaircase over the volumes + the half-open lookup of PointInCubeCollection.getSample (:93-105)."""
    cdf = np.cumsum(volumes)                                   # sequential, like the device engine's loop
    return int(min(np.searchsorted(cdf / cdf[-1], u, side="right"), len(volumes) - 1))


# ---------------------------------------------------------------------------------------------- telemetry
@dataclass
class DomainTelemetry:
    """QuadratureDomainTelemetry with rescalingEnabled = false (its default, :35): the scale factor stays 1."""
    acceptances: int = 0
    rejections: int = 0
    acceptances_since_rebuild: int = 0
    rejection_streak: int = 0

    def add_acceptance(self):
        self.acceptances += 1
        self.acceptances_since_rebuild += 1
        self.rejection_streak = 0

    def add_rejection(self):
        self.rejections += 1
        self.rejection_streak += 1

    @property
    def initiate_rebuild(self) -> bool:                        # :48-49
        return self.rejection_streak >= 1000 and self.acceptances_since_rebuild >= 1

    @property
    def exhausted(self) -> bool:                               # :43
        return self.rejection_streak >= 100000

    def reset_for_rebuild(self):                               # :45-46 (the streak is NOT reset)
        self.acceptances = self.rejections = self.acceptances_since_rebuild = 0


# ---------------------------------------------------------------------------------------------- engine loop
def run_reference_slq(points: np.ndarray, logpdf: Callable[[np.ndarray], float],
                      uniform: Callable[[int, int], float], max_iterations: int, batch: int = 1,
                      rebuild_streak: int = 1000, exhausted_streak: int = 100000):
    """The sampling recursion of SlqEngine (:120-197, :216-247, :309-345) for a pool given as `points` (P, d), one proposal
    at a time: a proposal replaces the current minimum iff its logPdf is above it; the minimum is retired; the cover is
    rebuilt from the pool when the telemetry asks for it.  `uniform(stream, n)` is the n-th uniform of a named stream
    (0: cube choice, 1 + k: coordinate k of the proposal).  Proposals are drawn in batches of `batch` from the current
    cover (the reference keeps sampleParallelism fibres in flight against the same domain); what is left of a batch when the
    cover is rebuilt is discarded.  Returns (retired logPdfs in order, final pool points, final pool logPdfs, proposals).
    The convergence tests on the quadrature are the caller's business: this runs until max_iterations retirements or
    until the domain is exhausted."""
    points = np.array(points, dtype=np.float64)
    P, d = points.shape
    lps = np.array([logpdf(x) for x in points])
    lower, upper = ready_for_sampling(points)
    vol = cube_volumes(lower, upper)
    tel = DomainTelemetry()
    retired: List[float] = []
    n = 0                                                      # proposal counter = RNG index
    while len(retired) < max_iterations and tel.rejection_streak < exhausted_streak:
        props = []
        for b in range(batch):
            i = choose_cube(vol, uniform(0, n + b))
            x = np.array([(uniform(1 + k, n + b) - 0.5) * 1.0 * (upper[i, k] - lower[i, k]) + points[i, k] for k in range(d)])
            props.append(x)
        vals = [logpdf(x) for x in props]
        rebuilt = False
        for b in range(batch):
            n += 1
            if len(retired) >= max_iterations:
                break
            imin = int(np.argmin(lps))
            if lps[imin] < vals[b]:
                retired.append(float(lps[imin]))
                points[imin], lps[imin] = props[b], vals[b]
                tel.add_acceptance()
            else:
                tel.add_rejection()
            if tel.rejection_streak >= rebuild_streak and tel.acceptances_since_rebuild >= 1:
                lower, upper = ready_for_sampling(points)
                vol = cube_volumes(lower, upper)
                tel.reset_for_rebuild()
                rebuilt = True
            if rebuilt or tel.rejection_streak >= exhausted_streak:
                n += batch - 1 - b                             # the rest of the batch is discarded, its draws are skipped
                break
    return np.array(retired), points, lps, n
