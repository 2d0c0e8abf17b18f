"""TEST INFRASTRUCTURE ONLY (like ref_oracle.py): host restatements of the BATCHED forms of the reference's engines.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package; the product never does
(tests/test_abi.py::test_product_never_imports_oracle).

The reference's Leapfrog-MCMC proposes strictly one point at a time against a P x P distance table
(/root/reference/src/main/scala/thylacine/model/sampling/leapfrog/LeapfrogMcmcEngine.scala:88-162); the device runs
half-pool sweeps in which every member of the active half is updated in parallel against the frozen other half
(DESIGN.md section 5).  `leapfrog_mcmc_half_pool_sweep` states THAT algorithm, step by step in the reference's own
terms, so that the device can be replayed draw for draw:
  partner choice   MathOps.vectorCdfStaircase over the distances to the frozen half   (util/MathOps.scala:76-91)
  proposal         reflection through the partner, x' = 2 x_j - x_i                  (LeapfrogMcmcEngine.scala:104-121)
  Hastings ratio   exp(logPdf(x') - logPdf(x_i)) * q(x' -> x_i) / q(x_i -> x')        (:123-146)
It is parity-pinned only against the reference's sequential engine statistically (tests/test_lfm_gpu.py): the reference
holds no test for this engine."""
from __future__ import annotations

import math
from typing import Callable

import numpy as np


def leapfrog_mcmc_half_pool_sweep(x: np.ndarray, lp: np.ndarray, logpdf: Callable[[np.ndarray], float],
                                  partner_uniform: Callable[[int], float], accept_uniform: Callable[[int], float],
                                  dist: Callable[[np.ndarray, np.ndarray], float]) -> None:
    """One sweep = one proposal for every pool member, two half-pool passes; x (P, d) and lp (P,) are updated in place.
    partner_uniform(i) / accept_uniform(i): the uniforms member i consumes in this sweep (the device's counter RNG)."""
    P = len(x)
    n0 = P // 2
    for half in (0, 1):
        active = list(range(0, n0)) if half == 0 else list(range(n0, P))
        frozen = list(range(n0, P)) if half == 0 else list(range(0, n0))
        XS = x[frozen].copy()
        updates = []
        for i in active:
            D = np.array([dist(x[i], s) for s in XS])
            total = D.sum()
            cdf = np.cumsum(D)
            j = min(int(np.searchsorted(cdf, partner_uniform(i) * total, side="right")), len(frozen) - 1)
            q_forward = D[j] / total
            proposal = 2.0 * XS[j] - x[i]
            D2 = np.array([dist(proposal, s) for s in XS])
            q_reverse = D2[j] / D2.sum()
            lp_new = logpdf(proposal)
            a = math.exp(min(lp_new - lp[i], 700)) * q_reverse / q_forward
            if a >= 1 or accept_uniform(i) < a:
                updates.append((i, proposal, lp_new))
        for i, proposal, lp_new in updates:
            x[i] = proposal
            lp[i] = lp_new
