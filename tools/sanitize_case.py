#!/usr/bin/env python
"""Small fixtures for `compute-sanitizer --tool memcheck|synccheck|racecheck python tools/sanitize_case.py [cases]`.

No torch import (the sanitizer slows process start-up enough already): everything goes through the C ABI with host
buffers.  Each case is checked against the oracle so a sanitizer-clean run is also a correct one.
Cases: stream (1 chain + DMMA row groups, ring wraps), fused, generic, hmc, slq, collapse."""
from __future__ import annotations

import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import thylacine_b200 as t  # noqa: E402
from thylacine_b200 import capi  # noqa: E402
from oracle import ref_oracle as o  # noqa: E402
from helpers import linear_gaussian_problem, rel_err_grad, rel_err_logp  # noqa: E402


def linear_case(d, n, B, path, tag):
    prob = linear_gaussian_problem(d, n, seed=d + n)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                   [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")], kernelPath=path)
    X = prob["theta"][None, :] + 0.05 * np.random.default_rng(B).standard_normal((B, d))
    want_lp, want_g = o.linear_gaussian_batch(prob["A"], prob["y"], prob["var"], X, prior_mean=prob["pm"], prior_var=prob["pvar"])
    lp, g = post.logPdfGradBatch(X)
    e1, e2 = rel_err_logp(lp, want_lp), rel_err_grad(g, want_g)
    print(f"{tag}: d={d} n={n} B={B} path={post.lastKernelPath} err {e1:.1e} {e2:.1e}", flush=True)
    assert e1 < 1e-10 and e2 < 1e-10
    post.close()


def main():
    cases = sys.argv[1:] or ["stream", "fused", "generic", "hmc", "slq", "collapse"]
    if "stream" in cases:
        linear_case(20, 3000, 1, capi.KERNEL_STREAM, "stream/1-chain")
        linear_case(5, 40_000, 3, capi.KERNEL_STREAM, "stream/row-group, ring wraps")      # 5000 8-row blocks / 148 CTAs > 32 stages
        linear_case(100, 30_000, 8, capi.KERNEL_STREAM, "stream/row-group, 100-d, ring wraps")
        linear_case(33, 257, 5, capi.KERNEL_STREAM, "stream/row-group ragged")
    if "fused" in cases:
        linear_case(100, 1000, 96, capi.KERNEL_FUSED_DMMA, "fused")
        linear_case(12, 200, 200, capi.KERNEL_FUSED_DMMA, "fused small")
    if "generic" in cases:
        linear_case(12, 200, 40, capi.KERNEL_GENERIC, "generic")
    if "collapse" in cases:
        prob = linear_gaussian_problem(40, 600, seed=4)
        post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                       [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")])
        col = post.collapsed()
        X = prob["theta"][None, :] + 0.02 * np.random.default_rng(1).standard_normal((300, 40))
        want_lp, want_g = o.linear_gaussian_batch(prob["A"], prob["y"], prob["var"], X, prior_mean=prob["pm"], prior_var=prob["pvar"])
        lp, g = col.logPdfGradBatch(X)
        print(f"collapse: path={col.lastKernelPath} err {rel_err_logp(lp, want_lp):.1e} {rel_err_grad(g, want_g):.1e}", flush=True)
        assert rel_err_logp(lp, want_lp) < 1e-9 and rel_err_grad(g, want_g) < 1e-9
        col.close()
        post.close()
    if "hmc" in cases:
        prob = linear_gaussian_problem(10, 300, seed=2)
        post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                       [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")])
        cfg = t.HmcmcConfig(stepsBetweenSamples=1, stepsInDynamicsSimulation=4, warmupStepCount=3,
                            dynamicsSimulationStepSize=0.2 * 0.1 * math.sqrt(10 / 300))
        hmc = t.HmcmcSampledPosterior(cfg, post, numberOfChains=70, rngSeed=1, seed={"theta": prob["theta"]})
        s = hmc.sampleChains(3)
        att, acc, _ = hmc.stats()
        print(f"hmc: samples {s.shape}, acceptance {acc.sum() / att.sum():.2f}", flush=True)
        assert np.all(np.isfinite(s))
        hmc.close()
        post.close()
    if "slq" in cases:
        d = 4
        y = np.random.default_rng(11).standard_normal(d)
        post = t.UnnormalisedPosterior([t.UniformPrior.fromBounds("theta", np.full(d, 10.0), np.full(d, -10.0))],
                                       [t.GaussianLikelihood(t.LinearForwardModel.identityOf("theta", d), y, np.full(d, 2.0))])
        want = o.analytic_uniform_box_log_evidence(np.full(d, -10.0), np.full(d, 10.0), np.eye(d), y, np.ones(d))
        P = 3000
        cfg = t.SlqConfig(poolSize=P, abscissaNumber=3, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                          sampleParallelism=200, maxIterationCount=30 * P, minIterationCount=P)
        for fused in (True, False):
            slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=5, constrainedWalkSteps=10, fusedWalk=fused)
            slq.rebuildSampleSimulation()
            print(f"slq fused={fused}: lnZ {slq.logEvidence:.3f} (analytic {want:.3f}), rounds {slq.stats.rounds}", flush=True)
            assert abs(slq.logEvidence - want) < 1.0
            slq.close()
        post.close()
    print("sanitize_case ok:", " ".join(cases), flush=True)


if __name__ == "__main__":
    main()
