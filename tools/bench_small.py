#!/usr/bin/env python
"""Focused timings (CUDA events, L2-rotating buffers): collapsed C2 model at 4096 / 65536 chains, C1-size and C2-size
HMC steps/s with the fused and the separate-kernel step.  Prints one JSON object."""
import json
import math
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import thylacine_b200 as t  # noqa: E402


def problem(d, n, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, d)) / np.sqrt(d)
    theta = rng.standard_normal(d)
    y = A @ theta + 0.1 * rng.standard_normal(n)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 20.0))],
                                   [t.GaussianLinearLikelihood.of(A, y, np.full(n, 0.2), "theta")])
    return post, theta


def time_eval(post, theta, B, steps=200):
    d = theta.size
    stream = torch.cuda.current_stream()
    post.setStream(stream.cuda_stream)
    nrot = max(3, int(140e6 // (2 * B * d * 8)) + 1)
    rng = np.random.default_rng(1)
    xs = [torch.from_numpy(theta[None, :] + 0.02 * rng.standard_normal((B, d))).cuda() for _ in range(nrot)]
    lps = [torch.empty(B, dtype=torch.float64, device="cuda") for _ in range(nrot)]
    gs = [torch.empty(B, d, dtype=torch.float64, device="cuda") for _ in range(nrot)]
    for i in range(5):
        post.logPdfGradBatchDevice(B, xs[i % nrot].data_ptr(), lps[i % nrot].data_ptr(), gs[i % nrot].data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        k = i % nrot
        post.logPdfGradBatchDevice(B, xs[k].data_ptr(), lps[k].data_ptr(), gs[k].data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"ms_per_step": ms, "evals_per_s": B / (ms * 1e-3), "path": post.lastKernelPath}


def time_hmc(post, theta, B, L, eps, fused, ntraj=20):
    stream = torch.cuda.current_stream()
    post.setStream(stream.cuda_stream)
    cfg = t.HmcmcConfig(1, L, 3, eps)
    hmc = t.HmcmcSampledPosterior(cfg, post, numberOfChains=B, rngSeed=2, seed={"theta": theta}, fusedSteps=fused)
    hmc.burnIn()
    hmc.runTrajectories(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    hmc.runTrajectories(ntraj)
    e1.record(stream)
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    att, acc, _ = hmc.stats()
    hmc.close()
    return {"steps_per_s": B * ntraj * L / sec, "us_per_leapfrog_step": 1e6 * sec / (ntraj * L),
            "acceptance": float(acc.sum() / att.sum()), "path": post.lastKernelPath}


def main():
    out = {"resident_warps": os.environ.get("THY_RESIDENT_WARPS", "12")}
    post, theta = problem(100, 10_000, 20260102)
    col = post.collapsed()
    for B in (4096, 16384, 65536):
        out[f"collapsed_c2_{B}"] = time_eval(col, theta, B)
        out[f"collapsed_c2_{B}"]["tflops_executed"] = 4.0 * 100 * 100 * out[f"collapsed_c2_{B}"]["evals_per_s"] / 1e12
    eps = 0.2 * 0.1 * math.sqrt(100 / 10_000)
    for fused in (True, False):
        out[f"hmc_c2_4096_fused{int(fused)}"] = time_hmc(post, theta, 4096, 16, eps, fused, ntraj=10)
        out[f"hmc_c2collapsed_4096_fused{int(fused)}"] = time_hmc(col, theta, 4096, 16, eps, fused, ntraj=50)
    post1, theta1 = problem(10, 1000, 20260101)
    eps1 = 0.2 * 0.1 * math.sqrt(10 / 1000)
    for fused in (True, False):
        out[f"hmc_c1_4096_fused{int(fused)}"] = time_hmc(post1, theta1, 4096, 16, eps1, fused, ntraj=50)
    out["c1_eval_4096"] = time_eval(post1, theta1, 4096)
    out["c1_eval_65536"] = time_eval(post1, theta1, 65536)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
