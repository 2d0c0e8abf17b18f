#!/usr/bin/env python
"""Sustained (>= 3 s back-to-back) fp64 figures with a clock record, next to the burst ones:
cuBLAS DGEMM 8192^3 (the roofline yardstick) and this library's config-2 evaluation (k_fused_dmma).  One JSON object."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402  (ClockSampler, synthetic_problem, chain_points)
import thylacine_b200 as t  # noqa: E402


def sustained(fn, flops_per_call, seconds=3.0):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    calls = 0
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(50):
            fn()
        calls += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    return {"calls": calls, "seconds": ms * 1e-3, "tflops": flops_per_call * calls / (ms * 1e-3) / 1e12, "clocks": clocks}


def burst(fn, flops_per_call, reps=10):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return {"ms": best, "tflops": flops_per_call / (best * 1e-3) / 1e12}


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty(n, n, dtype=torch.float64, device="cuda")
    out["cublas_dgemm_8192_burst"] = burst(lambda: torch.matmul(a, b, out=c), 2.0 * n ** 3)
    out["cublas_dgemm_8192_sustained"] = sustained(lambda: torch.matmul(a, b, out=c), 2.0 * n ** 3)
    del a, b, c
    prob = bench.synthetic_problem()
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                   [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")])
    post.setStream(torch.cuda.current_stream().cuda_stream)
    B, D = bench.CHAINS, bench.D
    xs = [torch.from_numpy(bench.chain_points(prob, B, seed=r)).cuda() for r in range(24)]
    lp = torch.empty(B, dtype=torch.float64, device="cuda")
    g = torch.empty(B, D, dtype=torch.float64, device="cuda")
    k = [0]

    def step():
        k[0] = (k[0] + 1) % 24
        post.logPdfGradBatchDevice(B, xs[k[0]].data_ptr(), lp.data_ptr(), g.data_ptr())

    out["c2_eval_burst"] = burst(step, bench.FLOPS_PER_EVAL * B)
    out["c2_eval_sustained"] = sustained(step, bench.FLOPS_PER_EVAL * B)
    out["c2_eval_sustained"]["evals_per_s"] = B * out["c2_eval_sustained"]["calls"] / out["c2_eval_sustained"]["seconds"]
    out["c2_sustained_fraction_of_sustained_dgemm"] = out["c2_eval_sustained"]["tflops"] / out["cublas_dgemm_8192_sustained"]["tflops"]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
