#!/usr/bin/env python
"""A few evaluations of one named case, for `ncu` launch lists / full captures (see profiles/*.md).

    python tools/profile_case.py --case c2|c3|c5_analytic|c5_fd|c5_fd_literal|stream|slq|collapsed|collapsed4k|hmc_c2|hmc_c1|lfm|moments [--reps 3]
"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    # Watchdog: a stalled device call must not hold a GPU box for the rest of its lease.  After the limit the Python
    # stacks are dumped to stderr and the process exits (works even while the main thread sits inside a C call).
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("THY_WATCHDOG_SECONDS", "600")), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", required=True)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import thylacine_b200 as t
    rng = np.random.default_rng(0)

    def lin(d, n):
        A = rng.standard_normal((n, d)) / np.sqrt(d)
        theta = rng.standard_normal(d)
        return A, theta, A @ theta + 0.1 * rng.standard_normal(n)

    def run(post, X, grad=True):
        Xd = torch.from_numpy(X).cuda()
        lp = torch.empty(X.shape[0], dtype=torch.float64, device="cuda")
        g = torch.empty_like(Xd)
        for _ in range(args.reps):
            post.logPdfGradBatchDevice(X.shape[0], Xd.data_ptr(), lp.data_ptr(), g.data_ptr() if grad else None)
        post.synchronize()
        print(args.case, post.lastKernelPath, float(lp[0]))

    if args.case in ("c2", "stream"):
        d, n, B = (100, 10_000, 4096) if args.case == "c2" else (100, 1_000_000, 8)
        A, theta, y = lin(d, n)
        post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 20.0))],
                                       [t.GaussianLinearLikelihood.of(A, y, np.full(n, 0.2), "theta")])
        run(post, theta[None, :] + 0.02 * rng.standard_normal((B, d)))
    elif args.case == "c3":
        d, n, B = 1000, 100_000, 8192
        A, theta, y = lin(d, n)
        post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 20.0))],
                                       [t.GaussianLinearLikelihood.of(A, y, np.full(n, 0.2), "theta")])
        run(post, theta[None, :] + 0.02 * rng.standard_normal((B, d)), grad=False)
    elif args.case in ("c5_analytic", "c5_fd", "c5_fd_literal"):
        d, n, B = 200, 256, 16384
        A = rng.standard_normal((n, d)) / np.sqrt(d)
        theta = 0.3 * rng.standard_normal(d)
        y = np.tanh(A @ theta) + 0.05 * rng.standard_cauchy(n)
        fm = t.NonLinearForwardModel.of("tanh", {"theta": A}, differential=None if args.case == "c5_analytic" else 1e-6)
        post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 2.0))],
                                       [t.CauchyLikelihood(fm, y, np.full(n, 0.1))], fdIncremental=(args.case == "c5_fd"),
                                       cauchyReferenceSign=False)
        run(post, theta[None, :] + 0.05 * rng.standard_normal((B, d)))
    elif args.case in ("collapsed", "collapsed4k"):
        # config 2 collapsed onto d observations: the resident-A variant of the fused kernel (one launch per evaluation)
        d, n, B = 100, 10_000, (65536 if args.case == "collapsed" else 4096)
        A, theta, y = lin(d, n)
        post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 20.0))],
                                       [t.GaussianLinearLikelihood.of(A, y, np.full(n, 0.2), "theta")])
        col = post.collapsed()
        run(col, theta[None, :] + 0.02 * rng.standard_normal((B, d)))
    elif args.case in ("hmc_c2", "hmc_c1"):
        d, n, B = (100, 10_000, 4096) if args.case == "hmc_c2" else (10, 1000, 4096)
        A, theta, y = lin(d, n)
        post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 20.0))],
                                       [t.GaussianLinearLikelihood.of(A, y, np.full(n, 0.2), "theta")])
        cfg = t.HmcmcConfig(1, 4, 1, 0.2 * 0.1 * float(np.sqrt(d / n)))
        hmc = t.HmcmcSampledPosterior(cfg, post, numberOfChains=B, rngSeed=1, seed={"theta": theta}, useGraph=False)
        hmc.burnIn()
        hmc.runTrajectories(args.reps)
        post.synchronize()
        print(args.case, post.lastKernelPath)
    elif args.case == "lfm":
        d, n, P = 100, 2000, 4096
        A, theta, y = lin(d, n)
        post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 20.0))],
                                       [t.GaussianLinearLikelihood.of(A, y, np.full(n, 0.2), "theta")])
        lfm = t.LeapfrogMcmcSampledPosterior.of(t.LeapfrogMcmcConfig(1, 1, P), "euclidean", post, rngSeed=3)
        lfm.runSweeps(args.reps)
        post.synchronize()
        print("lfm", post.lastKernelPath)
    elif args.case == "moments":
        d, n = 100, 200_000
        post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.ones(d))])
        mom = t.SampleMoments(post)
        X = torch.from_numpy(rng.standard_normal((n, d))).cuda()
        for _ in range(args.reps):
            mom.add(X)
        print("moments", mom.finish()[0])
    elif args.case == "slq":
        d, P = 50, 1_000_000
        y = rng.standard_normal(d)
        post = t.UnnormalisedPosterior([t.UniformPrior.fromBounds("theta", np.full(d, 10.0), np.full(d, -10.0))],
                                       [t.GaussianLikelihood(t.LinearForwardModel.identityOf("theta", d), y, np.full(d, 2.0))])
        cfg = t.SlqConfig(P, 8, 0.05, 0.3, 65536, 2_000_000_000, P)
        slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=4)
        slq.runRounds(args.reps)
        post.synchronize()
        print("slq rounds done")
    else:
        raise SystemExit("unknown case")


if __name__ == "__main__":
    main()
