#!/usr/bin/env python
"""Runs BASELINE.json's configs 1, 3, 4, 5 (and the HMC variant of config 2) on one GPU (or one rank's share of the
multi-GPU configs) and records throughput + canonical work rates.  bench.py covers config 2 itself.

    python tools/run_configs.py [--only c1,c2hmc,c3,c4,c5] [--out gpurun_out/configs.json]

Work accounting follows SURVEY.md section 8(d): 4 n d flops per logpost+grad eval, 2 n d per value.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def linear_problem(d, n, seed, sigma=0.1):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, d)) / np.sqrt(d)
    theta = rng.standard_normal(d)
    y = A @ theta + sigma * rng.standard_normal(n)
    return A, theta, y


def timed(fn, sync):
    sync()
    t0 = time.perf_counter()
    fn()
    sync()
    return time.perf_counter() - t0


def c1(t, torch, out):
    """HmcmcSampledPosterior, 10-d, 1k observations, single chain; CPU leg = the oracle's literal HmcmcEngine."""
    from oracle import ref_oracle as o
    d, n = 10, 1000
    A, theta, y = linear_problem(d, n, 20260101)
    ci, pm, pci = np.full(n, 0.2), np.zeros(d), np.full(d, 20.0)
    eps = 0.2 * 0.1 * math.sqrt(d / n)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", pm, pci)],
                                   [t.GaussianLinearLikelihood.of(A, y, ci, "theta")])
    stream = torch.cuda.Stream()
    post.setStream(stream.cuda_stream)
    res = {}
    for B in (1, 4096):
        cfg = t.HmcmcConfig(stepsBetweenSamples=1, stepsInDynamicsSimulation=16, warmupStepCount=200,
                            dynamicsSimulationStepSize=eps)
        hmc = t.HmcmcSampledPosterior(cfg, post, numberOfChains=B, rngSeed=1, seed={"theta": theta})
        hmc.burnIn()
        ntraj = 2000 if B == 1 else 500
        dt = timed(lambda: hmc.runTrajectories(ntraj), post.synchronize)
        att, acc, _ = hmc.stats()
        res[f"gpu_B{B}"] = {"leapfrog_steps_per_s": B * ntraj * 16 / dt, "trajectories": ntraj, "seconds": dt,
                            "acceptance": float(acc.sum() / att.sum())}
        hmc.close()
    # CPU: the reference's sequential algorithm with its dense n x n covariance (oracle, literal)
    ref = o.Posterior([o.gaussian_prior_from_ci("theta", pm, pci)], [o.gaussian_linear_likelihood(A, y, ci, "theta")])
    eng = o.HmcmcEngine(ref, o.HmcmcConfig(1, 16, 2, eps), seed_point=theta, rng=np.random.default_rng(0))
    t0 = time.perf_counter()
    eng.sample(3)
    dt = time.perf_counter() - t0
    res["cpu_literal_reference"] = {"leapfrog_steps_per_s": eng.jump_attempts * 16 / dt, "trajectories": eng.jump_attempts,
                                    "seconds": dt, "cores": os.cpu_count()}
    out["c1_hmc_10d_1k"] = res
    post.close()


def c2hmc(t, torch, out):
    """HMC leapfrog steps/s at the config-2 shape (100-d, 10k observations, 4096 chains)."""
    d, n, B, L = 100, 10_000, 4096, 16
    A, theta, y = linear_problem(d, n, 20260102)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 20.0))],
                                   [t.GaussianLinearLikelihood.of(A, y, np.full(n, 0.2), "theta")])
    stream = torch.cuda.Stream()
    post.setStream(stream.cuda_stream)
    cfg = t.HmcmcConfig(1, L, 10, 0.2 * 0.1 * math.sqrt(d / n))
    hmc = t.HmcmcSampledPosterior(cfg, post, numberOfChains=B, rngSeed=2, seed={"theta": theta})
    hmc.burnIn()
    ntraj = 40
    dt = timed(lambda: hmc.runTrajectories(ntraj), post.synchronize)
    att, acc, _ = hmc.stats()
    steps = B * ntraj * L
    out["c2_hmc_100d_10k_4096"] = {"leapfrog_steps_per_s": steps / dt, "seconds": dt, "acceptance": float(acc.sum() / att.sum()),
                                   "tflops_canonical": steps * 4.0 * n * d / dt / 1e12,
                                   "note": "one leapfrog step = one gradient = 4 n d flops; + one value per trajectory"}
    hmc.close()
    post.close()


def c3(t, torch, out, pool=8192, d=1000, n=100_000):
    """LeapfrogMcmcSampledPosterior, 1000-d, 100k observations; one GPU's share (8192) of the 65 536-member pool."""
    A, theta, y = linear_problem(d, n, 20260103)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 20.0))],
                                   [t.GaussianLinearLikelihood.of(A, y, np.full(n, 0.2), "theta")])
    stream = torch.cuda.Stream()
    post.setStream(stream.cuda_stream)
    cfg = t.LeapfrogMcmcConfig(stepsBetweenSamples=10, warmupStepCount=1, samplePoolSize=pool)
    t0 = time.perf_counter()
    lfm = t.LeapfrogMcmcSampledPosterior.of(cfg, "euclidean", post, rngSeed=3)
    t_create = time.perf_counter() - t0
    sweeps = 3
    dt = timed(lambda: lfm.runSweeps(sweeps), post.synchronize)
    prop, acc = lfm.stats()
    out["c3_leapfrog_mcmc_1000d_100k"] = {
        "pool_per_gpu": pool, "proposals_per_s": pool * sweeps / dt, "seconds_per_sweep": dt / sweeps,
        "create_seconds_incl_pool_eval_and_1_warmup_sweep": t_create, "acceptance": acc / max(prop, 1),
        "tflops_canonical": pool * sweeps * (2.0 * n * d + 3.0 * (pool / 2) * d) / dt / 1e12,
        "kernel_path": post.lastKernelPath}
    lfm.close()
    post.close()


def c4(t, torch, out, P=1_000_000, d=50):
    """SlqIntegratedPosterior, 50-d Gaussian posterior: throughput at 10^6 live points + evidence accuracy at 10^5."""
    from oracle import ref_oracle as o
    y = np.random.default_rng(20260104).standard_normal(d)

    def model(uniform):
        pr = (t.UniformPrior.fromBounds("theta", np.full(d, 10.0), np.full(d, -10.0)) if uniform
              else t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 10.0)))
        return t.UnnormalisedPosterior([pr], [t.GaussianLikelihood(t.LinearForwardModel.identityOf("theta", d), y, np.full(d, 2.0))])

    res = {}
    post = model(True)
    K = 65536
    cfg = t.SlqConfig(poolSize=P, abscissaNumber=8, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                      sampleParallelism=K, maxIterationCount=2_000_000_000, minIterationCount=P)
    slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=4)
    slq.runRounds(3)
    rounds = 30
    dt = timed(lambda: slq.runRounds(rounds), post.synchronize)
    res["throughput_P1e6"] = {"rounds_per_s": rounds / dt, "retired_per_s": rounds * K / dt,
                              "likelihood_evals_per_s": rounds * K * 20 / dt, "K": K, "walk_steps": 20}
    slq.close()
    post.close()
    for uniform in (True, False):
        Pa = 100_000
        post = model(uniform)
        want = (o.analytic_uniform_box_log_evidence(np.full(d, -10.0), np.full(d, 10.0), np.eye(d), y, np.ones(d)) if uniform
                else o.analytic_gaussian_log_evidence(np.zeros(d), np.full(d, 25.0), np.eye(d), y, np.ones(d)))
        cfg = t.SlqConfig(poolSize=Pa, abscissaNumber=8, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                          sampleParallelism=Pa // 16, maxIterationCount=int(1.6 * 80 * Pa), minIterationCount=Pa)
        slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=5, constrainedWalkSteps=int(os.environ.get("SLQ_WALK", "40")))
        t0 = time.perf_counter()
        slq.rebuildSampleSimulation()
        dt = time.perf_counter() - t0
        st = slq.stats
        res["evidence_P1e5_" + ("uniform" if uniform else "gaussian")] = {
            "log_evidence": st.log_evidence_mean, "analytic": want, "error": st.log_evidence_mean - want,
            "std_over_abscissas": st.log_evidence_std, "H": st.neg_entropy_mean,
            "tolerance_5sqrtHoverP": 5 * math.sqrt(max(st.neg_entropy_mean, 0) / Pa), "iterations": int(st.iterations),
            "rounds": int(st.rounds), "seconds": dt, "walk_acceptance": st.walk_acceptance, "walk_scale": st.walk_scale,
            "likelihood_evals": int(st.likelihood_evaluations)}
        slq.close()
        post.close()
    out["c4_slq_50d"] = res


def c5(t, torch, out, d=200, n=256, B=16384, L=8):
    """Non-linear forward model tanh(A theta), Cauchy likelihood, FD Jacobian, HMC over 16384 chains."""
    rng = np.random.default_rng(20260105)
    A = rng.standard_normal((n, d)) / np.sqrt(d)
    theta = 0.3 * rng.standard_normal(d)
    y = np.tanh(A @ theta) + 0.05 * rng.standard_cauchy(n)
    res = {}
    for name, kw in (("fd_literal", dict(differential=1e-6, fdInc=False)), ("fd_incremental", dict(differential=1e-6, fdInc=True)),
                     ("analytic_jacobian", dict(differential=None, fdInc=False))):
        fm = t.NonLinearForwardModel.of("tanh", {"theta": A}, differential=kw["differential"])
        post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 2.0))],
                                       [t.CauchyLikelihood(fm, y, np.full(n, 0.1))], fdIncremental=kw["fdInc"])
        stream = torch.cuda.Stream()
        post.setStream(stream.cuda_stream)
        cfg = t.HmcmcConfig(1, L, 2, 0.01)
        hmc = t.HmcmcSampledPosterior(cfg, post, numberOfChains=B, rngSeed=6, seed={"theta": theta})
        hmc.burnIn()
        ntraj = 3
        dt = timed(lambda: hmc.runTrajectories(ntraj), post.synchronize)
        att, acc, _ = hmc.stats()
        steps = B * ntraj * L
        res[name] = {"leapfrog_steps_per_s": steps / dt, "seconds": dt, "acceptance": float(acc.sum() / att.sum()),
                     "kernel_path": post.lastKernelPath}
        hmc.close()
        post.close()
    out["c5_nonlinear_cauchy_fd_200d"] = res


def stream(t, torch, out, d=100, n=1_000_000):
    """Small-batch streaming kernel against the HBM roofline: A (8 n d bytes = 800 MB >> L2) read once per launch."""
    import ctypes as C
    A, theta, y = linear_problem(d, n, 20260106)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 20.0))],
                                   [t.GaussianLinearLikelihood.of(A, y, np.full(n, 0.2), "theta")])
    stream_ = torch.cuda.Stream()
    post.setStream(stream_.cuda_stream)
    L = t.lib()
    res = {}
    for B in (1, 2, 4, 8, 16):
        X = torch.from_numpy(theta[None, :] + 0.01 * np.random.default_rng(B).standard_normal((B, d))).cuda()
        lp = torch.empty(B, dtype=torch.float64, device="cuda")
        g = torch.empty(B, d, dtype=torch.float64, device="cuda")
        for _ in range(3):
            post.logPdfGradBatchDevice(B, X.data_ptr(), lp.data_ptr(), g.data_ptr())
        post.synchronize()
        L.thy_model_enable_timing(post.handle, 1)
        reps = 20
        dt = timed(lambda: [post.logPdfGradBatchDevice(B, X.data_ptr(), lp.data_ptr(), g.data_ptr()) for _ in range(reps)],
                   post.synchronize)
        k_ms, k_n = C.c_double(), C.c_int64()
        L.thy_model_kernel_time(post.handle, C.byref(k_ms), C.byref(k_n))
        L.thy_model_enable_timing(post.handle, 0)
        kms = k_ms.value / max(k_n.value, 1)
        res[f"B{B}"] = {"kernel_path": post.lastKernelPath, "kernel_ms": kms, "wall_ms_per_eval_batch": 1e3 * dt / reps,
                        "algorithmic_GBps": 8.0 * n * d / (kms * 1e-3) / 1e9, "evals_per_s": B * reps / dt}
    out["stream_100d_1M_obs"] = res
    post.close()


def main():
    # Watchdog: a stalled device call must not hold a GPU box for the rest of its lease.  After the limit the Python
    # stacks are dumped to stderr and the process exits (works even while the main thread sits inside a C call).
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("THY_WATCHDOG_SECONDS", "600")), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c1,c2hmc,c3,c4,c5")
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "configs.json"))
    args = ap.parse_args()
    import torch
    import thylacine_b200 as t
    out = {"gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count()}
    fns = {"c1": c1, "c2hmc": c2hmc, "c3": c3, "c4": c4, "c5": c5, "stream": stream}
    for name in args.only.split(","):
        t0 = time.perf_counter()
        try:
            fns[name](t, torch, out)
        except Exception as e:  # record and keep going: one config must not hide the others
            out[name + "_error"] = repr(e)
        out.setdefault("wall_seconds", {})[name] = time.perf_counter() - t0
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(out, indent=1))
        print(name, "done in", round(out["wall_seconds"][name], 1), "s", flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
