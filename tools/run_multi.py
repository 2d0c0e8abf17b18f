#!/usr/bin/env python
"""BASELINE.json's multi-GPU configs under torchrun (one rank per GPU):

  config 3  LeapfrogMcmcSampledPosterior, 1000-d, 100k observations, 65 536 pool members = WORLD x sub-pools, no collective
  config 4  SlqIntegratedPosterior, 50-d, 10^6 live points sharded over the ranks, one NCCL all-gather per round,
            throughput at 10^6 live points + log-evidence against the analytic value at 10^5 live points

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
        tools/run_multi.py [--only c3,c4] [--out gpurun_out/multi_N.json]

Times are device-synchronised wall times bracketed by a barrier, max over ranks; rank 0 writes the JSON.
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
sys.path.insert(0, str(ROOT / "tools"))


def main():
    # Watchdog: a stalled device call must not hold a GPU box for the rest of its lease.  After the limit the Python
    # stacks are dumped to stderr and the process exits (works even while the main thread sits inside a C call).
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("THY_WATCHDOG_SECONDS", "600")), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c3,c4")
    ap.add_argument("--out", default=None)
    ap.add_argument("--c3-total-pool", type=int, default=65536)
    ap.add_argument("--c3-sweeps", type=int, default=2)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import thylacine_b200 as t
    from run_configs import linear_problem

    rank, world, local = t.distributed.rank_world()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = t.Communicator.fromTorchDistributed() if world > 1 else None

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        v = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        v = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
        return float(v.item())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = {"gpu": torch.cuda.get_device_name(local), "world": world, "host_cores": os.cpu_count()}
    todo = args.only.split(",")

    if "c3" in todo:
        d, n = 1000, 100_000
        lo, hi = t.shard_range(args.c3_total_pool, rank, world)
        pool = hi - lo
        A, theta, y = linear_problem(d, n, 20260103)
        post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 20.0))],
                                       [t.GaussianLinearLikelihood.of(A, y, np.full(n, 0.2), "theta")])
        cfg = t.LeapfrogMcmcConfig(stepsBetweenSamples=10, warmupStepCount=1, samplePoolSize=pool)
        lfm = t.LeapfrogMcmcSampledPosterior.of(cfg, "euclidean", post, rngSeed=3 + rank)   # one independent sub-pool per rank
        # Start the pool in the typical set (theta* + posterior-width noise, sd ~ sigma sqrt(d/n) = 0.01), as a user would
        # after an optimiser, instead of the sigma = 10 prior.  The acceptance still reads 0.0 at d = 1000, and that is the
        # reference's algorithm, not this implementation: the proposal reflects x_i through a partner, x' = 2 x_j - x_i
        # (LeapfrogMcmcEngine.scala:104-121), so for a pool distributed like the target x' - mu = 2 e_j - e_i has FIVE times
        # the target's variance and log p(x') - log p(x_i) ~ -2 d = -2000.  (tests/test_lfm_gpu.py: ~8 % acceptance at d = 3.)
        # The figure below is therefore proposal evaluations per second.
        lfm.setPool(theta[None, :] + 0.01 * np.random.default_rng(100 + rank).standard_normal((pool, d)))
        lfm.runSweeps(1)
        barrier()
        t0 = time.perf_counter()
        lfm.runSweeps(args.c3_sweeps)
        post.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        prop, acc = lfm.stats()
        total_pool = int(sum_over_ranks(pool))
        out["c3_leapfrog_mcmc_1000d_100k"] = {
            "total_pool": total_pool, "pool_per_gpu": pool, "sweeps": args.c3_sweeps,
            "proposals_per_s": total_pool * args.c3_sweeps / dt, "seconds_per_sweep": dt / args.c3_sweeps,
            "acceptance_rank0": acc / max(prop, 1),
            "tflops_canonical_all_gpus": total_pool * args.c3_sweeps * (2.0 * n * d + 3.0 * (pool / 2) * d) / dt / 1e12,
            "kernel_path": post.lastKernelPath, "collectives": "none"}
        lfm.close()
        post.close()

    if "c4" in todo or "c4full" in todo:
        from oracle import ref_oracle as o
        d = 50
        y = np.random.default_rng(20260104).standard_normal(d)

        def model(uniform):
            pr = (t.UniformPrior.fromBounds("theta", np.full(d, 10.0), np.full(d, -10.0)) if uniform
                  else t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.full(d, 10.0)))
            return t.UnnormalisedPosterior([pr], [t.GaussianLikelihood(t.LinearForwardModel.identityOf("theta", d), y,
                                                                      np.full(d, 2.0))])
        res = {}
        P, K = 1_000_000, 65536
        post = model(True)
        cfg = t.SlqConfig(poolSize=P, abscissaNumber=8, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                          sampleParallelism=K, maxIterationCount=2_000_000_000, minIterationCount=P)
        slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=4, comm=comm)
        slq.runRounds(3)
        rounds = 30
        barrier()
        t0 = time.perf_counter()
        slq.runRounds(rounds)
        post.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        res["throughput_P1e6"] = {"rounds_per_s": rounds / dt, "retired_per_s": rounds * K / dt,
                                  "likelihood_evals_per_s": rounds * K * 20 / dt, "K": K, "walk_steps": 20,
                                  "collective": "one NCCL all-gather of each rank's K smallest per round" if world > 1 else "none"}
        slq.close()
        post.close()
        for uniform in (True, False):
            Pa = 100_000
            post = model(uniform)
            want = (o.analytic_uniform_box_log_evidence(np.full(d, -10.0), np.full(d, 10.0), np.eye(d), y, np.ones(d))
                    if uniform else o.analytic_gaussian_log_evidence(np.zeros(d), np.full(d, 25.0), np.eye(d), y, np.ones(d)))
            cfg = t.SlqConfig(poolSize=Pa, abscissaNumber=8, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                              sampleParallelism=Pa // 16, maxIterationCount=int(1.6 * 80 * Pa), minIterationCount=Pa)
            slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=5, comm=comm, constrainedWalkSteps=40)
            barrier()
            t0 = time.perf_counter()
            slq.rebuildSampleSimulation()
            dt = max_over_ranks(time.perf_counter() - t0)
            st = slq.stats
            res["evidence_P1e5_" + ("uniform" if uniform else "gaussian")] = {
                "log_evidence": st.log_evidence_mean, "analytic": want, "error": st.log_evidence_mean - want,
                "std_over_abscissas": st.log_evidence_std, "H": st.neg_entropy_mean,
                "tolerance_5sqrtHoverP": 5 * math.sqrt(max(st.neg_entropy_mean, 0) / Pa),
                "iterations": int(st.iterations), "rounds": int(st.rounds), "seconds": dt}
            slq.close()
            post.close()
        if "c4full" in todo:
            # BASELINE config 4 as written: 10^6 live points, run to the engine's own termination, against the analytic value
            for uniform in (True, False):
                post = model(uniform)
                want = (o.analytic_uniform_box_log_evidence(np.full(d, -10.0), np.full(d, 10.0), np.eye(d), y, np.ones(d))
                        if uniform else o.analytic_gaussian_log_evidence(np.zeros(d), np.full(d, 25.0), np.eye(d), y, np.ones(d)))
                cfg = t.SlqConfig(poolSize=P, abscissaNumber=8, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                                  sampleParallelism=P // 16, maxIterationCount=int(1.6 * 80 * P), minIterationCount=P)
                slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=6, comm=comm, constrainedWalkSteps=40)
                barrier()
                t0 = time.perf_counter()
                slq.rebuildSampleSimulation()
                dt = max_over_ranks(time.perf_counter() - t0)
                st = slq.stats
                res["evidence_P1e6_" + ("uniform" if uniform else "gaussian")] = {
                    "log_evidence": st.log_evidence_mean, "analytic": want, "error": st.log_evidence_mean - want,
                    "std_over_abscissas": st.log_evidence_std, "H": st.neg_entropy_mean,
                    "tolerance_5sqrtHoverP": 5 * math.sqrt(max(st.neg_entropy_mean, 0) / P),
                    "iterations": int(st.iterations), "rounds": int(st.rounds), "seconds": dt,
                    "likelihood_evals": int(st.likelihood_evaluations)}
                slq.close()
                post.close()
        out["c4_slq_50d"] = res

    if comm is not None:
        comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        text = json.dumps(out, indent=1)
        if args.out:
            Path(args.out).parent.mkdir(parents=True, exist_ok=True)
            Path(args.out).write_text(text)
        print(text)


if __name__ == "__main__":
    main()
