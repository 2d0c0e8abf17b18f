#!/usr/bin/env python
"""SLQ round-time probe (single process or torchrun): rounds/s of the config-4 model for a given pool, with the round
replayed from a CUDA graph or launched plainly (THY_SLQ_GRAPH=0), plus the phase trace.  Prints one JSON line per rank 0."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import thylacine_b200 as t  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pool", type=int, default=1_000_000)
    ap.add_argument("--k", type=int, default=62_500)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--rounds", type=int, default=300)
    ap.add_argument("--d", type=int, default=50)
    ap.add_argument("--blocks", type=int, default=1, help="time this many consecutive blocks of --rounds rounds")
    ap.add_argument("--max-iterations", type=int, default=2_000_000_000)
    ap.add_argument("--full-run", action="store_true", help="time rebuildSampleSimulation instead of blocks of rounds")
    ap.add_argument("--warm", type=int, default=0, help="with --full-run: rounds run (and synchronised) before the timed call")
    ap.add_argument("--import-oracle", action="store_true")
    args = ap.parse_args()
    rank, world, local = t.distributed.rank_world()
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = t.Communicator.fromTorchDistributed()
    d = args.d
    y = np.random.default_rng(20260104).standard_normal(d)
    post = t.UnnormalisedPosterior([t.UniformPrior.fromBounds("theta", np.full(d, 10.0), np.full(d, -10.0))],
                                   [t.GaussianLikelihood(t.LinearForwardModel.identityOf("theta", d), y, np.full(d, 2.0))])
    stream = torch.cuda.current_stream()
    post.setStream(stream.cuda_stream)
    cfg = t.SlqConfig(poolSize=args.pool, abscissaNumber=8, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                      sampleParallelism=args.k, maxIterationCount=args.max_iterations,
                      minIterationCount=min(args.max_iterations, 2_000_000_000))
    slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=6, comm=comm, constrainedWalkSteps=args.steps)
    if args.import_oracle:
        from oracle import ref_oracle as o  # noqa: F401
    if args.full_run:
        if args.warm:
            slq.runRounds(args.warm)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        slq.rebuildSampleSimulation()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            print(json.dumps({"world": world, "full_run_seconds": dt, "rounds": int(slq.stats.rounds),
                              "us_per_round": 1e6 * dt / slq.stats.rounds}), flush=True)
        slq.close()
        post.close()
        if world > 1:
            dist.destroy_process_group()
        return
    slq.runRounds(5)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    per_block = []
    for _ in range(args.blocks):
        t0 = time.perf_counter()
        slq.runRounds(args.rounds)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        per_block.append(1e6 * dt / args.rounds)
    slq.setTrace(True)
    slq.runRounds(10)
    torch.cuda.synchronize()
    ph = slq.phaseTimes()
    ph["replacements_mean_max"] = slq.loadBalance()
    if rank == 0:
        print(json.dumps({"world": world, "pool": args.pool, "k": args.k, "walk_steps": args.steps,
                          "graph_env": os.environ.get("THY_SLQ_GRAPH", "1"), "us_per_round": per_block[0], "us_per_round_blocks": per_block,
                          "phases": ph}), flush=True)
    slq.close()
    post.close()
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
