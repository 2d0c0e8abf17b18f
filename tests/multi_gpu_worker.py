"""Worker for tests/test_multi_gpu.py (launched by torchrun, one rank per GPU): SLQ with live points sharded over the
ranks (NCCL all-gather threshold selection) against the analytic evidence, and the cross-rank moment sum."""
import math
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import thylacine_b200 as t  # noqa: E402
from oracle import ref_oracle as o  # noqa: E402


def main():
    rank, world, local = t.distributed.rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = t.Communicator.fromTorchDistributed()
    assert (comm.rank, comm.world) == (rank, world)

    # --- moments: each rank holds its shard of the samples
    d, n = 9, 4000
    X = 2.0 + 0.3 * np.random.default_rng(1).standard_normal((n, d))
    lo, hi = t.shard_range(n, rank, world)
    post0 = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", np.zeros(d), np.ones(d))])
    mom = t.SampleMoments(post0)
    mom.add(torch.from_numpy(X[lo:hi]).cuda())
    cnt, mean, cov = mom.finish(comm)
    assert cnt == n
    assert np.max(np.abs(mean - X.mean(0))) < 1e-12 and np.max(np.abs(cov - np.cov(X.T))) < 1e-12

    # --- SLQ: identical thresholds and evidence on every rank
    d, P, K = 6, 8000, 400
    y = np.random.default_rng(11).standard_normal(d)
    post = t.UnnormalisedPosterior([t.UniformPrior.fromBounds("theta", np.full(d, 10.0), np.full(d, -10.0))],
                                   [t.GaussianLikelihood(t.LinearForwardModel.identityOf("theta", d), y, np.full(d, 2.0))])
    want = o.analytic_uniform_box_log_evidence(np.full(d, -10.0), np.full(d, 10.0), np.eye(d), y, np.ones(d))
    cfg = t.SlqConfig(poolSize=P, abscissaNumber=8, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                      sampleParallelism=K, maxIterationCount=40 * P, minIterationCount=P)
    slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=2026, comm=comm, constrainedWalkSteps=30)
    slq.rebuildSampleSimulation()
    H = slq.stats.neg_entropy_mean
    tol = 5.0 * math.sqrt(H / P) + 3.0 * slq.stats.log_evidence_std / math.sqrt(8)
    assert abs(slq.logEvidence - want) < tol, (slq.logEvidence, want, tol)
    ll = slq.retiredLogLikelihoods()
    assert np.all(np.diff(ll) >= 0)
    mine = torch.tensor([slq.logEvidence, float(slq.stats.iterations), float(ll.sum())], dtype=torch.float64, device="cuda")
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    for v in allv:
        assert torch.equal(v, allv[0]), "ranks disagree on the retired sequence / evidence"
    slq.close()

    # --- positive log-likelihoods (tight variances; ADVICE r01: a search over uninitialised candidate memory used to bias
    # the retire counts when logL > 0) and a pool that is not a multiple of the slice size; fixed round budget
    d, P, K = 4, 6000, 1500
    y = np.random.default_rng(12).standard_normal(d) * 0.01
    post = t.UnnormalisedPosterior([t.UniformPrior.fromBounds("theta", np.full(d, 1.0), np.full(d, -1.0))],
                                   [t.GaussianLikelihood(t.LinearForwardModel.identityOf("theta", d), y, np.full(d, 0.02))])
    want = o.analytic_uniform_box_log_evidence(np.full(d, -1.0), np.full(d, 1.0), np.eye(d), y, np.full(d, 1e-4))
    cfg = t.SlqConfig(poolSize=P, abscissaNumber=8, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                      sampleParallelism=K, maxIterationCount=60 * P, minIterationCount=P)
    slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=31, comm=comm, constrainedWalkSteps=30)
    slq.rebuildSampleSimulation()
    ll = slq.retiredLogLikelihoods()
    assert ll.max() > 0 and np.all(np.diff(ll) >= 0) and ll.size == slq.stats.total_retired
    assert slq.stats.iterations % min(K, P // world // 2) == 0 or slq.stats.iterations == cfg.maxIterationCount
    H = slq.stats.neg_entropy_mean
    tol2 = 5.0 * math.sqrt(H / P) + 3.0 * slq.stats.log_evidence_std / math.sqrt(8)
    assert abs(slq.logEvidence - want) < tol2, (slq.logEvidence, want, tol2)
    mine2 = torch.tensor([slq.logEvidence, float(slq.stats.iterations), float(ll.sum())], dtype=torch.float64, device="cuda")
    allv = [torch.empty_like(mine2) for _ in range(world)]
    dist.all_gather(allv, mine2)
    for v in allv:
        assert torch.equal(v, allv[0]), "ranks disagree on the retired sequence / evidence (positive logL case)"
    slq.close()
    comm.close()
    dist.destroy_process_group()
    if rank == 0:
        print(f"multi_gpu_worker ok: world={world} lnZ={mine[0].item():.4f} want={want:.4f} tol={tol:.3f}")


if __name__ == "__main__":
    main()
