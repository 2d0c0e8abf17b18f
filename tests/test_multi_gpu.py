"""N > 1 on real GPUs (skipped on a single-GPU box): launches tests/multi_gpu_worker.py under torchrun."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2])
def test_slq_and_moments_across_ranks(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29531", str(ROOT / "tests" / "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "multi_gpu_worker ok" in r.stdout


def test_single_process_one_thread_per_gpu():
    """Single-process mode (thy_comm_create_all = ncclCommInitAll; thy_set_device binds a host thread to a GPU): ONE process,
    one thread per GPU -- how one JVM drives the box.  Both threads run the same sharded SLQ; they must agree bit for bit
    on the retired sequence and the evidence, and match the analytic value."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import math
    import threading
    import numpy as np
    import thylacine_b200 as t
    from thylacine_b200.distributed import deviceCount, setDevice
    from oracle import ref_oracle as o
    assert deviceCount() >= 2
    comms = t.Communicator.createAll([0, 1])
    d, P, K = 5, 6000, 500
    y = np.random.default_rng(3).standard_normal(d)
    want = o.analytic_uniform_box_log_evidence(np.full(d, -10.0), np.full(d, 10.0), np.eye(d), y, np.ones(d))
    out, err = [None, None], [None, None]

    def worker(i):
        try:
            setDevice(i)
            post = t.UnnormalisedPosterior([t.UniformPrior.fromBounds("theta", np.full(d, 10.0), np.full(d, -10.0))],
                                           [t.GaussianLikelihood(t.LinearForwardModel.identityOf("theta", d), y, np.full(d, 2.0))])
            cfg = t.SlqConfig(poolSize=P, abscissaNumber=4, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                              sampleParallelism=K, maxIterationCount=40 * P, minIterationCount=P)
            slq = t.SlqIntegratedPosterior.of(cfg, post, rngSeed=9, comm=comms[i], constrainedWalkSteps=25)
            slq.rebuildSampleSimulation()
            out[i] = (slq.logEvidence, slq.stats.neg_entropy_mean, slq.stats.log_evidence_std, int(slq.stats.iterations),
                      slq.retiredLogLikelihoods())
            slq.close()
            post.close()
        except Exception as e:  # surfaced by the main thread
            err[i] = e

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=600)
    assert err == [None, None], err
    assert out[0][0] == out[1][0] and out[0][3] == out[1][3] and np.array_equal(out[0][4], out[1][4])
    tol = 5.0 * math.sqrt(out[0][1] / P) + 3.0 * out[0][2] / 2.0
    assert abs(out[0][0] - want) < tol, (out[0][0], want, tol)
    for c in comms:
        c.close()
