#!/usr/bin/env python
"""Benchmark of the posterior-evaluation hot path (BASELINE.json metric, config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the hot path over one batch: batched fp64 log-posterior + gradient of a 100-d
linear-Gaussian model with 10k observations for 4096 chains per GPU (weak scaling: chains shard across ranks,
no data-path collective).  Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

D, N_OBS, CHAINS = 100, 10_000, 4096
METRIC = "fp64 logpost+grad evals/sec"
UNIT = "evals/s"
WORKLOAD = "batched log-posterior+gradient, 100-d linear-Gaussian, 10k observations, 4096 chains per GPU (BASELINE configs[1])"
FLOPS_PER_EVAL = 4.0 * N_OBS * D          # forward 2nd + adjoint 2nd (canonical form, as executed)
ROTATE = 24                               # distinct input/output buffer sets cycled through the timed region


def synthetic_problem(seed=20260102):
    """SURVEY 8(d) recipe for C2: A_ij ~ N(0,1)/sqrt(d), y = A theta* + 0.1 eps, likelihood CI 0.2, prior N(0, 10^2)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((N_OBS, D)) / np.sqrt(D)
    theta = rng.standard_normal(D)
    y = A @ theta + 0.1 * rng.standard_normal(N_OBS)
    return dict(A=A, theta=theta, y=y, ci=np.full(N_OBS, 0.2), pm=np.zeros(D), pci=np.full(D, 20.0))


def chain_points(prob, B, seed):
    """Typical-set points: theta* + O(sigma_post) noise (posterior sd ~ sigma sqrt(d/n) = 0.01)."""
    rng = np.random.default_rng(seed)
    return prob["theta"][None, :] + 0.02 * rng.standard_normal((B, D))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for k, name in enumerate(names):
            if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU legs are meant to use every host thread."""
    cores = os.cpu_count() or 1
    try:
        from threadpoolctl import ThreadpoolController
        global _THREADPOOL
        _THREADPOOL = ThreadpoolController()
        _THREADPOOL.limit(limits=cores)       # stays in force for the life of the process
    except Exception:
        pass
    return cores


# ------------------------------------------------------------------------------------------- reference arm
def literal_reference_evals(prob, points, budget_s):
    """The reference's algorithm executed literally (oracle port): dense n x n covariance, LU solve per logPdf
    (GaussianDistribution.scala:52-61 -> Breeze `covariance \\ centered`), dense inverse mat-vec per gradient (:63-68),
    one point per call.  Returns (evals done, seconds)."""
    from oracle import ref_oracle as o
    post = o.Posterior([o.gaussian_prior_from_ci("theta", prob["pm"], prob["pci"])],
                       [o.gaussian_linear_likelihood(prob["A"], prob["y"], prob["ci"], "theta")])
    t0 = time.perf_counter()
    done = 0
    for x in points:
        post.log_pdf(x)
        post.log_pdf_gradient(x)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return done, time.perf_counter() - t0, post


def structured_port_evals_per_s(prob, X, min_seconds):
    """Structure-aware CPU port (diagonal covariance, batched A@X / W@A through OpenBLAS on all cores)."""
    from oracle import ref_oracle as o
    var, pvar = o.ci_to_variance(prob["ci"]), o.ci_to_variance(prob["pci"])
    o.linear_gaussian_batch(prob["A"], prob["y"], var, X[:64], prior_mean=prob["pm"], prior_var=pvar)
    t0 = time.perf_counter()
    evals = 0
    while True:
        o.linear_gaussian_batch(prob["A"], prob["y"], var, X, prior_mean=prob["pm"], prior_var=pvar)
        evals += X.shape[0]
        if time.perf_counter() - t0 >= min_seconds:
            break
    return evals / (time.perf_counter() - t0), evals


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = use_all_host_cores()
    prob = synthetic_problem()
    pts = chain_points(prob, args.steps + args.warmup, seed=7)
    # warm-up builds the lazily cached Cholesky / inverse exactly as the reference's lazy vals would
    from oracle import ref_oracle as o
    post = o.Posterior([o.gaussian_prior_from_ci("theta", prob["pm"], prob["pci"])],
                       [o.gaussian_linear_likelihood(prob["A"], prob["y"], prob["ci"], "theta")])
    t_w = time.perf_counter()
    for x in pts[:max(1, args.warmup)]:
        post.log_pdf(x)
        post.log_pdf_gradient(x)
        if time.perf_counter() - t_w > 120:
            break
    t0 = time.perf_counter()
    done = 0
    for x in pts[args.warmup:]:
        post.log_pdf(x)
        post.log_pdf_gradient(x)
        done += 1
        if time.perf_counter() - t0 > 150:   # bounded: the literal algorithm costs ~0.7 TFLOP per logPdf at n = 10^4
            break
    dt = time.perf_counter() - t0
    value = done / dt
    structured, _ = structured_port_evals_per_s(prob, chain_points(prob, 1024, seed=8), 5.0)
    sample = (f"{done} sequential logPdfAt+logPdfGradientAt calls on the C2 model, executed literally as the reference does "
              f"(dense {N_OBS}x{N_OBS} covariance, LU solve per logPdf, dense inverse mat-vec per gradient), numpy/OpenBLAS on all cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(done, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "d": D, "n_obs": N_OBS, "step": "one logpost+grad evaluation (bounded sample of the workload)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "structured_port_evals_per_s": structured,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def load_fp64_peak():
    """fp64 tensor/vector peak: MEASURED_PEAKS.json has no fp64 entry, so the denominator is this repo's own
    measurement of cuBLAS DGEMM 8192^3 on the pool's B200 (profiles/fp64_peaks.json), else the datasheet figure."""
    p = ROOT / "profiles" / "fp64_peaks.json"
    if p.exists():
        try:
            j = json.loads(p.read_text())
            return float(j["cublas_dgemm_8192_tflops"]), "measured: cuBLAS DGEMM 8192^3 on this pool (profiles/fp64_peaks.json)"
        except Exception:
            pass
    return 37.0, "fallback: HGX B200 datasheet fp64 (37 TFLOP/s); MEASURED_PEAKS.json has no fp64 entry"


def run_slq_leg(t, torch, dist, rank, world, stream, barrier):
    """BASELINE configs[3] at full size.  Returns the dict that goes under the bench line's "slq" key (rank 0), timed on
    the device (CUDA events on the model's stream), max over ranks."""
    import math
    from oracle import ref_oracle as o     # analytic evidence only (checker, not on the timed path)
    d, P, A = 50, 1_000_000, 8
    K, M = P // 16, 40
    y = np.random.default_rng(20260104).standard_normal(d)
    want = o.analytic_uniform_box_log_evidence(np.full(d, -10.0), np.full(d, 10.0), np.eye(d), y, np.ones(d))
    comm = t.Communicator.fromTorchDistributed() if world > 1 else None

    def build(seed):
        post = t.UnnormalisedPosterior([t.UniformPrior.fromBounds("theta", np.full(d, 10.0), np.full(d, -10.0))],
                                       [t.GaussianLikelihood(t.LinearForwardModel.identityOf("theta", d), y, np.full(d, 2.0))])
        post.setStream(stream.cuda_stream)
        cfg = t.SlqConfig(poolSize=P, abscissaNumber=A, domainScalingIncrement=0.05, targetAcceptanceProbability=0.3,
                          sampleParallelism=K, maxIterationCount=int(1.6 * 80 * P), minIterationCount=P)
        return post, t.SlqIntegratedPosterior.of(cfg, post, rngSeed=seed, comm=comm, constrainedWalkSteps=M)

    def max_over_ranks(v):
        x = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(x, op=dist.ReduceOp.MAX)
        return float(x.item())

    # (1) the whole simulation: live phase to the engine's own termination + drain, against the analytic evidence
    post, slq = build(6)
    launches0 = t.kernelLaunchCount()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    slq.rebuildSampleSimulation()
    e1.record(stream)
    barrier()
    sec = max_over_ranks(e0.elapsed_time(e1)) * 1e-3
    st = slq.stats
    launches = t.kernelLaunchCount() - launches0
    lb_mean, lb_max = slq.loadBalance()
    lb_max_all = max_over_ranks(lb_max)
    H = st.neg_entropy_mean
    tol = 5.0 * math.sqrt(max(H, 0.0) / P) + 3.0 * st.log_evidence_std / math.sqrt(A)
    out = {
        "workload": "SlqIntegratedPosterior log-evidence, 50-d Gaussian likelihood x uniform box prior, 10^6 live points "
                    "(BASELINE configs[3]), K = P/16 retired per round, 40 constrained walk steps per replacement, 8 abscissas",
        "scaling": "strong", "n_gpus": world, "live_points_per_gpu": P // world,
        "seconds": sec, "rounds": int(st.rounds), "rounds_per_s": st.rounds / sec, "retired_per_s": st.iterations / sec,
        "likelihood_evals_per_s": st.likelihood_evaluations / sec, "iterations": int(st.iterations),
        "log_evidence": st.log_evidence_mean, "analytic_log_evidence": want, "error": st.log_evidence_mean - want,
        "tolerance": tol, "within_tolerance": bool(abs(st.log_evidence_mean - want) < tol),
        "neg_entropy": H, "std_over_abscissas": st.log_evidence_std, "walk_acceptance": st.walk_acceptance,
        "gpu_launches": int(launches),
        "replacements_per_round_per_rank": {"ideal": K / world, "rank0_mean": lb_mean, "max_over_ranks_and_rounds": lb_max_all},
        "collective": ("one NCCL all-gather of every rank's log-likelihoods per round on the product communicator "
                       "(thy_comm_create, not torch.distributed)" if world > 1 else "none (single rank)"),
    }
    slq.close()
    post.close()
    # (2) where a round's time goes: a second instance, rounds as plain launches with CUDA events between the phases
    post, slq = build(7)
    slq.runRounds(3)
    slq.setTrace(True)
    slq.runRounds(12)
    post.synchronize()
    ph = slq.phaseTimes()
    slq.setTrace(False)
    out["phase_us_per_round"] = {k: (max_over_ranks(v) if k != "rounds" else v) for k, v in ph.items()}
    mode = int(ph.get("graph_replay_bit0_peer_memory_bit1", 0))
    out["round_replayed_from_cuda_graph"] = bool(mode & 1)
    if world > 1:
        out["collective"] = ("none per round: new log-likelihoods are written into every rank's memory over NVLink peer memory by "
                             "the walk kernel (slq_p2p.cu); NCCL on the product communicator (thy_comm_create) for set-up and "
                             "the final drain" if (mode & 2) else out["collective"])
    slq.close()
    post.close()
    if comm is not None:
        comm.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import thylacine_b200 as t
    from thylacine_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (thylacine_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    if args.only_slq:
        def _barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
        out = run_slq_leg(t, torch, dist, rank, world, torch.cuda.current_stream(), _barrier)
        if rank == 0:
            print(json.dumps({"slq": out}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    prob = synthetic_problem()
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                   [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")])
    stream = torch.cuda.current_stream()
    post.setStream(stream.cuda_stream)
    L = t.lib()

    # inputs resident in HBM: ROTATE distinct batches (footprint > L2), each rank its own chains
    Xs, LPs, Gs = [], [], []
    for r in range(ROTATE):
        X = chain_points(prob, CHAINS, seed=1000 * rank + r)
        Xs.append(torch.from_numpy(X).cuda())
        LPs.append(torch.empty(CHAINS, dtype=torch.float64, device="cuda"))
        Gs.append(torch.empty(CHAINS, D, dtype=torch.float64, device="cuda"))

    def step(i):
        k = i % ROTATE
        post.logPdfGradBatchDevice(CHAINS, Xs[k].data_ptr(), LPs[k].data_ptr(), Gs[k].data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    # quick correctness guard on the warm-up output (oracle as checker only)
    if rank == 0:
        from oracle import ref_oracle as o
        k = (args.warmup - 1) % ROTATE if args.warmup > 0 else 0
        if args.warmup == 0:
            step(0)
            torch.cuda.synchronize()
        want_lp, want_g = o.linear_gaussian_batch(prob["A"], prob["y"], o.ci_to_variance(prob["ci"]), Xs[k][:256].cpu().numpy(),
                                                  prior_mean=prob["pm"], prior_var=o.ci_to_variance(prob["pci"]))
        got_lp, got_g = LPs[k][:256].cpu().numpy(), Gs[k][:256].cpu().numpy()
        err_lp = float(np.max(np.abs(got_lp - want_lp) / np.maximum(1, np.abs(want_lp))))
        err_g = float(np.max(np.abs(got_g - want_g)) / np.max(np.abs(want_g)))
        if not (err_lp < 1e-10 and err_g < 1e-10):
            raise SystemExit(f"bench.py: parity check failed (logp {err_lp:.2e}, grad {err_g:.2e})")

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = t.kernelLaunchCount()
    L.thy_model_enable_timing(post.handle, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = t.kernelLaunchCount() - launches0
    import ctypes as C
    k_ms, k_n = C.c_double(), C.c_int64()
    capi.check(L.thy_model_kernel_time(post.handle, C.byref(k_ms), C.byref(k_n)))
    L.thy_model_enable_timing(post.handle, 0)
    clocks = sampler.stop() if rank == 0 else None

    tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_max = float(tmax.item())
    value = world * CHAINS * args.steps / (ms_max * 1e-3)

    # ---- e2e: the public host-buffer call (pinned host memory -> H2D -> kernels -> D2H of logp and grad), every step's
    # copies inside the timed region.  The caller keeps TWO calls in flight (thy_logpdf_grad_batch_async + thy_model_wait,
    # the library alternates two staging sets): step k+1's input copy and step k-1's output copy run under step k's
    # kernels, and the host takes delivery of step k-1's results before it queues step k+1.
    NBUF = 3
    Xh = [torch.from_numpy(chain_points(prob, CHAINS, seed=5000 + 100 * rank + r)).pin_memory() for r in range(NBUF)]
    LPh = [torch.empty(CHAINS, dtype=torch.float64).pin_memory() for _ in range(NBUF)]
    Gh = [torch.empty(CHAINS, D, dtype=torch.float64).pin_memory() for _ in range(NBUF)]
    import ctypes
    dp = ctypes.POINTER(ctypes.c_double)

    def e2e_run(nsteps):
        tickets = []
        for i in range(nsteps):
            k = i % NBUF
            tk = ctypes.c_int64()
            capi.check(L.thy_logpdf_grad_batch_async(post.handle, CHAINS, ctypes.cast(Xh[k].data_ptr(), dp),
                                                     ctypes.cast(LPh[k].data_ptr(), dp), ctypes.cast(Gh[k].data_ptr(), dp),
                                                     ctypes.byref(tk)))
            tickets.append(tk.value)
            if i >= 1:
                capi.check(L.thy_model_wait(post.handle, tickets[i - 1]))   # results of the previous step are on the host
        capi.check(L.thy_model_wait(post.handle, 0))

    e2e_run(max(3, args.warmup))
    # delivered results must be the synchronous call's results
    chk_lp, chk_g = torch.empty(CHAINS, dtype=torch.float64).pin_memory(), torch.empty(CHAINS, D, dtype=torch.float64).pin_memory()
    k_last = (max(3, args.warmup) - 1) % NBUF
    capi.check(L.thy_logpdf_grad_batch(post.handle, CHAINS, ctypes.cast(Xh[k_last].data_ptr(), dp),
                                       ctypes.cast(chk_lp.data_ptr(), dp), ctypes.cast(chk_g.data_ptr(), dp)))
    # (the blocking call evaluates in three chunks, so its stream-K partial sums combine in another order: rounding only)
    d_lp = float(((chk_lp - LPh[k_last]).abs() / chk_lp.abs().clamp(min=1.0)).max())
    d_g = float(((chk_g - Gh[k_last]).abs().amax(1) / chk_g.abs().amax(1)).max())
    if not (d_lp < 1e-13 and d_g < 1e-12):
        raise SystemExit(f"bench.py: pipelined host-buffer call disagrees with the synchronous call ({d_lp:.1e}, {d_g:.1e})")
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * CHAINS * args.steps / float(te.item())
    # the single blocking call (latency of ONE evaluation from host memory, chunked [small, large, small])
    t0 = time.perf_counter()
    nsync = max(5, min(50, args.steps))
    for i in range(nsync):
        capi.check(L.thy_logpdf_grad_batch(post.handle, CHAINS, ctypes.cast(Xh[i % NBUF].data_ptr(), dp),
                                           ctypes.cast(LPh[i % NBUF].data_ptr(), dp), ctypes.cast(Gh[i % NBUF].data_ptr(), dp)))
    e2e_blocking_value = CHAINS * nsync / (time.perf_counter() - t0)

    # ---- side by side: the reduced-flop form.  The same posterior with its linear-Gaussian term
    # collapsed onto d observations (thy_model_collapse); flops are counted as executed (4 d^2 per eval), never mixed
    # with the canonical count.  Reported next to, not instead of, `value`.
    reduced = None
    try:
        col = post.collapsed()
        col.setStream(stream.cuda_stream)
        red = {}
        for chains in (CHAINS, 16 * CHAINS):
            nrot = max(3, int(140e6 // (2 * chains * D * 8)) + 1)
            if chains == CHAINS:
                xs, lps, gs = Xs[:nrot], LPs[:nrot], Gs[:nrot]
                ref_lp, ref_g = LPs[0].clone(), Gs[0].clone()      # canonical output for Xs[0] (written in the timed loop)
                post.logPdfGradBatchDevice(CHAINS, Xs[0].data_ptr(), ref_lp.data_ptr(), ref_g.data_ptr())
            else:
                xs = [torch.cat([Xs[(r + j) % ROTATE] for j in range(16)]) for r in range(nrot)]
                lps = [torch.empty(chains, dtype=torch.float64, device="cuda") for _ in range(nrot)]
                gs = [torch.empty(chains, D, dtype=torch.float64, device="cuda") for _ in range(nrot)]
            nsteps = max(20, min(args.steps, 200))
            for i in range(3):
                col.logPdfGradBatchDevice(chains, xs[i % nrot].data_ptr(), lps[i % nrot].data_ptr(), gs[i % nrot].data_ptr())
            barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            for i in range(nsteps):
                k = i % nrot
                col.logPdfGradBatchDevice(chains, xs[k].data_ptr(), lps[k].data_ptr(), gs[k].data_ptr())
            r1.record(stream)
            barrier()
            rt = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(rt, op=dist.ReduceOp.MAX)
            rms = float(rt.item()) / nsteps
            entry = {"value": world * chains / (rms * 1e-3), "unit": UNIT, "ms_per_step": rms, "steps": nsteps,
                     "tflops_executed_per_gpu": 4.0 * D * D * chains / (rms * 1e-3) / 1e12}
            if chains == CHAINS:
                col.logPdfGradBatchDevice(CHAINS, Xs[0].data_ptr(), lps[0].data_ptr(), gs[0].data_ptr())
                torch.cuda.synchronize()
                entry["max_rel_diff_vs_canonical"] = {
                    "logp": float(((lps[0] - ref_lp).abs() / ref_lp.abs().clamp(min=1.0)).max().item()),
                    "grad": float(((gs[0] - ref_g).abs().amax(1) / ref_g.abs().amax(1)).max().item())}
            red[f"chains_per_gpu_{chains}"] = entry
        reduced = {"form": "linear-Gaussian term collapsed onto d observations: A' = chol(A^T Sigma^-1 A)^T built on the device "
                           "once; same log-posterior and gradient (tests/test_collapse_gpu.py, 1e-10)",
                   "flops_per_eval_executed": 4.0 * D * D, "flops_per_eval_canonical": FLOPS_PER_EVAL,
                   "kernel_path": col.lastKernelPath, **red}
        col.close()
    except Exception as e:  # the canonical line must not be lost to the side-by-side leg
        reduced = {"error": repr(e)}

    # ---- the second half of BASELINE.json's metric: HMC leapfrog steps/s on the same model (one leapfrog step = one
    # position update of one chain = one gradient evaluation); chains resident on the device
    hmc_line = None
    try:
        import math
        L_STEPS = 16
        sigma = 0.1
        eps = 0.2 * sigma * math.sqrt(D / N_OBS)
        cfg = t.HmcmcConfig(stepsBetweenSamples=1, stepsInDynamicsSimulation=L_STEPS, warmupStepCount=5,
                            dynamicsSimulationStepSize=eps)
        hmc = t.HmcmcSampledPosterior(cfg, post, numberOfChains=CHAINS, rngSeed=2 + rank, seed={"theta": prob["theta"]})
        hmc.burnIn()
        ntraj = max(5, min(40, args.steps // 8))
        hmc.runTrajectories(3)
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record(stream)
        hmc.runTrajectories(ntraj)
        h1.record(stream)
        barrier()
        ht = torch.tensor([h0.elapsed_time(h1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ht, op=dist.ReduceOp.MAX)
        hsec = float(ht.item()) * 1e-3
        att, acc, _ = hmc.stats()
        steps_total = world * CHAINS * ntraj * L_STEPS
        hmc_line = {"metric": "HMC leapfrog steps/sec", "value": steps_total / hsec, "unit": "steps/s",
                    "chains_per_gpu": CHAINS, "trajectories": ntraj, "leapfrog_steps_per_trajectory": L_STEPS,
                    "step_size": eps, "acceptance_rank0": float(acc.sum() / max(att.sum(), 1)),
                    "tflops_canonical_per_gpu": CHAINS * ntraj * L_STEPS * FLOPS_PER_EVAL / hsec / 1e12,
                    "kernel_path": post.lastKernelPath,
                    "note": "one CUDA-graph launch per trajectory; leapfrog step = fused DMMA gradient + elementwise kick/drift kernels "
                            "(at this size the library's cost model keeps the step out of the gradient kernel's output step: "
                            "558 vs 547 us measured; short models -- config 1, collapsed -- run one launch per step)"}
        hmc.close()
    except Exception as e:  # the evals/s line must not be lost to the HMC leg
        hmc_line = {"error": repr(e)}

    # ---- the one path with a collective (BASELINE configs[3]): SLQ log-evidence of a 50-d Gaussian posterior with 10^6 live
    # points SHARDED over the ranks through the product's own NCCL communicator (thy_comm_create): one all-gather of the
    # ranks' log-likelihoods per round, radix select of the K lowest on every rank.  Strong scaling: the pool is fixed.
    slq_line = None
    if not args.no_slq:
        try:
            slq_line = run_slq_leg(t, torch, dist, rank, world, stream, barrier)
        except Exception as e:  # the evals/s line must not be lost to the SLQ leg
            slq_line = {"error": repr(e)}

    if rank == 0:
        peak, peak_src = load_fp64_peak()
        k_avg_ms = k_ms.value / max(k_n.value, 1)
        achieved = FLOPS_PER_EVAL * CHAINS / (k_avg_ms * 1e-3) / 1e12 if k_n.value else None
        traffic = None   # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
        for tp, get in ((ROOT / "profiles" / "r02_ncu_fused_c2.json",
                         lambda j: j["launches"][0]["dram_read_bytes"] + j["launches"][0]["dram_write_bytes"]),
                        (ROOT / "profiles" / "ncu_fused_summary.json", lambda j: j.get("dram_bytes_per_launch"))):
            if tp.exists():
                try:
                    traffic = get(json.loads(tp.read_text()))
                    break
                except Exception:
                    traffic = None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = use_all_host_cores()
            v, evals = structured_port_evals_per_s(prob, chain_points(prob, CHAINS, seed=9), 10.0)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{evals} evals of the same C2 model in batches of {CHAINS}: oracle/ref_oracle.linear_gaussian_batch "
                             f"(diagonal-covariance restatement, numpy/OpenBLAS threads = all cores), >= 10 s of CPU work; the literal "
                             f"dense-covariance reference algorithm is timed by --impl reference"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "d": D, "n_obs": N_OBS, "chains_per_gpu": CHAINS, "kernel_path": post.lastKernelPath,
                       "l2": f"inputs/outputs rotate over {ROTATE} buffer sets ({ROTATE * 2 * CHAINS * D * 8 / 1e6:.0f} MB > 126 MB L2); "
                             f"A (8.6 MB) is a model constant reused by every chain tile within a step",
                       "parallelism": f"chains sharded over {world} GPU(s), no data-path collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": CHAINS * D * 8,
                    "d2h_bytes_per_step": CHAINS * (D + 1) * 8,
                    "call": "thy_logpdf_grad_batch_async + thy_model_wait, two calls in flight, pinned host buffers",
                    "blocking_call_value_per_gpu": e2e_blocking_value},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "kernel": "k_fused_dmma", "kernel_ms": k_avg_ms, "kernel_launches": int(k_n.value),
                         "flops_per_launch": FLOPS_PER_EVAL * CHAINS, "peak_source": peak_src},
            "cpu_baseline": cpu,
            "reduced_flop": reduced,
            "hmc": hmc_line,
            "slq": slq_line,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    # Watchdog: a stalled device call must not hold a GPU box for the rest of its lease.  After the limit the Python
    # stacks are dumped to stderr and the process exits (works even while the main thread sits inside a C call).
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("THY_WATCHDOG_SECONDS", "1500")), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-slq", action="store_true")
    ap.add_argument("--only-slq", action="store_true", help="run just the SLQ leg and print its dict (diagnostics)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
