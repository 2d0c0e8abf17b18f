"""HMC on the device vs the reference's algorithm (oracle HmcmcEngine) and vs the analytic Gaussian posterior.

Reference: /root/reference/src/main/scala/thylacine/model/sampling/hmcmc/HmcmcEngine.scala:55-214.
The reference's RNG is unseeded, so sample-level parity is defined by REPLAY: the oracle runs the reference's
sequential single-chain algorithm with the very normals / uniforms the device's counter-based generator produced.
"""
import numpy as np
import pytest

import thylacine_b200 as t
from oracle import ref_oracle as o
from helpers import linear_gaussian_problem
import philox_ref as ph

pytestmark = pytest.mark.gpu


def _posteriors(prob):
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                   [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")])
    ref = o.Posterior([o.gaussian_prior_from_ci("theta", prob["pm"], prob["pci"])],
                      [o.gaussian_linear_likelihood(prob["A"], prob["y"], prob["ci"], "theta")])
    return post, ref


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("use_graph", [False, True])
def test_hmc_replay_matches_reference_algorithm(use_graph, fused):
    import torch
    d, n = 6, 80
    prob = linear_gaussian_problem(d, n, seed=42)
    post, ref = _posteriors(prob)
    stream = torch.cuda.Stream()
    post.setStream(stream.cuda_stream)
    cfg = t.HmcmcConfig(stepsBetweenSamples=2, stepsInDynamicsSimulation=5, warmupStepCount=7,
                        dynamicsSimulationStepSize=0.02)
    B, rng_seed, n_samples = 5, 1234567, 4
    mean, _ = o.analytic_gaussian_posterior(prob["pm"], prob["pvar"], prob["A"], prob["y"], prob["var"])
    seeds = mean[None, :] + 0.05 * np.random.default_rng(0).standard_normal((B, d))
    hmc = t.HmcmcSampledPosterior(cfg, post, seed=seeds, numberOfChains=B, rngSeed=rng_seed, useGraph=use_graph,
                                  fusedSteps=fused)   # fused: one launch per leapfrog step (kernels_fused_dmma.cu emit step)
    got = hmc.sampleChains(n_samples)
    att, acc, dh = hmc.stats()
    for b in range(B):
        eng = o.HmcmcEngine(ref, o.HmcmcConfig(2, 5, 7, 0.02), seed_point=seeds[b],
                            momentum_fn=lambda tr, dd, b=b: np.array([ph.normal(rng_seed, b, j, ph.RNG_HMC_MOMENTUM, tr) for j in range(dd)]),
                            uniform_fn=lambda tr, b=b: ph.uniform2(rng_seed, b, 0, ph.RNG_HMC_ACCEPT, tr)[0])
        want = np.array(eng.sample(n_samples))
        # chains advance in lock-step: chain b has run as many trajectories as the slowest chain needed
        assert np.max(np.abs(got[b] - want)) < 1e-9 * max(1.0, np.max(np.abs(want)))
        assert acc[b] >= eng.jump_acceptances and att[b] >= eng.jump_attempts
    assert np.all(att == att[0])


def test_hmc_moments_match_analytic_posterior():
    """C1-like model (10-d, 1k observations), many chains: sample mean / variance vs the analytic Gaussian posterior
    within 5 sigma / sqrt(ESS) (SURVEY 8c tolerances; ESS taken as chains x samples / 2 for thinned HMC)."""
    import torch
    d, n = 10, 1000
    prob = linear_gaussian_problem(d, n, seed=20260101)
    post, _ = _posteriors(prob)
    post.setStream(torch.cuda.Stream().cuda_stream)
    eps = 0.2 * 0.1 * np.sqrt(d / n)           # SURVEY 8(d) C1 step size: 0.2 sigma sqrt(d/n) = 0.002
    # trajectory length L eps = 0.016 ~ a quarter period of the posterior oscillator (omega ~ 100): near-independent
    # successive samples; (L = 16 at this step size is a full period and returns close to the start)
    cfg = t.HmcmcConfig(stepsBetweenSamples=1, stepsInDynamicsSimulation=4, warmupStepCount=50,
                        dynamicsSimulationStepSize=eps * 2)
    B, S = 1024, 8
    mean, cov = o.analytic_gaussian_posterior(prob["pm"], prob["pvar"], prob["A"], prob["y"], prob["var"])
    sd = np.sqrt(np.diag(cov))
    # over-dispersed starts around the mode (as after an optimiser), not the sigma_p = 10 prior: burn-in from the
    # prior is a property of the problem (energy 5e5 per coordinate), not of the implementation under test
    seeds = mean[None, :] + 3.0 * sd[None, :] * np.random.default_rng(1).standard_normal((B, d))
    hmc = t.HmcmcSampledPosterior(cfg, post, seed=seeds, numberOfChains=B, rngSeed=99)
    samples = hmc.sampleChains(S).reshape(B * S, d)
    ess = B * S / 2.0
    att, acc, _ = hmc.stats()
    assert 0.5 < acc.sum() / att.sum() <= 1.0
    assert np.max(np.abs(samples.mean(0) - mean) / sd) < 5.0 / np.sqrt(ess)
    assert np.max(np.abs(samples.var(0) / sd ** 2 - 1.0)) < 5.0 * np.sqrt(2.0 / ess)
    # off-diagonal structure: correlation matrix within 5/sqrt(ESS)
    corr = np.corrcoef(samples.T)
    want = cov / np.outer(sd, sd)
    assert np.max(np.abs(corr - want)) < 6.0 / np.sqrt(ess)


def test_hmc_single_chain_reference_surface():
    d, n = 4, 50
    prob = linear_gaussian_problem(d, n, seed=3)
    post, _ = _posteriors(prob)
    cfg = t.HmcmcConfig(1, 8, 20, 0.02)
    hmc = t.HmcmcSampledPosterior(cfg, post, seed={"theta": prob["theta"]})
    out = hmc.sample(5)
    assert len(out) == 5 and all(set(s) == {"theta"} and s["theta"].shape == (d,) for s in out)
    assert len({s["theta"].tobytes() for s in out}) == 5            # Set semantics: all distinct
    att, acc, dh = hmc.stats()
    assert att[0] >= (20 + 1) + 5 * 2 and np.isfinite(dh[0])


def test_hmc_uses_correct_cauchy_sign():
    """A sampler on a Cauchy model must climb towards the mode even when the model reproduces the reference's
    wrong-signed gradient for logPdfGradientAt."""
    rng = np.random.default_rng(1)
    d, n = 3, 40
    A = rng.standard_normal((n, d))
    theta = np.array([0.5, -1.0, 2.0])
    y = A @ theta + 0.05 * rng.standard_normal(n)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("c", np.zeros(d), np.full(d, 20.0))],
                                   [t.CauchyLikelihood.of(A, y, np.full(n, 0.1), "c")], cauchyReferenceSign=True)
    cfg = t.HmcmcConfig(1, 10, 300, 0.004)
    hmc = t.HmcmcSampledPosterior(cfg, post, seed=np.tile(theta + 0.3, (64, 1)), numberOfChains=64, rngSeed=5)
    s = hmc.sampleChains(4).reshape(-1, d)
    assert np.max(np.abs(s.mean(0) - theta)) < 0.05


def test_hmc_graph_survives_workspace_growth_on_the_same_posterior():
    """A captured trajectory graph holds the addresses of the model's workspace buffers.  An unrelated, larger
    evaluation on the same posterior re-allocates them; the sampler must notice (workspace epoch) and re-capture
    instead of replaying stale pointers.  Same seed with and without the interleaved evaluation => identical chains."""
    d, n, B = 16, 400, 256
    prob = linear_gaussian_problem(d, n, seed=9)
    cfg = t.HmcmcConfig(stepsBetweenSamples=1, stepsInDynamicsSimulation=6, warmupStepCount=3, dynamicsSimulationStepSize=0.01)
    states = []
    for disturb in (False, True):
        post, _ = _posteriors(prob)
        hmc = t.HmcmcSampledPosterior(cfg, post, numberOfChains=B, rngSeed=11, seed={"theta": prob["theta"]})
        hmc.burnIn()
        hmc.runTrajectories(3)                                   # graph captured and replayed
        if disturb:
            big = np.random.default_rng(1).standard_normal((40 * B, d))
            post.logPdfGradBatch(big)                            # grows X / logp / grad staging and the ticket array
        hmc.runTrajectories(4)
        states.append(hmc.state())
        hmc.close()
        post.close()
    assert np.array_equal(states[0][0], states[1][0]) and np.array_equal(states[0][1], states[1][1])


@pytest.mark.parametrize("d,n,B,L", [(10, 1000, 700, 5), (100, 104, 1500, 3), (100, 10_000, 256, 4), (37, 64, 130, 1)])
def test_hmc_fused_leapfrog_step_agrees_with_the_separate_kernels(d, n, B, L):
    """One launch per leapfrog step (kicks, drift, Hamiltonian, Metropolis test and state update inside the gradient
    kernel's output step -- streaming and resident variants, stream-K combine included) against the round-1 path of
    separate elementwise kernels: same draws, same accept decisions, states equal to rounding."""
    prob = linear_gaussian_problem(d, n, seed=d + n)
    eps = 0.3 * 0.1 * np.sqrt(d / max(n, d))
    cfg = t.HmcmcConfig(stepsBetweenSamples=1, stepsInDynamicsSimulation=L, warmupStepCount=2, dynamicsSimulationStepSize=eps)
    mean, cov = o.analytic_gaussian_posterior(prob["pm"], prob["pvar"], prob["A"], prob["y"], prob["var"])
    seeds = mean[None, :] + 2.0 * np.sqrt(np.diag(cov))[None, :] * np.random.default_rng(2).standard_normal((B, d))
    out = []
    for fused in (True, False):
        post, _ = _posteriors(prob)
        hmc = t.HmcmcSampledPosterior(cfg, post, seed=seeds, numberOfChains=B, rngSeed=77, fusedSteps=fused)
        hmc.burnIn()
        hmc.runTrajectories(6)
        x, lp = hmc.state()
        att, acc, dh = hmc.stats()
        out.append((x, lp, att, acc, dh, hmc.trajectoryCounter()))
        hmc.close()
        post.close()
    (xa, la, ata, aca, dha, ta), (xb, lb, atb, acb, dhb, tb) = out
    assert ta == tb == 3 + 6 and np.array_equal(ata, atb)
    same = aca == acb                                   # a decision can only differ where |u - exp(-dH)| ~ 1e-15
    assert same.mean() > 0.999
    sd = np.sqrt(np.diag(cov))
    assert np.max(np.abs(xa[same] - xb[same]) / sd[None, :]) < 1e-8
    assert np.max(np.abs(la[same] - lb[same]) / np.maximum(1.0, np.abs(lb[same]))) < 1e-10
    assert 0.3 < aca.sum() / ata.sum() <= 1.0


def test_hmc_checkpoint_resume_continues_the_same_chains():
    """thy_hmc_set_state + trajectory counter (checkpoint row): a new sampler given the positions and
    the counter of a checkpoint continues bit for bit like the run that was checkpointed."""
    d, n, B = 12, 300, 96
    prob = linear_gaussian_problem(d, n, seed=6)
    cfg = t.HmcmcConfig(stepsBetweenSamples=1, stepsInDynamicsSimulation=4, warmupStepCount=3, dynamicsSimulationStepSize=0.004)
    post, _ = _posteriors(prob)
    a = t.HmcmcSampledPosterior(cfg, post, numberOfChains=B, rngSeed=21, seed={"theta": prob["theta"]})
    a.burnIn()
    a.runTrajectories(5)
    x_ck, _ = a.state()
    t_ck = a.trajectoryCounter()
    a.runTrajectories(7)
    x_end, lp_end = a.state()
    b = t.HmcmcSampledPosterior(cfg, post, numberOfChains=B, rngSeed=21, seed={"theta": prob["theta"]})
    b.setState(x_ck, t_ck)
    b.runTrajectories(7)
    x_res, lp_res = b.state()
    assert t_ck == 4 + 5 and b.trajectoryCounter() == t_ck + 7
    assert np.array_equal(x_end, x_res) and np.array_equal(lp_end, lp_res)
    a.close()
    b.close()
    post.close()


@pytest.mark.parametrize("fused", [True, False])
def test_hmc_zero_dynamics_steps_always_accepts_with_zero_energy_error(fused):
    """stepsInDynamicsSimulation = 0: the reference's leapfrog fold is empty, the proposal is the current point with the
    momentum untouched, dH == 0 and the jump is always accepted (HmcmcEngine.scala:55-80, :128)."""
    prob = linear_gaussian_problem(5, 60, seed=8)
    post, _ = _posteriors(prob)
    cfg = t.HmcmcConfig(stepsBetweenSamples=0, stepsInDynamicsSimulation=0, warmupStepCount=0, dynamicsSimulationStepSize=0.01)
    hmc = t.HmcmcSampledPosterior(cfg, post, numberOfChains=40, rngSeed=3, seed={"theta": prob["theta"]}, fusedSteps=fused)
    hmc.burnIn()
    x0, lp0 = hmc.state()
    hmc.runTrajectories(3)
    x1, lp1 = hmc.state()
    att, acc, dh = hmc.stats()
    assert np.array_equal(x0, x1) and np.array_equal(lp0, lp1)
    assert np.all(att == 4) and np.all(acc == 4) and np.all(dh == 0.0)
    hmc.close()
    post.close()
