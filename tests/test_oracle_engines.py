"""CPU checks of the oracle's sequential engines themselves (HmcmcEngine.scala:55-214, LeapfrogMcmcEngine.scala:88-294,
QuadratureAbscissa.scala:33-73 restated in oracle/ref_oracle.py).  The device replay tests compare against these
engines, so they have to be right on their own: both samplers must recover an analytic Gaussian posterior, and the
reference's quirks (Q5) must be there."""
import math

import numpy as np

from oracle import ref_oracle as o


def _model(d=3, n=12, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, d))
    y = A @ rng.standard_normal(d) + 0.3 * rng.standard_normal(n)
    ci, pm, pci = np.full(n, 0.6), np.zeros(d), np.full(d, 6.0)
    post = o.Posterior([o.gaussian_prior_from_ci("theta", pm, pci)], [o.gaussian_linear_likelihood(A, y, ci, "theta")])
    mean, cov = o.analytic_gaussian_posterior(pm, o.ci_to_variance(pci), A, y, o.ci_to_variance(ci))
    return post, mean, cov


def test_oracle_hmc_recovers_the_analytic_posterior_and_keeps_the_reference_quirks():
    post, mean, cov = _model()
    sd = np.sqrt(np.diag(cov))
    cfg = o.HmcmcConfig(stepsBetweenSamples=1, stepsInDynamicsSimulation=7, warmupStepCount=30,
                        dynamicsSimulationStepSize=0.35 * float(sd.min()))
    eng = o.HmcmcEngine(post, cfg, rng=np.random.default_rng(3), seed_point=mean.copy())
    s = np.array(eng.sample(600))
    # Q5: warmupStepCount + 1 burn-in trajectories, stepsBetweenSamples + 1 per block; rejected blocks emit nothing
    assert eng.jump_attempts >= (cfg.warmupStepCount + 1) + 600 * (cfg.stepsBetweenSamples + 1)
    assert (eng.jump_attempts - (cfg.warmupStepCount + 1)) % (cfg.stepsBetweenSamples + 1) == 0
    assert len({x.tobytes() for x in s}) == 600                                # Set semantics: no repeated sample
    ess = 600 / 3
    assert np.max(np.abs(s.mean(0) - mean) / sd) < 5.0 / math.sqrt(ess)
    assert np.max(np.abs(s.std(0) / sd - 1.0)) < 5.0 / math.sqrt(2 * ess) + 0.05
    assert 0.5 < eng.jump_acceptances / eng.jump_attempts <= 1.0


def test_oracle_leapfrog_mcmc_recovers_the_analytic_posterior():
    post, mean, cov = _model(d=2, n=10, seed=1)
    sd = np.sqrt(np.diag(cov))
    cfg = o.LeapfrogMcmcConfig(stepsBetweenSamples=4, warmupStepCount=60, samplePoolSize=24)
    eng = o.LeapfrogMcmcEngine(post, cfg, rng=np.random.default_rng(5), seed_point=mean.copy())
    s = np.array(eng.sample(1500))
    assert eng.proposals > 0 and 0.02 < eng.acceptances / eng.proposals < 0.98
    # an ensemble of 24 members mixes slowly: generous ESS, the point is the stationary distribution
    ess = 1500 / 25
    assert np.max(np.abs(s.mean(0) - mean) / sd) < 5.0 / math.sqrt(ess)
    assert np.all(s.std(0) / sd > 0.5) and np.all(s.std(0) / sd < 1.8)


def test_oracle_abscissa_shrinks_like_order_statistics():
    # E[ln X_i] = -i / P for the i-th retired point (QuadratureAbscissa.scala:33-37: max of P uniforms, refilled below it)
    P, steps, reps = 40, 60, 200
    rng = np.random.default_rng(9)
    acc = np.zeros(steps)
    for _ in range(reps):
        qa = o.QuadratureAbscissa(P, rng)
        for _ in range(steps):
            qa.extend_by_one()
        acc += np.log(np.array(qa.abscissa[:steps]))
    got = acc / reps
    want = -np.arange(steps) / P                                               # the first abscissa is X = 1 exactly
    assert abs(got[0]) < 1e-12
    assert np.max(np.abs(got[1:] - want[1:])) < 5.0 * math.sqrt(steps) / P / math.sqrt(reps) + 0.02
