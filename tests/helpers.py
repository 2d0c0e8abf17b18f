"""Synthetic models shared by the parity tests (recipe: SURVEY.md section 8d)."""
import numpy as np

from oracle import ref_oracle as o


def linear_gaussian_problem(d, n, seed, sigma=0.1, prior_sigma=10.0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, d)) / np.sqrt(d)
    theta = rng.standard_normal(d)
    y = A @ theta + sigma * rng.standard_normal(n)
    ci = np.full(n, 2.0 * sigma)
    pm = np.zeros(d)
    pci = np.full(d, 2.0 * prior_sigma)
    return dict(A=A, theta=theta, y=y, ci=ci, pm=pm, pci=pci, var=o.ci_to_variance(ci), pvar=o.ci_to_variance(pci))


def typical_set_points(prob, B, seed, scale=2.0):
    """X ~ N(mu_post, scale^2 Sigma_post): residuals O(sigma), where cancellation is worst."""
    rng = np.random.default_rng(seed)
    mean, cov = o.analytic_gaussian_posterior(prob["pm"], prob["pvar"], prob["A"], prob["y"], prob["var"])
    L = np.linalg.cholesky(cov)
    return mean[None, :] + scale * rng.standard_normal((B, mean.size)) @ L.T


def rel_err_logp(got, want):
    return float(np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))))


def rel_err_grad(got, want):
    """max-norm error relative to ||grad||_inf of the oracle, per point (SURVEY 8c tolerances)."""
    scale = np.maximum(np.max(np.abs(want), axis=-1, keepdims=True), 1e-300)
    return float(np.max(np.abs(got - want) / scale))
