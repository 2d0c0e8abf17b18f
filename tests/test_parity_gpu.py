"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's golden vectors.

Tolerance (north_star): log-posterior and gradient within 1e-10 relative in fp64 -- gradient relative to
||grad||_inf of the oracle, log-pdf relative to |logpdf| (absolute 1e-10 when |logpdf| < 1).
"""
import json
import math
from pathlib import Path

import numpy as np
import pytest

import thylacine_b200 as t
from thylacine_b200 import capi
from oracle import ref_oracle as o
from helpers import linear_gaussian_problem, typical_set_points, rel_err_logp, rel_err_grad

pytestmark = pytest.mark.gpu
TOL = 1e-10
GOLDEN = json.loads((Path(__file__).parent / "golden" / "reference_specs.json").read_text())
PATHS = [capi.KERNEL_GENERIC, capi.KERNEL_FUSED_DMMA]


# ---- the reference's own specs, through the drop-in surface --------------------------------------------
def test_GaussianPriorSpec_gradient_exact():
    g = GOLDEN["gaussian_prior_gradient"]
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("p", g["mean"], g["ci"])])
    got = post.logPdfGradientAt({"p": g["x"]})["p"]
    assert np.array_equal(got, np.array([-0.25, 4.0 / 9.0, -1.0]))           # GaussianPriorSpec.scala:57


@pytest.mark.parametrize("path", PATHS)
def test_GaussianLinearLikelihoodSpec_gradient_exact(path):
    f = GOLDEN["foo_likelihood"]
    # a flat improper-ish prior is not available in the reference either: use the uniform fixture prior, whose
    # gradient is exactly zero (UniformDistribution.scala:78-81), so the posterior gradient IS the likelihood's.
    post = t.UnnormalisedPosterior([t.UniformPrior.fromBounds("foo", [5, 5], [-5, -5])],
                                   [t.GaussianLinearLikelihood.of(f["A"], f["y"], f["ci"], "foo")], kernelPath=path)
    assert np.array_equal(post.logPdfGradientAt({"foo": [1, 2]})["foo"], [0.0, 0.0])          # :36-38
    assert np.array_equal(post.logPdfGradientAt({"foo": [3, 2]})["foo"], [-4e5, -8.8e5])      # :45-47
    lv = -2 * math.log(10.0)
    d = GOLDEN["derived"]
    assert post.logPdfAt({"foo": [1, 2]}) == pytest.approx(d["foo_likelihood_logpdf_at_1_2"] + lv, rel=1e-13)
    assert post.logPdfAt({"foo": [3, 2]}) == pytest.approx(d["foo_likelihood_logpdf_at_3_2"] + lv, rel=1e-13)


@pytest.mark.parametrize("incremental", [False, True])
def test_GaussianLikelihoodSpec_fd_gradient(incremental):
    f = GOLDEN["foo_likelihood"]
    fm = t.NonLinearForwardModel.of("identity", {"foo": f["A"]}, differential=f["fd_differential"],
                                    domainDimensions={"foo": 2}, rangeDimension=2)
    post = t.UnnormalisedPosterior([t.UniformPrior.fromBounds("foo", [5, 5], [-5, -5])],
                                   [t.GaussianLikelihood(fm, f["y"], f["ci"])], fdIncremental=incremental)
    assert np.array_equal(post.logPdfGradientAt({"foo": [1, 2]})["foo"], [0.0, 0.0])          # GaussianLikelihoodSpec:37-39
    got = post.logPdfGradientAt({"foo": [3, 2]})["foo"]
    assert np.max(np.abs(got - np.array([-4e5, -8.8e5]))) < f["fd_abs_tol"]                    # :46-48


def test_fixture_posterior_values_and_label_order():
    foo, bar = GOLDEN["foo_likelihood"], GOLDEN["bar_likelihood"]
    priors = [t.GaussianPrior.fromConfidenceIntervals("foo", [1, 2], [3, 5]),
              t.GaussianPrior.fromConfidenceIntervals("bar", [5], [0.1])]
    liks = [t.GaussianLinearLikelihood.of(foo["A"], foo["y"], foo["ci"], "foo"),
            t.GaussianLinearLikelihood.of(bar["A"], bar["y"], bar["ci"], "bar")]
    post = t.UnnormalisedPosterior(priors, liks)
    assert post.layout == {"bar": (0, 1), "foo": (1, 2)}                      # Posterior.scala:40-48
    ref = o.Posterior([o.gaussian_prior_from_ci("foo", [1, 2], [3, 5]), o.gaussian_prior_from_ci("bar", [5], [0.1])],
                      [o.gaussian_linear_likelihood(foo["A"], foo["y"], foo["ci"], "foo"),
                       o.gaussian_linear_likelihood(bar["A"], bar["y"], bar["ci"], "bar")])
    for x in ([5.0, 1.0, 2.0], [4.9, 3.0, 2.0], [5.2, -1.0, 0.5]):
        params = {"bar": x[:1], "foo": x[1:]}
        assert post.logPdfAt(params) == pytest.approx(ref.log_pdf(np.array(x)), rel=TOL)
        g = post.logPdfGradientAt(params)
        want = ref.log_pdf_gradient(np.array(x))
        got = np.concatenate([g["bar"], g["foo"]])
        assert rel_err_grad(got[None], want[None]) < TOL


def test_uniform_prior_half_open_and_minus_inf():
    post = t.UnnormalisedPosterior([t.UniformPrior.fromBounds("u", [5, 5], [-5, -5])])
    lp = post.logPdfBatch([[0, 0], [-5, 0], [5, 0], [0, 4.999999], [0, np.nan]])
    assert lp[0] == pytest.approx(-2 * math.log(10)) and lp[1] == lp[0] and lp[3] == lp[0]
    assert lp[2] == -math.inf and lp[4] == -math.inf                          # UniformDistribution.scala:61-76
    _, g = post.logPdfGradBatch([[0, 0], [7, 0]])
    assert np.array_equal(g, np.zeros((2, 2)))


# ---- random linear-Gaussian models: every kernel path vs the oracle ---------------------------------------
SHAPES = [(1, 1, 1), (2, 2, 3), (3, 5, 7), (8, 32, 64), (10, 1000, 65), (12, 200, 96), (17, 333, 130),
          (24, 100, 1), (33, 257, 200), (50, 50, 129), (56, 640, 64), (80, 500, 70), (100, 1000, 256), (128, 300, 100)]


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("d,n,B", SHAPES)
def test_linear_gaussian_random(path, d, n, B):
    prob = linear_gaussian_problem(d, n, seed=1000 + d * 7 + n)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                   [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")],
                                   kernelPath=path)
    rng = np.random.default_rng(d + n + B)
    for X in (typical_set_points(prob, B, seed=B) if n >= d else rng.standard_normal((B, d)),
              prob["pm"] + 10.0 * rng.standard_normal((B, d))):               # typical set and prior far field
        want_lp, want_g = o.linear_gaussian_batch(prob["A"], prob["y"], prob["var"], X, prior_mean=prob["pm"],
                                                  prior_var=prob["pvar"])
        lp, g = post.logPdfGradBatch(X)
        assert rel_err_logp(lp, want_lp) < TOL
        assert rel_err_grad(g, want_g) < TOL
        lp_only = post.logPdfBatch(X)
        assert np.array_equal(lp_only, lp) or rel_err_logp(lp_only, want_lp) < TOL
    # the fused path has two variants: A streamed through a TMA ring, or resident in shared memory when it fits
    assert post.lastKernelPath in (("generic",) if path == capi.KERNEL_GENERIC else ("fused_dmma", "resident_dmma"))


def test_literal_dense_reference_agrees_at_small_size():
    """The literal restatement (dense covariance, LU solve per logPdf) vs the CUDA path, one point at a time."""
    prob = linear_gaussian_problem(6, 40, seed=5)
    ref = o.Posterior([o.gaussian_prior_from_ci("theta", prob["pm"], prob["pci"])],
                      [o.gaussian_linear_likelihood(prob["A"], prob["y"], prob["ci"], "theta")])
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                   [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")])
    X = typical_set_points(prob, 9, seed=1)
    lp, g = post.logPdfGradBatch(X)
    for b in range(X.shape[0]):
        assert lp[b] == pytest.approx(ref.log_pdf(X[b]), rel=TOL)
        assert rel_err_grad(g[b][None], ref.log_pdf_gradient(X[b])[None]) < TOL


@pytest.mark.parametrize("path", PATHS)
def test_multi_label_blocks_offset_and_two_likelihoods(path):
    rng = np.random.default_rng(11)
    n1, n2 = 70, 45
    dims = {"alpha": 3, "beta": 5, "gamma": 2}
    A1 = {"gamma": rng.standard_normal((n1, 2)), "alpha": rng.standard_normal((n1, 3))}      # skips beta: non-contiguous
    A2 = {"beta": rng.standard_normal((n2, 5))}
    off = rng.standard_normal(n1)
    y1, y2 = rng.standard_normal(n1), rng.standard_normal(n2)
    ci1, ci2 = rng.uniform(0.1, 2, n1), rng.uniform(0.1, 2, n2)
    cov = np.array([[2.0, 0.3], [0.3, 1.0]])
    priors = [t.GaussianPrior.fromConfidenceIntervals("beta", rng.standard_normal(5), rng.uniform(1, 3, 5)),
              t.UniformPrior.fromBounds("alpha", [4, 4, 4], [-4, -4, -4]),
              t.GaussianPrior.fromCovarianceMatrix("gamma", [0.5, -0.5], cov)]
    liks = [t.GaussianLikelihood(t.LinearForwardModel.of(transform=A1, vectorOffset=off), y1, ci1),
            t.GaussianLinearLikelihood.of(A2["beta"], y2, ci2, "beta")]
    post = t.UnnormalisedPosterior(priors, liks, kernelPath=path)
    ref = o.Posterior([o.gaussian_prior_from_ci("beta", priors[0].a, priors[0].b),
                       o.uniform_prior_from_bounds("alpha", [4, 4, 4], [-4, -4, -4]),
                       o.gaussian_prior_from_cov("gamma", [0.5, -0.5], cov)],
                      [o.gaussian_likelihood(o.LinearForwardModel(A1, off), y1, ci1),
                       o.gaussian_linear_likelihood(A2["beta"], y2, ci2, "beta")])
    assert post.layout == {"alpha": (0, 3), "beta": (3, 5), "gamma": (8, 2)}
    X = rng.uniform(-3.9, 3.9, (37, 10))
    X[5, 1] = 4.5                                                              # outside the uniform box
    lp, g = post.logPdfGradBatch(X)
    for b in range(X.shape[0]):
        want = ref.log_pdf(X[b])
        if math.isinf(want):
            assert lp[b] == -math.inf
        else:
            assert lp[b] == pytest.approx(want, rel=TOL)
        assert rel_err_grad(g[b][None], ref.log_pdf_gradient(X[b])[None]) < TOL


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("faithful", [True, False])
def test_cauchy_linear_likelihood_and_prior(path, faithful):
    rng = np.random.default_rng(3)
    d, n, B = 7, 60, 33
    A = rng.standard_normal((n, d))
    y = rng.standard_normal(n)
    ci = rng.uniform(0.5, 2.0, n)
    pm, pci = rng.standard_normal(d), rng.uniform(1, 3, d)
    post = t.UnnormalisedPosterior([t.CauchyPrior("c", pm, pci)], [t.CauchyLikelihood.of(A, y, ci, "c")],
                                   cauchyReferenceSign=faithful, kernelPath=path)
    ref = o.Posterior([o.cauchy_prior("c", pm, pci, faithful)],
                      [o.cauchy_likelihood(o.LinearForwardModel({"c": A}), y, ci, faithful)])
    X = rng.standard_normal((B, d))
    lp, g = post.logPdfGradBatch(X)
    for b in range(B):
        assert lp[b] == pytest.approx(ref.log_pdf(X[b]), rel=TOL)
        assert rel_err_grad(g[b][None], ref.log_pdf_gradient(X[b])[None]) < TOL


@pytest.mark.parametrize("link", ["tanh", "exp", "square", "softplus", "sigmoid"])
@pytest.mark.parametrize("dist", ["gaussian", "cauchy"])
def test_nonlinear_analytic_jacobian(link, dist):
    rng = np.random.default_rng(21)
    d, n, B = 9, 50, 20
    A = rng.standard_normal((n, d)) / 3
    off = 0.1 * rng.standard_normal(n)
    fn = {"tanh": np.tanh, "exp": np.exp, "square": np.square, "softplus": lambda z: np.logaddexp(0, z),
          "sigmoid": lambda z: 1 / (1 + np.exp(-z))}[link]
    dfn = {"tanh": lambda z: 1 - np.tanh(z) ** 2, "exp": np.exp, "square": lambda z: 2 * z,
           "softplus": lambda z: 1 / (1 + np.exp(-z)), "sigmoid": lambda z: fn(z) * (1 - fn(z))}[link]
    y = fn(A @ rng.standard_normal(d) + off) + 0.05 * rng.standard_normal(n)
    ci = np.full(n, 0.1)
    pm, pci = np.zeros(d), np.full(d, 4.0)
    fm = t.NonLinearForwardModel.of(link, {"w": A}, vectorOffset=off)
    lk = t.GaussianLikelihood(fm, y, ci) if dist == "gaussian" else t.CauchyLikelihood(fm, y, ci)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("w", pm, pci)], [lk], cauchyReferenceSign=False)
    ofm = o.NonLinearForwardModel(lambda p: fn(A @ p["w"] + off), None, {"w": d}, n,
                                  jacobian=lambda p: {"w": dfn(A @ p["w"] + off)[:, None] * A})
    olk = o.gaussian_likelihood(ofm, y, ci) if dist == "gaussian" else o.cauchy_likelihood(ofm, y, ci, False)
    ref = o.Posterior([o.gaussian_prior_from_ci("w", pm, pci)], [olk])
    X = 0.5 * rng.standard_normal((B, d))
    lp, g = post.logPdfGradBatch(X)
    for b in range(B):
        assert lp[b] == pytest.approx(ref.log_pdf(X[b]), rel=TOL)
        assert rel_err_grad(g[b][None], ref.log_pdf_gradient(X[b])[None]) < TOL


@pytest.mark.parametrize("incremental", [False, True])
def test_nonlinear_fd_jacobian_matches_reference_fd(incremental):
    """FiniteDifferenceJacobian.scala:30-55.  A forward difference divides rounding noise of size eps*|f| by h, so
    two correct implementations that sum A.theta in different orders agree only to ~ sqrt(d) eps |f| / (h |J|);
    the tolerance below states that bound (1e-7 at h = 1e-6), not 1e-10."""
    rng = np.random.default_rng(5)
    d, n, B = 20, 64, 16
    h = 1e-6
    A = rng.standard_normal((n, d)) / np.sqrt(d)
    y = np.tanh(A @ rng.standard_normal(d)) + 0.05 * rng.standard_normal(n)
    ci = np.full(n, 0.1)
    pm, pci = np.zeros(d), np.full(d, 4.0)
    fm = t.NonLinearForwardModel.of("tanh", {"w": A}, differential=h)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("w", pm, pci)],
                                   [t.CauchyLikelihood(fm, y, ci)], cauchyReferenceSign=True, fdIncremental=incremental)
    ofm = o.NonLinearForwardModel(lambda p: np.tanh(A @ p["w"]), h, {"w": d}, n)
    ref = o.Posterior([o.gaussian_prior_from_ci("w", pm, pci)], [o.cauchy_likelihood(ofm, y, ci, True)])
    X = 0.5 * rng.standard_normal((B, d))
    lp, g = post.logPdfGradBatch(X)
    for b in range(B):
        assert lp[b] == pytest.approx(ref.log_pdf(X[b]), rel=TOL)
        assert rel_err_grad(g[b][None], ref.log_pdf_gradient(X[b])[None]) < 1e-7


def test_uniform_likelihood():
    rng = np.random.default_rng(9)
    A = rng.standard_normal((5, 3))
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("a", [0, 0, 0], [2, 2, 2])],
                                   [t.UniformLikelihood(t.LinearForwardModel.of("a", A), np.full(5, 1.0), np.full(5, -1.0))])
    ref = o.Posterior([o.gaussian_prior_from_ci("a", [0, 0, 0], [2, 2, 2])],
                      [o.Likelihood(o.LinearForwardModel({"a": A}), o.UniformDistribution(np.full(5, 1.0), np.full(5, -1.0)))])
    X = 0.6 * rng.standard_normal((64, 3))
    lp, g = post.logPdfGradBatch(X)
    n_in = 0
    for b in range(64):
        want = ref.log_pdf(X[b])
        if math.isinf(want):
            assert lp[b] == -math.inf
        else:
            n_in += 1
            assert lp[b] == pytest.approx(want, rel=TOL)
        assert rel_err_grad(g[b][None], ref.log_pdf_gradient(X[b])[None]) < TOL
    assert 0 < n_in < 64


def test_empty_batch_and_ragged_tail():
    prob = linear_gaussian_problem(10, 100, seed=2)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                   [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")])
    assert post.logPdfBatch(np.zeros((0, 10))).shape == (0,)
    for B in (1, 7, 63, 64, 65, 127, 129, 1000):
        X = typical_set_points(prob, B, seed=B)
        want_lp, want_g = o.linear_gaussian_batch(prob["A"], prob["y"], prob["var"], X, prior_mean=prob["pm"],
                                                  prior_var=prob["pvar"])
        lp, g = post.logPdfGradBatch(X)
        assert rel_err_logp(lp, want_lp) < TOL and rel_err_grad(g, want_g) < TOL


def test_run_to_run_determinism():
    prob = linear_gaussian_problem(40, 3000, seed=8)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                   [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")])
    X = typical_set_points(prob, 700, seed=3)
    a = post.logPdfGradBatch(X)
    b = post.logPdfGradBatch(X)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_full_size_config2_properties():
    """BASELINE config 2 (100-d, 10k observations, 4096 chains): oracle parity on a slice, plus size-independent
    properties on the whole batch -- the gradient vanishes at the analytic posterior mean, and
    logp(x) - logp(mu) = -1/2 (x-mu)^T Lambda (x-mu) (exact for a Gaussian posterior)."""
    prob = linear_gaussian_problem(100, 10_000, seed=20260102)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                   [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")])
    X = typical_set_points(prob, 4096, seed=4)
    lp, g = post.logPdfGradBatch(X)
    assert post.lastKernelPath in ("fused_dmma", "resident_dmma")
    want_lp, want_g = o.linear_gaussian_batch(prob["A"], prob["y"], prob["var"], X, prior_mean=prob["pm"],
                                              prior_var=prob["pvar"])
    assert rel_err_logp(lp, want_lp) < TOL and rel_err_grad(g, want_g) < TOL
    mean, cov = o.analytic_gaussian_posterior(prob["pm"], prob["pvar"], prob["A"], prob["y"], prob["var"])
    lam = np.linalg.inv(cov)
    lp0, g0 = post.logPdfGradBatch(mean[None, :])
    gscale = np.max(np.abs(g))
    assert np.max(np.abs(g0)) < 1e-9 * gscale                                 # stationary point
    D = X - mean[None, :]
    quad = -0.5 * np.einsum("bi,ij,bj->b", D, lam, D)
    assert np.max(np.abs((lp - lp0[0]) - quad)) < 1e-9 * np.max(np.abs(quad)) + 1e-7
    assert np.max(np.abs(g + D @ lam)) < 1e-9 * gscale                        # grad = -Lambda (x - mu)


# ---- user-defined links compiled with NVRTC (NonLinearForwardModel.of(evaluation = closure)) ----------------------------
@pytest.mark.parametrize("d,n,B", [(7, 60, 33), (200, 256, 300), (64, 5, 1), (129, 700, 130)])
def test_user_link_matches_the_oracle_closure_and_the_registry(d, n, B):
    """thy_model_register_link: a softsign family given as CUDA source runs between the two DMMA GEMMs of the wide path
    and must agree with the oracle's NonLinearForwardModel built from the same function as Python closures
    (NonLinearForwardModel.scala:44-85, analytic Jacobian overload :64-85); the registry's tanh re-registered as a user link
    must reproduce the built-in family bit for bit (same expression, same kernels either side)."""
    rng = np.random.default_rng(d + n)
    A = rng.standard_normal((n, d)) / np.sqrt(d)
    off = 0.1 * rng.standard_normal(n)
    y = 0.5 * rng.standard_normal(n)
    ci = np.full(n, 0.4)
    pm, pci = np.zeros(d), np.full(d, 4.0)
    X = 0.7 * rng.standard_normal((B, d))
    soft = t.UserLink("z / (1.0 + fabs(z))", "1.0 / ((1.0 + fabs(z)) * (1.0 + fabs(z)))")
    fm = t.NonLinearForwardModel.of(soft, {"w": A}, vectorOffset=off)
    # (Gaussian observations: the reference's Cauchy normaliser overflows to +inf beyond n ~ 210)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("w", pm, pci)], [t.GaussianLikelihood(fm, y, ci)])
    f = lambda p: (A @ p["w"] + off) / (1.0 + np.abs(A @ p["w"] + off))
    jac = lambda p: {"w": A / ((1.0 + np.abs(A @ p["w"] + off)) ** 2)[:, None]}
    ofm = o.NonLinearForwardModel(f, None, {"w": d}, n, jacobian=jac)
    ref = o.Posterior([o.gaussian_prior_from_ci("w", pm, pci)], [o.gaussian_likelihood(ofm, y, ci)])
    lp, g = post.logPdfGradBatch(X)
    assert post.lastKernelPath == "wide_dmma"
    for b in range(min(B, 40)):
        assert lp[b] == pytest.approx(ref.log_pdf(X[b]), rel=TOL)
        assert rel_err_grad(g[b][None], ref.log_pdf_gradient(X[b])[None]) < TOL
    assert rel_err_logp(post.logPdfBatch(X), lp) < 1e-14
    # forward difference of the link itself when no derivative is given (same limit as FiniteDifferenceJacobian.scala:30-55)
    fd = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("w", pm, pci)],
                                 [t.GaussianLikelihood(t.NonLinearForwardModel.of(t.UserLink("z / (1.0 + fabs(z))"), {"w": A},
                                                                                  differential=1e-6, vectorOffset=off), y, ci)])
    lp_fd, g_fd = fd.logPdfGradBatch(X)
    assert np.array_equal(lp_fd, lp) and rel_err_grad(g_fd, g) < 1e-5
    # the registry's tanh, re-registered as source
    builtin = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("w", pm, pci)],
                                      [t.GaussianLikelihood(t.NonLinearForwardModel.of("tanh", {"w": A}, vectorOffset=off), y, ci)],
                                      kernelPath=capi.KERNEL_WIDE)
    user = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("w", pm, pci)],
                                   [t.GaussianLikelihood(t.NonLinearForwardModel.of(t.UserLink("tanh(z)", "1.0 - f * f"), {"w": A},
                                                                                    vectorOffset=off), y, ci)])
    lp_b, g_b = builtin.logPdfGradBatch(X)
    lp_u, g_u = user.logPdfGradBatch(X)
    assert np.array_equal(lp_b, lp_u) and np.array_equal(g_b, g_u)
    for p_ in (post, fd, builtin, user):
        p_.close()


# ---- resident variant of the fused kernel: A held in shared memory for the CTA's lifetime ---------------------------
RESIDENT_SHAPES = [(100, 100, 4096), (100, 100, 4099), (100, 104, 700), (10, 1000, 5000), (128, 128, 1200), (3, 250, 9000),
                   (64, 8, 333), (1, 1, 3000)]


@pytest.mark.parametrize("d,n,B", RESIDENT_SHAPES)
@pytest.mark.parametrize("fuse", [0, 2])
def test_resident_kernel_matches_oracle_and_the_streaming_variant(d, n, B, fuse):
    """kernels_fused_dmma.cu k_resident_dmma (the collapsed-model / config-1 shape) against the oracle, with the priors
    folded into the output step (fusePrior=2) and added to a prior pass (0); the variant that streams A through the TMA
    ring (kernelPath 5) must agree to rounding on the same inputs."""
    prob = linear_gaussian_problem(d, n, seed=d * 31 + n)
    mk = lambda path: t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                              [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")],
                                              kernelPath=path, fusePrior=fuse)
    res, stm = mk(capi.KERNEL_FUSED_DMMA), mk(capi.KERNEL_FUSED_STREAMING)
    rng = np.random.default_rng(B)
    X = (typical_set_points(prob, B, seed=B) if n >= d else rng.standard_normal((B, d)))
    want_lp, want_g = o.linear_gaussian_batch(prob["A"], prob["y"], prob["var"], X, prior_mean=prob["pm"], prior_var=prob["pvar"])
    lp, g = res.logPdfGradBatch(X)
    assert res.lastKernelPath == "resident_dmma"
    assert rel_err_logp(lp, want_lp) < TOL and rel_err_grad(g, want_g) < TOL
    lp2, g2 = res.logPdfGradBatch(X)
    assert np.array_equal(lp, lp2) and np.array_equal(g, g2)                # deterministic
    assert rel_err_logp(res.logPdfBatch(X), want_lp) < TOL                  # value-only instantiation
    lps, gs = stm.logPdfGradBatch(X)
    assert stm.lastKernelPath == "fused_dmma"
    assert rel_err_logp(lps, want_lp) < TOL and rel_err_grad(gs, want_g) < TOL
    res.close()
    stm.close()


def test_async_host_buffer_calls_deliver_the_synchronous_results():
    """thy_logpdf_grad_batch_async / thy_model_wait: calls pipelined two deep over two staging sets, batches of changing
    size (a staging set has to grow while the other is in flight), value-only calls; every delivery equals the blocking call."""
    import torch
    prob = linear_gaussian_problem(24, 900, seed=12)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                   [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")])
    sizes = [300, 300, 1200, 64, 5000, 5000, 7, 2048]
    Xs = [torch.from_numpy(typical_set_points(prob, B, seed=i)).pin_memory() for i, B in enumerate(sizes)]
    lps = [torch.empty(B, dtype=torch.float64).pin_memory() for B in sizes]
    gs = [torch.empty(B, 24, dtype=torch.float64).pin_memory() for B in sizes]
    tickets = []
    for i, B in enumerate(sizes):
        grad = None if i == 3 else gs[i].numpy()
        tickets.append(post.logPdfGradBatchAsync(Xs[i].numpy(), lps[i].numpy(), grad))
        if i >= 1:
            post.wait(tickets[i - 1])
            # blocking call on the same model, in between.  It cuts batches >= 2048 into [small, large, small] chunks, so
            # its stream-K partial sums combine in a different order: equal to rounding, not bit for bit
            want_lp, want_g = post.logPdfGradBatch(Xs[i - 1].numpy())
            assert rel_err_logp(lps[i - 1].numpy(), want_lp) < 1e-13
            if i - 1 != 3:
                assert rel_err_grad(gs[i - 1].numpy(), want_g) < 1e-12
    assert tickets == list(range(1, len(sizes) + 1))
    post.wait()
    want_lp, want_g = post.logPdfGradBatch(Xs[-1].numpy())
    assert rel_err_logp(lps[-1].numpy(), want_lp) < 1e-13 and rel_err_grad(gs[-1].numpy(), want_g) < 1e-12
    # the same call twice delivers the same bits
    lp2, g2 = torch.empty(2048, dtype=torch.float64).pin_memory(), torch.empty(2048, 24, dtype=torch.float64).pin_memory()
    post.wait(post.logPdfGradBatchAsync(Xs[-1].numpy(), lp2.numpy(), g2.numpy()))
    assert np.array_equal(lp2.numpy(), lps[-1].numpy()) and np.array_equal(g2.numpy(), gs[-1].numpy())
    post.close()


# ---- small-batch streaming kernel (B <= 8): TMA-staged memory-bound pass -----------------------------------
STREAM_SHAPES = [(1, 1, 1), (3, 5, 2), (10, 1000, 1), (10, 1000, 8), (31, 64, 3), (32, 33, 4), (33, 257, 5), (64, 4000, 7),
                 (100, 10_000, 1), (100, 10_000, 6), (128, 300, 8), (97, 20_000, 2),
                 # long models: every CTA goes round its TMA ring more than once (static stage ownership keeps the
                 # row-group kernel sound under reuse, kernels_stream.cu)
                 (5, 40_000, 1), (5, 40_000, 3), (8, 50_000, 8), (20, 60_000, 16), (100, 300_000, 8), (128, 200_000, 5)]


@pytest.mark.parametrize("d,n,B", STREAM_SHAPES)
def test_stream_kernel_linear_gaussian(d, n, B):
    prob = linear_gaussian_problem(d, n, seed=77 + d + n)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                   [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")],
                                   kernelPath=capi.KERNEL_STREAM)
    rng = np.random.default_rng(B)
    for X in (typical_set_points(prob, B, seed=B) if n >= d else rng.standard_normal((B, d)),
              prob["pm"] + 10.0 * rng.standard_normal((B, d))):
        want_lp, want_g = o.linear_gaussian_batch(prob["A"], prob["y"], prob["var"], X, prior_mean=prob["pm"],
                                                  prior_var=prob["pvar"])
        lp, g = post.logPdfGradBatch(X)
        assert post.lastKernelPath == ("stream" if B == 1 else "stream_dmma")   # 2+ chains: the DMMA row-group kernel
        assert rel_err_logp(lp, want_lp) < TOL
        assert rel_err_grad(g, want_g) < TOL
        lp2, g2 = post.logPdfGradBatch(X)                       # deterministic, and the ticket was reset
        assert np.array_equal(lp, lp2) and np.array_equal(g, g2)
        assert rel_err_logp(post.logPdfBatch(X), want_lp) < TOL


def test_stream_kernel_is_the_small_batch_default_and_rejects_large_batches():
    prob = linear_gaussian_problem(20, 500, seed=9)
    pri = [t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])]
    lik = [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")]
    auto = t.UnnormalisedPosterior(pri, lik)
    auto.logPdfGradBatch(typical_set_points(prob, 1, seed=1))
    assert auto.lastKernelPath == "stream"
    auto.logPdfGradBatch(typical_set_points(prob, 4, seed=1))
    assert auto.lastKernelPath == "stream_dmma"
    auto.logPdfGradBatch(typical_set_points(prob, 40, seed=1))
    assert auto.lastKernelPath in ("fused_dmma", "resident_dmma")
    forced = t.UnnormalisedPosterior(pri, lik, kernelPath=capi.KERNEL_STREAM)
    X = typical_set_points(prob, 27, seed=1)                   # 4 groups of <= 8 chains, one pass over A each
    lp, g = forced.logPdfGradBatch(X)
    want_lp, want_g = o.linear_gaussian_batch(prob["A"], prob["y"], prob["var"], X, prior_mean=prob["pm"], prior_var=prob["pvar"])
    assert forced.lastKernelPath == "stream_dmma" and rel_err_logp(lp, want_lp) < TOL and rel_err_grad(g, want_g) < TOL
    with pytest.raises(t.ThylacineError) as e:
        forced.logPdfGradBatch(typical_set_points(prob, 33, seed=1))
    assert e.value.status == capi.THY_ERR_UNSUPPORTED


@pytest.mark.parametrize("faithful", [True, False])
def test_stream_kernel_cauchy_multi_label_offset(faithful):
    rng = np.random.default_rng(21)
    n = 90
    A = {"gamma": rng.standard_normal((n, 2)), "alpha": rng.standard_normal((n, 3))}
    off, y = rng.standard_normal(n), rng.standard_normal(n)
    ci = 0.5 + rng.random(n)
    pri = [t.GaussianPrior.fromConfidenceIntervals("alpha", np.zeros(3), np.full(3, 4.0)),
           t.UniformPrior.fromBounds("beta", np.full(4, 9.0), np.full(4, -9.0)),
           t.GaussianPrior.fromConfidenceIntervals("gamma", np.ones(2), np.full(2, 2.0))]
    lik = t.CauchyLikelihood(t.LinearForwardModel.of(transform=A, vectorOffset=off), y, ci)
    post = t.UnnormalisedPosterior(pri, [lik], cauchyReferenceSign=faithful, kernelPath=capi.KERNEL_STREAM)
    ref = o.Posterior([o.gaussian_prior_from_ci("alpha", np.zeros(3), np.full(3, 4.0)),
                       o.uniform_prior_from_bounds("beta", np.full(4, 9.0), np.full(4, -9.0)),
                       o.gaussian_prior_from_ci("gamma", np.ones(2), np.full(2, 2.0))],
                      [o.cauchy_likelihood(o.LinearForwardModel(A, off), y, ci, reference_faithful=faithful)])
    X = rng.standard_normal((5, 9))
    lp, g = post.logPdfGradBatch(X)
    assert post.lastKernelPath == "stream_dmma"
    for b in range(5):
        assert lp[b] == pytest.approx(ref.log_pdf(X[b]), rel=TOL)
        assert rel_err_grad(g[b][None], ref.log_pdf_gradient(X[b])[None]) < TOL


@pytest.mark.parametrize("B", [2, 4, 8])
def test_stream_row_group_kernel_on_a_million_rows_repeated_launches_bit_identical(B):
    """100-d / 10^6 observations / 2, 4, 8 chains runs the DMMA row-group kernel (every CTA wraps its
    ring ~26 times), agrees with the oracle, and 1000 repeated launches are bit-identical (a stage-reuse race would show
    up as a differing partial sum or a stall; the test is bounded by pytest-timeout)."""
    import torch
    d, n = 100, 1_000_000
    rng = np.random.default_rng(4242)
    A = rng.standard_normal((n, d)) / np.sqrt(d)
    theta = rng.standard_normal(d)
    y = A @ theta + 0.1 * rng.standard_normal(n)
    ci = np.full(n, 0.2)
    pm, pci = np.zeros(d), np.full(d, 20.0)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", pm, pci)],
                                   [t.GaussianLinearLikelihood.of(A, y, ci, "theta")])
    X = theta[None, :] + 0.002 * rng.standard_normal((B, d))
    want_lp, want_g = o.linear_gaussian_batch(A, y, o.ci_to_variance(ci), X, prior_mean=pm, prior_var=o.ci_to_variance(pci))
    lp, g = post.logPdfGradBatch(X)
    assert post.lastKernelPath == "stream_dmma"
    assert rel_err_logp(lp, want_lp) < TOL and rel_err_grad(g, want_g) < TOL
    Xd = torch.from_numpy(X).cuda()
    lpd = torch.empty(B, dtype=torch.float64, device="cuda")
    gd = torch.empty(B, d, dtype=torch.float64, device="cuda")
    post.logPdfGradBatchDevice(B, Xd.data_ptr(), lpd.data_ptr(), gd.data_ptr())
    torch.cuda.synchronize()
    lp0, g0 = lpd.clone(), gd.clone()
    assert np.array_equal(lp0.cpu().numpy(), lp) and np.array_equal(g0.cpu().numpy(), g)
    bad = torch.zeros((), dtype=torch.int64, device="cuda")
    for _ in range(1000):
        post.logPdfGradBatchDevice(B, Xd.data_ptr(), lpd.data_ptr(), gd.data_ptr())
        bad += (lpd != lp0).sum() + (gd != g0).sum()
    torch.cuda.synchronize()
    assert int(bad.item()) == 0
    post.close()


# ---- wide DMMA GEMM path: d > 128 and / or non-linear links on the tensor cores -----------------------------
WIDE_SHAPES = [(200, 300, 150), (129, 129, 64), (1000, 517, 130), (333, 1000, 257), (7, 60, 33), (64, 5, 40), (100, 1000, 129)]


@pytest.mark.parametrize("d,n,B", WIDE_SHAPES)
def test_wide_path_linear_gaussian(d, n, B):
    prob = linear_gaussian_problem(d, n, seed=500 + d + n)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                   [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")],
                                   kernelPath=capi.KERNEL_WIDE)
    rng = np.random.default_rng(B)
    for X in (typical_set_points(prob, B, seed=B) if n >= d else rng.standard_normal((B, d)),
              prob["pm"] + 10.0 * rng.standard_normal((B, d))):
        want_lp, want_g = o.linear_gaussian_batch(prob["A"], prob["y"], prob["var"], X, prior_mean=prob["pm"],
                                                  prior_var=prob["pvar"])
        lp, g = post.logPdfGradBatch(X)
        assert post.lastKernelPath == "wide_dmma"
        assert rel_err_logp(lp, want_lp) < TOL
        assert rel_err_grad(g, want_g) < TOL
        assert rel_err_logp(post.logPdfBatch(X), want_lp) < TOL           # value only: W is never written
        lp2, g2 = post.logPdfGradBatch(X)
        assert np.array_equal(lp, lp2) and np.array_equal(g, g2)          # deterministic


def test_wide_path_is_the_default_beyond_128_columns():
    prob = linear_gaussian_problem(160, 400, seed=4)
    post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("theta", prob["pm"], prob["pci"])],
                                   [t.GaussianLinearLikelihood.of(prob["A"], prob["y"], prob["ci"], "theta")])
    post.logPdfGradBatch(typical_set_points(prob, 64, seed=1))
    assert post.lastKernelPath == "wide_dmma"
    post.logPdfGradBatch(typical_set_points(prob, 3, seed=1))
    assert post.lastKernelPath == "generic"


@pytest.mark.parametrize("link", ["tanh", "softplus", "identity"])
@pytest.mark.parametrize("faithful", [True, False])
def test_wide_path_cauchy_links_multi_label_offset(link, faithful):
    rng = np.random.default_rng(31)
    n, B = 210, 140
    dims = {"alpha": 70, "beta": 40, "gamma": 90}
    A = {"gamma": rng.standard_normal((n, 90)) / 10, "alpha": rng.standard_normal((n, 70)) / 10}   # skips beta
    off = 0.1 * rng.standard_normal(n)
    fn = {"tanh": np.tanh, "softplus": lambda z: np.logaddexp(0, z), "identity": lambda z: z}[link]
    dfn = {"tanh": lambda z: 1 - np.tanh(z) ** 2, "softplus": lambda z: 1 / (1 + np.exp(-z)), "identity": np.ones_like}[link]
    y = fn(rng.standard_normal(n)) + 0.05 * rng.standard_normal(n)
    ci = rng.uniform(1.5, 2.5, n)      # keeps the reference's det / Gamma normaliser finite at n = 210
    pri = [t.GaussianPrior.fromConfidenceIntervals(k, np.zeros(w), np.full(w, 4.0)) for k, w in dims.items()]
    fm = (t.LinearForwardModel.of(transform=A, vectorOffset=off) if link == "identity"
          else t.NonLinearForwardModel.of(link, A, vectorOffset=off))
    post = t.UnnormalisedPosterior(pri, [t.CauchyLikelihood(fm, y, ci)], cauchyReferenceSign=faithful,
                                   kernelPath=capi.KERNEL_WIDE)
    Acat = np.concatenate([A["alpha"], A["gamma"]], axis=1)

    def pre(p):
        return Acat @ np.concatenate([p["alpha"], p["gamma"]]) + off
    ofm = o.NonLinearForwardModel(lambda p: fn(pre(p)), None, dims, n,
                                  jacobian=lambda p: {"alpha": dfn(pre(p))[:, None] * A["alpha"],
                                                      "beta": np.zeros((n, 40)),
                                                      "gamma": dfn(pre(p))[:, None] * A["gamma"]})
    ref = o.Posterior([o.gaussian_prior_from_ci(k, np.zeros(w), np.full(w, 4.0)) for k, w in dims.items()],
                      [o.cauchy_likelihood(ofm, y, ci, faithful)])
    X = 0.5 * rng.standard_normal((B, 200))
    lp, g = post.logPdfGradBatch(X)
    assert post.lastKernelPath == "wide_dmma"
    for b in range(0, B, 7):
        assert lp[b] == pytest.approx(ref.log_pdf(X[b]), rel=TOL)
        assert rel_err_grad(g[b][None], ref.log_pdf_gradient(X[b])[None]) < TOL


# ---- single-kernel evaluation: priors folded into the fused kernel's output step -------------------------------------
@pytest.mark.parametrize("B", [1, 63, 64, 700, 4096])
def test_fused_prior_single_kernel_matches_the_separate_prior_kernel_and_the_oracle(B):
    import torch
    rng = np.random.default_rng(B)
    da, db, n = 5, 9, 300
    A = {"a": rng.standard_normal((n, da)), "b": rng.standard_normal((n, db))}
    y = rng.standard_normal(n)
    ci = rng.uniform(0.2, 1.0, n)
    priors = [t.UniformPrior.fromBounds("a", np.full(da, 2.0), np.full(da, -2.0)),
              t.GaussianPrior.fromConfidenceIntervals("b", rng.standard_normal(db), rng.uniform(1, 3, db))]
    lik = [t.GaussianLikelihood(t.LinearForwardModel.of(transform=A), y, ci)]
    X = rng.uniform(-1.9, 1.9, (B, da + db))
    X[B // 2, 2] = 2.0                                                         # on the open upper bound: outside
    outs = []
    for fuse in (True, False):
        post = t.UnnormalisedPosterior(priors, lik, kernelPath=capi.KERNEL_FUSED_DMMA, fusePrior=2 if fuse else 0)
        launches0 = t.kernelLaunchCount()
        lp, g = post.logPdfGradBatch(X)                                        # pageable numpy buffers: staged copies
        launches = t.kernelLaunchCount() - launches0
        assert launches == (1 if fuse else 2) * (1 if B < 2048 else 3)
        xp = torch.from_numpy(X).pin_memory()                                  # pinned buffers: the same chunk plan
        lpp, gp = torch.empty(B, dtype=torch.float64).pin_memory(), torch.empty(B, da + db, dtype=torch.float64).pin_memory()
        import ctypes
        dp = ctypes.POINTER(ctypes.c_double)
        capi.check(t.lib().thy_logpdf_grad_batch(post.handle, B, ctypes.cast(xp.data_ptr(), dp), ctypes.cast(lpp.data_ptr(), dp),
                                                 ctypes.cast(gp.data_ptr(), dp)))
        assert np.array_equal(lpp.numpy(), lp) and np.array_equal(gp.numpy(), g)
        lv = post.logPdfBatch(X)                                               # value only
        assert np.array_equal(lv, lp)
        outs.append((lp, g))
        post.close()
    (lf, gf), (ls, gs) = outs
    assert lf[B // 2] == -math.inf and ls[B // 2] == -math.inf
    ok = np.isfinite(ls)
    assert np.max(np.abs(lf[ok] - ls[ok]) / np.maximum(1.0, np.abs(ls[ok]))) < 1e-13 if ok.any() else True
    assert rel_err_grad(gf, gs) < 1e-13
    ref = o.Posterior([o.uniform_prior_from_bounds("a", np.full(da, 2.0), np.full(da, -2.0)),
                       o.gaussian_prior_from_ci("b", priors[1].a, priors[1].b)],
                      [o.gaussian_likelihood(o.LinearForwardModel(A), y, ci)])
    for b in range(0, B, max(1, B // 7)):
        want = ref.log_pdf(X[b])
        if math.isinf(want):
            assert lf[b] == -math.inf
        else:
            assert lf[b] == pytest.approx(want, rel=TOL)
        assert rel_err_grad(gf[b][None], ref.log_pdf_gradient(X[b])[None]) < TOL


@pytest.mark.parametrize("d,n,B", [(20, 64, 16), (200, 256, 96), (7, 5, 3), (131, 70, 40)])
def test_fd_literal_prefix_kernel_is_bit_identical_to_the_plain_literal_kernel(d, n, B):
    """The perturbed dot products start from the unperturbed prefix sums (k_fd_adjoint_prefix): the same fma sequence,
    so the gradient must equal the plain literal kernel's BIT FOR BIT (fdIncremental=2 selects the plain kernel)."""
    rng = np.random.default_rng(d + n)
    A = rng.standard_normal((n, d)) / np.sqrt(d)
    off = 0.1 * rng.standard_normal(n)
    y = np.tanh(A @ rng.standard_normal(d) + off) + 0.05 * rng.standard_normal(n)
    X = 0.5 * rng.standard_normal((B, d))
    outs = []
    for mode in (0, 2):
        fm = t.NonLinearForwardModel.of("tanh", {"w": A}, differential=1e-6, vectorOffset=off)
        post = t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("w", np.zeros(d), np.full(d, 4.0))],
                                       [t.CauchyLikelihood(fm, y, np.full(n, 0.1))], cauchyReferenceSign=False, fdIncremental=mode)
        outs.append(post.logPdfGradBatch(X))
        post.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.all(np.isfinite(outs[0][1])) and np.max(np.abs(outs[0][1])) > 0
