// Gaussian collapse of the linear-Gaussian likelihood terms, and the analytic Gaussian posterior on the device.
//   /root/reference/src/main/scala/thylacine/model/components/posterior/GaussianAnalyticPosterior.scala:141-169
//     Lambda = Sigma_p^-1 + A^T Sigma_d^-1 A,  mean = Lambda^-1 (Sigma_p^-1 mu_p + A^T Sigma_d^-1 y)
// SURVEY.md section 8(d) "reduced-flop form" / section 8(f) row 2.
//
// For GaussianLinearLikelihood terms the summed log-likelihood is exactly quadratic:
//     sum_l logL_l(x) = C - 1/2 (x^T G x - 2 b^T x + S),   G = sum A^T W A,  b = sum A^T W (y - offset).
// G and b come from the SAME device kernels that evaluate the likelihood (forward + adjoint on the fp64 tensor cores):
// the adjoint of the homogeneous problem (y = offset = 0) at the unit vectors e_j is -G e_j, with no cancellation
// against b.  With G = L L^T (Cholesky, host, d x d) the d-observation likelihood
//     A' = L^T,  y' = L^-1 b,  sigma' = 1,  log_const' = logL(x^) + 1/2 |y' - L^T x^|^2
// has the same value and gradient as the original terms at every x, for 4 d^2 instead of 4 n d flops per evaluation.
// The collapsed model keeps the source's priors (any kind) and label layout, so every engine runs on it unchanged.
#include "kernels.h"
#include <algorithm>
#include <cmath>

namespace thy {

// G (d x d, row-major, symmetrised), b (d) and logL(0) of all likelihood terms, computed on the device.
static thy_status likelihood_gram(thy_model_s* m, std::vector<double>& G, std::vector<double>& b, double* logl0) {
  const int d = m->d;
  for (auto& own : m->lik_dev) {
    const LikDev& lk = own->view;
    if (lk.distribution != THY_DIST_GAUSSIAN || lk.link != THY_LINK_IDENTITY || lk.jacobian != THY_JAC_CONSTANT)
      return fail(THY_ERR_UNSUPPORTED, "only GaussianLinearLikelihood terms (Gaussian observations, linear forward model) collapse");
  }
  if (m->liks.size() != m->lik_dev.size()) return fail(THY_ERR_STATE, "likelihood records out of step");
  cudaStream_t s = m->stream;
  const size_t nx = (size_t)(d + 1) * d;
  std::vector<double> hX(nx, 0.0);
  for (int j = 0; j < d; ++j) hX[(size_t)(1 + j) * d + j] = 1.0;
  DevBuf<double> X, Gd, LP;
  THY_TRY(X.upload(hX, s));
  THY_TRY(Gd.reserve(nx));
  THY_TRY(LP.reserve((size_t)d + 1));
  THY_CUDA_CHECK(cudaMemsetAsync(Gd.ptr, 0, nx * sizeof(double), s));
  THY_CUDA_CHECK(cudaMemsetAsync(LP.ptr, 0, ((size_t)d + 1) * sizeof(double), s));
  for (size_t l = 0; l < m->lik_dev.size(); ++l) {
    LikDevOwner& own = *m->lik_dev[l];
    const LikelihoodHost& L = m->liks[l];
    // row 0: the real observations at x = 0  ->  b and C - S/2
    THY_TRY(launch_likelihood(m, own, 1, X.ptr, LP.ptr, Gd.ptr, -1.0, s));
    // rows 1..d: the homogeneous problem at the unit vectors  ->  -G e_j
    std::vector<double2> y0((size_t)own.view.n_pad, make_double2(0.0, 0.0));
    for (int i = 0; i < own.view.n; ++i) y0[i] = make_double2(0.0, 1.0 / L.unc[i]);
    std::vector<double> off0((size_t)own.view.n_pad, 0.0);
    DevBuf<double2> y0d;
    DevBuf<double> off0d;
    THY_TRY(y0d.upload(y0, s));
    THY_TRY(off0d.upload(off0, s));
    const LikDev saved = own.view;
    own.view.yiv = own.view.yiv_lin = y0d.ptr;
    own.view.off = off0d.ptr;
    thy_status st = launch_likelihood(m, own, d, X.ptr + d, LP.ptr + 1, Gd.ptr + d, -1.0, s);
    own.view = saved;
    THY_TRY(st);
    THY_CUDA_CHECK(cudaStreamSynchronize(s));  // y0d / off0d die at the end of this iteration
  }
  std::vector<double> hG(nx), hLP((size_t)d + 1);
  THY_CUDA_CHECK(cudaMemcpyAsync(hG.data(), Gd.ptr, nx * sizeof(double), cudaMemcpyDeviceToHost, s));
  THY_CUDA_CHECK(cudaMemcpyAsync(hLP.data(), LP.ptr, ((size_t)d + 1) * sizeof(double), cudaMemcpyDeviceToHost, s));
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  b.assign(hG.begin(), hG.begin() + d);
  G.assign((size_t)d * d, 0.0);
  for (int j = 0; j < d; ++j)
    for (int k = 0; k <= j; ++k) {
      const double v = -0.5 * (hG[(size_t)(1 + j) * d + k] + hG[(size_t)(1 + k) * d + j]);
      G[(size_t)j * d + k] = G[(size_t)k * d + j] = v;
    }
  if (logl0) *logl0 = hLP[0];
  return THY_OK;
}

// Likelihood-only log-density at one host point.
static thy_status likelihood_logp_at(thy_model_s* m, const std::vector<double>& x, double* out) {
  cudaStream_t s = m->stream;
  DevBuf<double> X, LP;
  THY_TRY(X.upload(x, s));
  THY_TRY(LP.reserve(1));
  THY_CUDA_CHECK(cudaMemsetAsync(LP.ptr, 0, sizeof(double), s));
  for (auto& own : m->lik_dev) THY_TRY(launch_likelihood(m, *own, 1, X.ptr, LP.ptr, nullptr, -1.0, s));
  THY_CUDA_CHECK(cudaMemcpyAsync(out, LP.ptr, sizeof(double), cudaMemcpyDeviceToHost, s));
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  return THY_OK;
}

// y = L^-1 b (forward substitution) and x = L^-T y (back substitution), long-double accumulation.
static void chol_solve(const std::vector<double>& L, int n, const std::vector<double>& b, std::vector<double>& y,
                       std::vector<double>& x) {
  y.assign(n, 0.0);
  x.assign(n, 0.0);
  for (int i = 0; i < n; ++i) {
    long double s = b[i];
    for (int k = 0; k < i; ++k) s -= (long double)L[(size_t)i * n + k] * y[k];
    y[i] = (double)(s / L[(size_t)i * n + i]);
  }
  for (int i = n - 1; i >= 0; --i) {
    long double s = y[i];
    for (int k = i + 1; k < n; ++k) s -= (long double)L[(size_t)k * n + i] * x[k];
    x[i] = (double)(s / L[(size_t)i * n + i]);
  }
}

}  // namespace thy

using namespace thy;

extern "C" {

thy_status thy_model_collapse(thy_model_t m, thy_model_t* out) {
  thy::ErrScope _err_scope(m);
  if (!out) return fail(THY_ERR_INVALID_ARGUMENT, "null out pointer");
  *out = nullptr;
  if (!m || !m->finalized) return fail(THY_ERR_STATE, "model is not finalized");
  THY_TRY(require_device());
  if (m->lik_dev.empty()) return fail(THY_ERR_UNSUPPORTED, "the model has no likelihood term to collapse");
  const int d = m->d;
  std::vector<double> G, b;
  THY_TRY(likelihood_gram(m, G, b, nullptr));
  // columns any likelihood reads (a likelihood only touches its own labels, LinearForwardModel.scala:111-114)
  std::vector<int> U;
  for (auto& own : m->lik_dev) U.insert(U.end(), own->col_host.begin(), own->col_host.end());
  std::sort(U.begin(), U.end());
  U.erase(std::unique(U.begin(), U.end()), U.end());
  const int du = (int)U.size();
  std::vector<double> GU((size_t)du * du), bU(du), L;
  for (int i = 0; i < du; ++i) {
    bU[i] = b[U[i]];
    for (int j = 0; j < du; ++j) GU[(size_t)i * du + j] = G[(size_t)U[i] * d + U[j]];
  }
  if (!cholesky_lower(GU, du, L))
    return fail(THY_ERR_UNSUPPORTED,
                "A^T Sigma^-1 A is singular (fewer independent observations than parameters): nothing to collapse onto");
  std::vector<double> yp, xh;
  chol_solve(L, du, bU, yp, xh);
  std::vector<double> xfull(d, 0.0);
  for (int i = 0; i < du; ++i) xfull[U[i]] = xh[i];
  double lp_hat = 0.0;
  THY_TRY(likelihood_logp_at(m, xfull, &lp_hat));
  long double q = 0.0L;  // |y' - L^T x^|^2, zero up to rounding
  std::vector<double> At((size_t)du * du, 0.0);
  for (int i = 0; i < du; ++i) {
    long double z = 0.0L;
    for (int j = i; j < du; ++j) {
      At[(size_t)i * du + j] = L[(size_t)j * du + i];  // A' = L^T, upper triangular
      z += (long double)L[(size_t)j * du + i] * xh[j];
    }
    q += ((long double)yp[i] - z) * ((long double)yp[i] - z);
  }
  const double lc = (double)((long double)lp_hat + 0.5L * q);

  thy_model_t c = nullptr;
  THY_TRY(thy_model_create(&c));
  c->priors = m->priors;
  c->opt_cauchy_ref_sign = m->opt_cauchy_ref_sign;
  c->opt_fd_incremental = m->opt_fd_incremental;
  c->opt_fuse_prior = m->opt_fuse_prior;
  c->opt_kernel_path = m->opt_kernel_path == 4 ? 0 : m->opt_kernel_path;
  c->stream = (m->stream == m->own_stream) ? (cudaStream_t)0 : m->stream;   // an external stream is shared; an owned one is not
  std::vector<double> ones(du, 1.0);
  thy_status st = thy_model_finalize(c);
  if (st == THY_OK) {
    st = build_lik_dev(c, THY_DIST_GAUSSIAN, THY_LINK_IDENTITY, THY_JAC_CONSTANT, 0.0, du, U, {At.data()}, {du}, yp.data(),
                       ones.data(), nullptr, &lc);
  }
  if (st == THY_OK) {
    LikelihoodHost rec;  // observation record of the synthetic term (coefficients live on the device only)
    rec.distribution = THY_DIST_GAUSSIAN;
    rec.link = THY_LINK_IDENTITY;
    rec.jacobian = THY_JAC_CONSTANT;
    rec.n = du;
    rec.meas = yp;
    rec.unc = ones;
    c->liks.push_back(std::move(rec));
    // build_lik_dev uploads on stream 0, which is not ordered against an external non-blocking stream: wait for the device
    if (cudaDeviceSynchronize() != cudaSuccess) st = fail(THY_ERR_CUDA, "upload of the collapsed term failed");
  }
  if (st != THY_OK) {
    thy_model_destroy(c);
    return st;
  }
  *out = c;
  return THY_OK;
}

thy_status thy_model_gaussian_posterior(thy_model_t m, double* out_mean, double* out_covariance, double* out_precision,
                                        double* out_log_evidence, double* out_max_logpdf) {
  thy::ErrScope _err_scope(m);
  if (!m || !m->finalized) return fail(THY_ERR_STATE, "model is not finalized");
  THY_TRY(require_device());
  const int d = m->d;
  for (auto& p : m->priors)
    if (p.kind != PK_GAUSS_DIAG && p.kind != PK_GAUSS_FULL)
      return fail(THY_ERR_UNSUPPORTED, "GaussianAnalyticPosterior takes Gaussian priors (GaussianAnalyticPosterior.scala:39)");
  std::vector<double> Lam((size_t)d * d, 0.0), b(d, 0.0);
  if (!m->lik_dev.empty()) THY_TRY(likelihood_gram(m, Lam, b, nullptr));
  for (auto& p : m->priors) {
    if (p.kind == PK_GAUSS_DIAG) {
      for (int i = 0; i < p.dim; ++i) {
        Lam[(size_t)(p.offset + i) * d + p.offset + i] += 1.0 / p.b[i];
        b[p.offset + i] += p.a[i] / p.b[i];
      }
    } else {
      std::vector<double> Lc, P;
      if (!cholesky_lower(p.b, p.dim, Lc)) return fail(THY_ERR_INVALID_ARGUMENT, "covariance not positive definite");
      spd_inverse(Lc, p.dim, P);
      for (int i = 0; i < p.dim; ++i) {
        long double s = 0.0L;
        for (int j = 0; j < p.dim; ++j) {
          Lam[(size_t)(p.offset + i) * d + p.offset + j] += P[(size_t)i * p.dim + j];
          s += (long double)P[(size_t)i * p.dim + j] * p.a[j];
        }
        b[p.offset + i] += (double)s;
      }
    }
  }
  std::vector<double> L, y, mean;
  if (!cholesky_lower(Lam, d, L)) return fail(THY_ERR_UNSUPPORTED, "posterior precision is not positive definite");
  chol_solve(L, d, b, y, mean);
  if (out_mean) std::copy(mean.begin(), mean.end(), out_mean);
  if (out_precision) std::copy(Lam.begin(), Lam.end(), out_precision);
  if (out_covariance) {
    std::vector<double> C;
    spd_inverse(L, d, C);
    std::copy(C.begin(), C.end(), out_covariance);
  }
  if (out_log_evidence || out_max_logpdf) {
    // ln Z = ln p~(x^) + d/2 ln 2 pi - 1/2 ln det Lambda;  the normalised density peaks at -d/2 ln 2 pi + 1/2 ln det Lambda
    DevBuf<double> X, LP;
    THY_TRY(X.upload(mean, m->stream));
    THY_TRY(LP.reserve(1));
    THY_TRY(eval_batch_dev(m, 1, X.ptr, LP.ptr, nullptr, false));
    double lp = 0.0;
    THY_CUDA_CHECK(cudaMemcpyAsync(&lp, LP.ptr, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    THY_CUDA_CHECK(cudaStreamSynchronize(m->stream));
    long double half_logdet = 0.0L;
    for (int i = 0; i < d; ++i) half_logdet += std::log((long double)L[(size_t)i * d + i]);
    const long double gauss = 0.5L * d * std::log(2.0L * (long double)M_PI) - half_logdet;
    if (out_log_evidence) *out_log_evidence = (double)((long double)lp + gauss);
    if (out_max_logpdf) *out_max_logpdf = (double)(-gauss);
  }
  return THY_OK;
}

}  // extern "C"
