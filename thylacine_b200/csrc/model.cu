// Model construction: validation (what the reference asserts), label-sorted flattening, upload.
// Reference semantics followed here (paths relative to /root/reference/src/main/scala/thylacine/):
//   model/core/RecordedData.scala:48-72                 variance = Math.pow(ci/2, 2), CI > 0
//   model/components/posterior/Posterior.scala:40-48    flat order = prior labels sorted ascending
//   model/components/forwardmodel/LinearForwardModel.scala:87-101  column blocks merged in sorted label order
//   model/distributions/UniformDistribution.scala:36-39 upper > lower
//   model/components/posterior/UnnormalisedPosterior.scala:35-47   no duplicate labels
#include "kernels.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>

namespace thy {

static thread_local std::string t_last_error = "";
std::atomic<uint64_t> g_kernel_launches{0};

static thread_local thy_model_s* t_scope_model = nullptr;
ErrScope::ErrScope(thy_model_s* m) : prev(t_scope_model) { t_scope_model = m; }
ErrScope::~ErrScope() { t_scope_model = prev; }

void set_last_error(const std::string& msg) { t_last_error = msg; }
thy_status fail(thy_status code, const std::string& msg) {
  t_last_error = msg;
  if (t_scope_model) {
    std::lock_guard<std::mutex> lock(t_scope_model->err_mutex);
    t_scope_model->last_error = msg;
  }
  return code;
}

thy_status require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(THY_ERR_NO_DEVICE, "no CUDA device available: thylacine_b200 has no CPU fallback");
  }
  return THY_OK;
}

int sm_count() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  cached = n;
  return n;
}

// Lower Cholesky factor of a dense SPD matrix (row-major), in place into L; false if not SPD.
bool cholesky_lower(const std::vector<double>& S, int n, std::vector<double>& L) {
  L.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j <= i; ++j) {
      long double s = S[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= (long double)L[(size_t)i * n + k] * L[(size_t)j * n + k];
      if (i == j) {
        if (!(s > 0)) return false;
        L[(size_t)i * n + i] = (double)sqrtl(s);
      } else {
        L[(size_t)i * n + j] = (double)(s / L[(size_t)j * n + j]);
      }
    }
  }
  return true;
}

// Inverse of an SPD matrix from its Cholesky factor (long-double accumulation).
void spd_inverse(const std::vector<double>& L, int n, std::vector<double>& P) {
  std::vector<long double> Li((size_t)n * n, 0.0L);  // inverse of L (lower)
  for (int i = 0; i < n; ++i) {
    Li[(size_t)i * n + i] = 1.0L / L[(size_t)i * n + i];
    for (int j = 0; j < i; ++j) {
      long double s = 0;
      for (int k = j; k < i; ++k) s -= (long double)L[(size_t)i * n + k] * Li[(size_t)k * n + j];
      Li[(size_t)i * n + j] = s / L[(size_t)i * n + i];
    }
  }
  P.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      long double s = 0;
      for (int k = i; k < n; ++k) s += Li[(size_t)k * n + i] * Li[(size_t)k * n + j];
      P[(size_t)i * n + j] = P[(size_t)j * n + i] = (double)s;
    }
}

static thy_status check_model(thy_model_t m, bool want_finalized) {
  if (!m) return fail(THY_ERR_INVALID_ARGUMENT, "null model handle");
  if (m->poisoned) return fail(THY_ERR_STATE, "an earlier thy_model_finalize failed on this handle; destroy it and build a new model");
  if (want_finalized && !m->finalized) return fail(THY_ERR_STATE, "model is not finalized");
  if (!want_finalized && m->finalized) return fail(THY_ERR_STATE, "model is already finalized (immutable)");
  return THY_OK;
}

static thy_status check_label(thy_model_t m, const char* label, int dim) {
  if (!label || !*label) return fail(THY_ERR_INVALID_ARGUMENT, "empty label");
  if (dim <= 0) return fail(THY_ERR_INVALID_ARGUMENT, "dimension must be positive");
  for (auto& p : m->priors)
    if (p.label == label) return fail(THY_ERR_INVALID_ARGUMENT, std::string("duplicate prior label '") + label + "'");
  return THY_OK;
}

static thy_status ci_to_var(const double* ci, int n, std::vector<double>& var) {
  var.resize(n);
  for (int i = 0; i < n; ++i) {
    if (!(ci[i] > 0)) return fail(THY_ERR_INVALID_ARGUMENT, "confidence intervals must be > 0 (RecordedData.scala:57)");
    var[i] = std::pow(ci[i] / 2, 2);  // Math.pow(ci / 2, 2)
  }
  return THY_OK;
}

}  // namespace thy

namespace thy {

// Uploads one likelihood term: the column blocks (ascending label order, already resolved to flat columns `cols`)
// are merged into the padded coefficient matrix, observations into (y, 1/var) pairs.  `log_const_override` replaces
// the distribution's own normaliser (used by the collapsed Gaussian term, collapse.cu).
thy_status build_lik_dev(thy_model_s* m, int distribution, int link, int jacobian, double differential, int n,
                         const std::vector<int>& cols, const std::vector<const double*>& blocks,
                         const std::vector<int>& widths, const double* meas, const double* unc, const double* offset,
                         const double* log_const_override) {
  const double LOG_2PI = std::log(2.0 * M_PI);
  auto own = std::make_unique<LikDevOwner>();
  const int dl = (int)cols.size();
  const int n_pad = (int)round_up(n, 32);
  // row pitch = 4 (mod 8) doubles: conflict-free DMMA fragment loads (kernels_fused_dmma.cu)
  const int nt = fused_nt_for(dl);
  const int lds = (nt > 0 ? 8 * nt : (int)round_up(dl, 8)) + 4;
  std::vector<double> A((size_t)n_pad * lds, 0.0);
  own->col_host = cols;
  int c0 = 0;
  for (size_t b = 0; b < blocks.size(); ++b) {
    const int w = widths[b];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < w; ++j) A[(size_t)i * lds + c0 + j] = blocks[b][(size_t)i * w + j];
    c0 += w;
  }
  own->contiguous = true;
  for (int k = 1; k < dl; ++k)
    if (own->col_host[k] != own->col_host[0] + k) own->contiguous = false;
  std::vector<double2> yiv(n_pad, make_double2(0.0, 0.0)), yiv_lin(n_pad, make_double2(0.0, 0.0));
  std::vector<double> offv(n_pad, 0.0);
  double lc = 0.0;
  if (distribution == THY_DIST_UNIFORM) {
    for (int i = 0; i < n; ++i) {
      yiv[i] = make_double2(meas[i], unc[i]);  // (upper, lower)
      lc -= std::log(meas[i] - unc[i]);
    }
    // padded rows must always pass the bounds test on f = 0
    for (int i = n; i < n_pad; ++i) yiv[i] = make_double2(1.0, -1.0);
  } else {
    double logdet = 0.0, lognorm = n / 2.0 * LOG_2PI;
    for (int i = 0; i < n; ++i) {
      const double o = offset ? offset[i] : 0.0;
      yiv[i] = make_double2(meas[i], 1.0 / unc[i]);
      yiv_lin[i] = make_double2(meas[i] - o, 1.0 / unc[i]);
      logdet += std::log(unc[i]);
      lognorm += std::log(std::sqrt(unc[i]));
    }
    if (distribution == THY_DIST_GAUSSIAN)
      lc = -lognorm;
    else
      lc = std::lgamma((1 + n) / 2.0) - std::lgamma(0.5) - n / 2.0 * std::log(M_PI) - logdet / 2.0;
  }
  if (log_const_override) lc = *log_const_override;
  if (offset)
    for (int i = 0; i < n; ++i) offv[i] = offset[i];
  THY_TRY(own->A.upload(A));
  THY_TRY(own->yiv.upload(yiv));
  THY_TRY(own->yiv_lin.upload(yiv_lin));
  THY_TRY(own->off.upload(offv));
  THY_TRY(own->col.upload(own->col_host));
  LikDev& v = own->view;
  v.distribution = distribution;
  v.link = link;
  v.jacobian = jacobian;
  v.n = n;
  v.dl = dl;
  v.n_pad = n_pad;
  v.lds = lds;
  v.A = own->A.ptr;
  v.yiv = own->yiv.ptr;
  v.yiv_lin = own->yiv_lin.ptr;
  v.off = own->off.ptr;
  v.col = own->col.ptr;
  v.log_const = lc;
  v.differential = differential;
  v.cauchy_sign = m->opt_cauchy_ref_sign ? 1.0 : -1.0;
  m->lik_dev.push_back(std::move(own));
  return THY_OK;
}

}  // namespace thy

using namespace thy;

extern "C" {

const char* thy_version(void) { return "thylacine_b200 0.1 (sm_100a)"; }
const char* thy_last_error(void) { return thy::t_last_error.c_str(); }
thy_status thy_model_last_error(thy_model_t m, char* out, size_t capacity) {
  if (!m || !out || capacity == 0) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  std::lock_guard<std::mutex> lock(m->err_mutex);
  std::snprintf(out, capacity, "%s", m->last_error.c_str());
  return THY_OK;
}
uint64_t thy_kernel_launch_count(void) { return g_kernel_launches.load(); }

thy_status thy_model_create(thy_model_t* out) {
  if (!out) return fail(THY_ERR_INVALID_ARGUMENT, "null out pointer");
  *out = new (std::nothrow) thy_model_s();
  if (!*out) return fail(THY_ERR_INVALID_ARGUMENT, "allocation failed");
  return THY_OK;
}

thy_status thy_model_destroy(thy_model_t m) {
  if (m) {
    for (auto& set : m->async_set) {
      if (set.d2h_done) cudaEventSynchronize(set.d2h_done);   // an outstanding async call still owns the staging buffers
      if (set.h2d_done) cudaEventDestroy(set.h2d_done);
      if (set.compute_done) cudaEventDestroy(set.compute_done);
      if (set.d2h_done) cudaEventDestroy(set.d2h_done);
    }
  }
  if (m && m->own_stream) cudaStreamDestroy(m->own_stream);
  if (m && m->h2d_stream) {
    cudaStreamDestroy(m->h2d_stream);
    cudaStreamDestroy(m->d2h_stream);
    cudaEventDestroy(m->ev_start);
    for (int i = 0; i < 8; ++i) {
      cudaEventDestroy(m->ev_h2d[i]);
      cudaEventDestroy(m->ev_done[i]);
    }
  }
  delete m;
  return THY_OK;
}

thy_status thy_model_add_gaussian_prior_ci(thy_model_t m, const char* label, int dim, const double* values,
                                           const double* ci) {
  thy::ErrScope _err_scope(m);
  THY_TRY(check_model(m, false));
  THY_TRY(check_label(m, label, dim));
  if (!values || !ci) return fail(THY_ERR_INVALID_ARGUMENT, "null array");
  PriorHost p;
  p.label = label;
  p.dim = dim;
  p.kind = PK_GAUSS_DIAG;
  p.a.assign(values, values + dim);
  THY_TRY(ci_to_var(ci, dim, p.b));
  m->priors.push_back(std::move(p));
  return THY_OK;
}

thy_status thy_model_add_gaussian_prior_cov(thy_model_t m, const char* label, int dim, const double* values,
                                            const double* cov) {
  thy::ErrScope _err_scope(m);
  THY_TRY(check_model(m, false));
  THY_TRY(check_label(m, label, dim));
  if (!values || !cov) return fail(THY_ERR_INVALID_ARGUMENT, "null array");
  PriorHost p;
  p.label = label;
  p.dim = dim;
  p.kind = PK_GAUSS_FULL;
  p.a.assign(values, values + dim);
  p.b.assign(cov, cov + (size_t)dim * dim);
  for (int i = 0; i < dim; ++i)
    for (int j = 0; j < i; ++j) {
      double x = p.b[(size_t)i * dim + j], y = p.b[(size_t)j * dim + i];
      if (std::fabs(x - y) > 1e-12 * (std::fabs(x) + std::fabs(y)) + 1e-300)
        return fail(THY_ERR_INVALID_ARGUMENT, "covariance matrix is not symmetric");
    }
  std::vector<double> L;
  if (!cholesky_lower(p.b, dim, L)) return fail(THY_ERR_INVALID_ARGUMENT, "covariance matrix is not positive definite");
  m->priors.push_back(std::move(p));
  return THY_OK;
}

thy_status thy_model_add_uniform_prior(thy_model_t m, const char* label, int dim, const double* max_bounds,
                                       const double* min_bounds) {
  thy::ErrScope _err_scope(m);
  THY_TRY(check_model(m, false));
  THY_TRY(check_label(m, label, dim));
  if (!max_bounds || !min_bounds) return fail(THY_ERR_INVALID_ARGUMENT, "null array");
  for (int i = 0; i < dim; ++i)
    if (!(max_bounds[i] > min_bounds[i]))
      return fail(THY_ERR_INVALID_ARGUMENT, "upper bound must exceed lower bound (UniformDistribution.scala:36-39)");
  PriorHost p;
  p.label = label;
  p.dim = dim;
  p.kind = PK_UNIFORM;
  p.a.assign(max_bounds, max_bounds + dim);
  p.b.assign(min_bounds, min_bounds + dim);
  m->priors.push_back(std::move(p));
  return THY_OK;
}

thy_status thy_model_add_cauchy_prior(thy_model_t m, const char* label, int dim, const double* values, const double* ci) {
  thy::ErrScope _err_scope(m);
  THY_TRY(check_model(m, false));
  THY_TRY(check_label(m, label, dim));
  if (!values || !ci) return fail(THY_ERR_INVALID_ARGUMENT, "null array");
  PriorHost p;
  p.label = label;
  p.dim = dim;
  p.kind = PK_CAUCHY;
  p.a.assign(values, values + dim);
  THY_TRY(ci_to_var(ci, dim, p.b));
  m->priors.push_back(std::move(p));
  return THY_OK;
}

thy_status thy_model_add_likelihood(thy_model_t m, const thy_likelihood_desc* d) {
  thy::ErrScope _err_scope(m);
  THY_TRY(check_model(m, false));
  if (!d) return fail(THY_ERR_INVALID_ARGUMENT, "null descriptor");
  if (d->n <= 0 || d->nblocks <= 0) return fail(THY_ERR_INVALID_ARGUMENT, "n and nblocks must be positive");
  if (d->distribution < THY_DIST_GAUSSIAN || d->distribution > THY_DIST_UNIFORM)
    return fail(THY_ERR_INVALID_ARGUMENT, "unknown distribution");
  const bool user_link = d->link >= THY_LINK_USER_BASE && d->link < THY_LINK_USER_BASE + (int)m->user_links.size();
  if (!user_link && (d->link < THY_LINK_IDENTITY || d->link > THY_LINK_SIGMOID))
    return fail(THY_ERR_INVALID_ARGUMENT, "unknown link");
  if (user_link) {
    if (d->distribution == THY_DIST_UNIFORM)
      return fail(THY_ERR_UNSUPPORTED, "user-defined links take Gaussian or Cauchy observations");
    if (d->jacobian == THY_JAC_CONSTANT)
      return fail(THY_ERR_INVALID_ARGUMENT, "a constant Jacobian requires the identity link (LinearForwardModel)");
    if (d->jacobian == THY_JAC_ANALYTIC && !m->user_links[d->link - THY_LINK_USER_BASE]->has_deriv)
      return fail(THY_ERR_INVALID_ARGUMENT, "THY_JAC_ANALYTIC needs a link registered with a derivative expression");
    if (d->jacobian == THY_JAC_FINITE_DIFFERENCE && !(d->differential > 0))
      return fail(THY_ERR_INVALID_ARGUMENT, "finite differences need a positive differential");
  }
  if (d->jacobian < THY_JAC_CONSTANT || d->jacobian > THY_JAC_FINITE_DIFFERENCE)
    return fail(THY_ERR_INVALID_ARGUMENT, "unknown jacobian mode");
  if (d->jacobian == THY_JAC_CONSTANT && d->link != THY_LINK_IDENTITY)
    return fail(THY_ERR_INVALID_ARGUMENT, "a constant Jacobian requires the identity link (LinearForwardModel)");
  if (d->jacobian == THY_JAC_FINITE_DIFFERENCE && !(d->differential > 0))
    return fail(THY_ERR_INVALID_ARGUMENT, "finite-difference differential must be > 0");
  if (!d->labels || !d->dims || !d->blocks || !d->measurements || !d->uncertainties)
    return fail(THY_ERR_INVALID_ARGUMENT, "null array in descriptor");
  LikelihoodHost L;
  L.distribution = d->distribution;
  L.link = d->link;
  L.jacobian = d->jacobian;
  L.n = d->n;
  L.differential = d->differential;
  for (int b = 0; b < d->nblocks; ++b) {
    if (!d->labels[b] || !d->blocks[b] || d->dims[b] <= 0) return fail(THY_ERR_INVALID_ARGUMENT, "bad coefficient block");
    std::string lab = d->labels[b];
    for (auto& l : L.labels)
      if (l == lab) return fail(THY_ERR_INVALID_ARGUMENT, "duplicate label inside one forward model");
    const PriorHost* pr = nullptr;
    for (auto& p : m->priors)
      if (p.label == lab) pr = &p;
    if (!pr) return fail(THY_ERR_UNKNOWN_LABEL, "likelihood refers to unknown prior label '" + lab + "'");
    if (pr->dim != d->dims[b])
      return fail(THY_ERR_INVALID_ARGUMENT, "coefficient block width differs from the prior dimension of '" + lab + "'");
    L.labels.push_back(lab);
    L.dims.push_back(d->dims[b]);
    L.blocks.emplace_back(d->blocks[b], d->blocks[b] + (size_t)d->n * d->dims[b]);
  }
  if (d->offset) {
    L.has_offset = true;
    L.offset.assign(d->offset, d->offset + d->n);
  }
  L.meas.assign(d->measurements, d->measurements + d->n);
  if (d->distribution == THY_DIST_UNIFORM) {
    L.unc.assign(d->uncertainties, d->uncertainties + d->n);
    for (int i = 0; i < d->n; ++i)
      if (!(L.meas[i] > L.unc[i]))
        return fail(THY_ERR_INVALID_ARGUMENT, "upper bound must exceed lower bound (UniformDistribution.scala:36-39)");
  } else {
    THY_TRY(ci_to_var(d->uncertainties, d->n, L.unc));  // unc now holds variances
  }
  m->liks.push_back(std::move(L));
  return THY_OK;
}

thy_status thy_model_register_link(thy_model_t m, const char* eval_expr, const char* deriv_expr, int* out_link) {
  thy::ErrScope _err_scope(m);
  THY_TRY(check_model(m, false));
  if (!out_link) return fail(THY_ERR_INVALID_ARGUMENT, "null out pointer");
  std::unique_ptr<UserLink> link;
  THY_TRY(user_link_compile(eval_expr, deriv_expr, &link));
  m->user_links.push_back(std::move(link));
  *out_link = THY_LINK_USER_BASE + (int)m->user_links.size() - 1;
  return THY_OK;
}

thy_status thy_model_set_option(thy_model_t m, int option, int value) {
  thy::ErrScope _err_scope(m);
  THY_TRY(check_model(m, false));
  switch (option) {
    case THY_OPT_CAUCHY_REFERENCE_SIGN: m->opt_cauchy_ref_sign = value ? 1 : 0; return THY_OK;
    case THY_OPT_FD_INCREMENTAL:
      if (value < 0 || value > 2) return fail(THY_ERR_INVALID_ARGUMENT, "finite-difference mode must be 0, 1 or 2");
      m->opt_fd_incremental = value;
      return THY_OK;
    case THY_OPT_KERNEL_PATH:
      if (value < 0 || value > 5) return fail(THY_ERR_INVALID_ARGUMENT, "kernel path must be 0..5");
      m->opt_kernel_path = value;
      return THY_OK;
    case THY_OPT_FUSE_PRIOR:
      if (value < 0 || value > 2) return fail(THY_ERR_INVALID_ARGUMENT, "fuse-prior option must be 0, 1 or 2");
      m->opt_fuse_prior = value;
      return THY_OK;
    default: return fail(THY_ERR_INVALID_ARGUMENT, "unknown option");
  }
}

// A finalize that fails half-way (non-SPD covariance, unknown label, failed upload) leaves host vectors and device
// buffers partially filled and the coefficient blocks of earlier likelihoods already released: the handle is marked
// poisoned and every later call on it reports THY_ERR_STATE (the caller destroys it and builds a new model).
static thy_status finalize_impl(thy_model_t m);
thy_status thy_model_finalize(thy_model_t m) {
  thy::ErrScope _err_scope(m);
  THY_TRY(check_model(m, false));
  const thy_status st = finalize_impl(m);
  if (st != THY_OK && st != THY_ERR_NO_DEVICE) m->poisoned = true;
  return st;
}
static thy_status finalize_impl(thy_model_t m) {
  m->h_chol_off.clear();
  m->pr_precisions.clear();
  m->lik_dev.clear();
  m->layout.clear();
  if (m->priors.empty()) return fail(THY_ERR_INVALID_ARGUMENT, "a posterior needs at least one prior");
  THY_TRY(require_device());
  THY_CUDA_CHECK(cudaGetDevice(&m->device));

  // flat layout: labels ascending by String.compareTo (UTF-16 code units; == byte order for ASCII labels)
  std::sort(m->priors.begin(), m->priors.end(), [](const PriorHost& a, const PriorHost& b) { return a.label < b.label; });
  int off = 0;
  for (auto& p : m->priors) {
    p.offset = off;
    m->layout[p.label] = {off, p.dim};
    off += p.dim;
  }
  m->d = off;
  const int d = m->d;
  const double LOG_2PI = std::log(2.0 * M_PI);

  // ---- priors ----
  m->h_kind.assign(d, PK_GAUSS_DIAG);
  m->h_p0.assign(d, 0.0);
  m->h_p1.assign(d, 0.0);
  double log_const = 0.0;
  m->h_special.clear();
  std::vector<double> chol_all;
  for (auto& p : m->priors) {
    for (int i = 0; i < p.dim; ++i) m->h_kind[p.offset + i] = p.kind;
    if (p.kind == PK_GAUSS_DIAG) {
      double lognorm = p.dim / 2.0 * LOG_2PI;
      for (int i = 0; i < p.dim; ++i) {
        m->h_p0[p.offset + i] = p.a[i];
        m->h_p1[p.offset + i] = 1.0 / p.b[i];
        lognorm += std::log(std::sqrt(p.b[i]));  // sum(log(diag(cholesky)))
      }
      log_const -= lognorm;
    } else if (p.kind == PK_UNIFORM) {
      double neg_log_vol = 0.0;
      for (int i = 0; i < p.dim; ++i) {
        m->h_p0[p.offset + i] = p.b[i];  // lower
        m->h_p1[p.offset + i] = p.a[i];  // upper
        neg_log_vol -= std::log(p.a[i] - p.b[i]);
      }
      log_const += neg_log_vol;
    } else if (p.kind == PK_CAUCHY) {
      double logdet = 0.0;
      for (int i = 0; i < p.dim; ++i) {
        m->h_p0[p.offset + i] = p.a[i];
        m->h_p1[p.offset + i] = 1.0 / p.b[i];
        logdet += std::log(p.b[i]);
      }
      SpecialTermDev t{};
      t.kind = PK_CAUCHY;
      t.offset = p.offset;
      t.dim = p.dim;
      // CauchyDistribution.scala:51-53 with lgamma / log-det (identical wherever the reference is finite, Q4)
      t.log_mult = std::lgamma((1 + p.dim) / 2.0) - std::lgamma(0.5) - p.dim / 2.0 * std::log(M_PI) - logdet / 2.0;
      t.grad_sign = m->opt_cauchy_ref_sign ? 1.0 : -1.0;
      t.precision = nullptr;
      m->h_special.push_back(t);
      m->h_chol_off.push_back(-1);
    } else {  // PK_GAUSS_FULL
      std::vector<double> L, P;
      if (!cholesky_lower(p.b, p.dim, L)) return fail(THY_ERR_INVALID_ARGUMENT, "covariance not positive definite");
      spd_inverse(L, p.dim, P);
      double lognorm = p.dim / 2.0 * LOG_2PI;
      for (int i = 0; i < p.dim; ++i) {
        lognorm += std::log(L[(size_t)i * p.dim + i]);
        m->h_p0[p.offset + i] = p.a[i];
      }
      auto buf = std::make_unique<DevBuf<double>>();
      THY_TRY(buf->upload(P));
      SpecialTermDev t{};
      t.kind = PK_GAUSS_FULL;
      t.offset = p.offset;
      t.dim = p.dim;
      t.log_mult = -lognorm;
      t.grad_sign = 1.0;
      t.precision = buf->ptr;
      m->pr_precisions.push_back(std::move(buf));
      m->h_special.push_back(t);
      m->h_chol_off.push_back((int)chol_all.size());
      chol_all.insert(chol_all.end(), L.begin(), L.end());
    }
  }
  THY_TRY(m->pr_kind.upload(m->h_kind));
  THY_TRY(m->pr_p0.upload(m->h_p0));
  THY_TRY(m->pr_p1.upload(m->h_p1));
  THY_TRY(m->pr_special.upload(m->h_special));
  THY_TRY(m->pr_chol.upload(chol_all));
  THY_TRY(m->pr_chol_off.upload(m->h_chol_off));
  m->prior_view.d = d;
  m->prior_view.kind = m->pr_kind.ptr;
  m->prior_view.p0 = m->pr_p0.ptr;
  m->prior_view.p1 = m->pr_p1.ptr;
  m->prior_view.n_special = (int)m->h_special.size();
  m->prior_view.special = m->pr_special.ptr;
  m->prior_view.log_const = log_const;

  // ---- likelihoods ----
  for (auto& L : m->liks) {
    // blocks in ascending label order (LinearForwardModel.scala:87-101)
    std::vector<int> order(L.labels.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return L.labels[a] < L.labels[b]; });
    std::vector<int> cols;
    std::vector<const double*> blocks;
    std::vector<int> widths;
    for (int b : order) {
      auto it = m->layout.find(L.labels[b]);
      if (it == m->layout.end()) return fail(THY_ERR_UNKNOWN_LABEL, "unknown label '" + L.labels[b] + "'");
      for (int j = 0; j < L.dims[b]; ++j) cols.push_back(it->second.first + j);
      blocks.push_back(L.blocks[b].data());
      widths.push_back(L.dims[b]);
    }
    THY_TRY(build_lik_dev(m, L.distribution, L.link, L.jacobian, L.differential, L.n, cols, blocks, widths, L.meas.data(),
                          L.unc.data(), L.has_offset ? L.offset.data() : nullptr, nullptr));
    // host copies of the big blocks are no longer needed
    L.blocks.clear();
    L.blocks.shrink_to_fit();
  }
  THY_CUDA_CHECK(cudaDeviceSynchronize());
  if (m->stream == 0) {
    THY_CUDA_CHECK(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamDefault));
    m->stream = m->own_stream;
  }
  m->finalized = true;
  return THY_OK;
}

thy_status thy_model_dimension(thy_model_t m, int* out_dim) {
  thy::ErrScope _err_scope(m);
  THY_TRY(check_model(m, true));
  if (!out_dim) return fail(THY_ERR_INVALID_ARGUMENT, "null out pointer");
  *out_dim = m->d;
  return THY_OK;
}

thy_status thy_model_label_layout(thy_model_t m, const char* label, int* out_offset, int* out_dim) {
  thy::ErrScope _err_scope(m);
  THY_TRY(check_model(m, true));
  if (!label) return fail(THY_ERR_INVALID_ARGUMENT, "null label");
  auto it = m->layout.find(label);
  if (it == m->layout.end()) return fail(THY_ERR_UNKNOWN_LABEL, std::string("unknown label '") + label + "'");
  if (out_offset) *out_offset = it->second.first;
  if (out_dim) *out_dim = it->second.second;
  return THY_OK;
}

thy_status thy_model_set_stream(thy_model_t m, void* cuda_stream) {
  thy::ErrScope _err_scope(m);
  if (!m) return fail(THY_ERR_INVALID_ARGUMENT, "null model handle");
  m->stream = cuda_stream ? (cudaStream_t)cuda_stream : (m->own_stream ? m->own_stream : (cudaStream_t)0);
  return THY_OK;
}

thy_status thy_model_synchronize(thy_model_t m) {
  thy::ErrScope _err_scope(m);
  THY_TRY(check_model(m, true));
  THY_CUDA_CHECK(cudaStreamSynchronize(m->stream));
  return THY_OK;
}

const char* thy_model_last_kernel_path(thy_model_t m) { return m ? m->last_path : "none"; }

}  // extern "C"
