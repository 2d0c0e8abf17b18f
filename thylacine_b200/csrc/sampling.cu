// Prior draws (Posterior.samplePriors, model/components/posterior/Posterior.scala:64-65):
//   Gaussian: mu + chol(Sigma) z          (Breeze MultivariateGaussian.sample; GaussianPrior.scala:47-48)
//   Uniform : lo + u (hi - lo), u in [0,1) (distributions/UniformDistribution.scala:88-93)
//   Cauchy  : MVN(mu, Sigma / chi2_1)     (distributions/CauchyDistribution.scala:84-87)
#include "kernels.h"
#include "philox.cuh"

namespace thy {

__global__ void __launch_bounds__(256) k_sample_priors(PriorDev pr, const double* __restrict__ chol,
                                                       const int* __restrict__ chol_off, long long B, uint64_t seed,
                                                       double* __restrict__ X) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * pr.d) return;
  const long long b = idx / pr.d;
  const int j = (int)(idx % pr.d);
  const int kind = pr.kind[j];
  double v;
  if (kind == PK_GAUSS_DIAG) {
    v = pr.p0[j] + rng_normal(seed, b, j, RNG_PRIOR, 0) / sqrt(pr.p1[j]);
  } else if (kind == PK_UNIFORM) {
    const RngDraw d = rng_uniform2(seed, b, j, RNG_PRIOR, 0);
    v = pr.p0[j] + d.u0 * (pr.p1[j] - pr.p0[j]);
  } else {
    // locate the owning special term
    int t = 0;
    for (; t < pr.n_special; ++t)
      if (j >= pr.special[t].offset && j < pr.special[t].offset + pr.special[t].dim) break;
    const SpecialTermDev st = pr.special[t];
    if (kind == PK_CAUCHY) {
      const double z0 = rng_normal(seed, b, t, RNG_PRIOR_CHI, 0);  // chi2_1 = z0^2
      v = pr.p0[j] + rng_normal(seed, b, j, RNG_PRIOR, 0) / sqrt(pr.p1[j]) / fabs(z0);
    } else {
      const double* L = chol + chol_off[t];
      const int i = j - st.offset;
      double s = 0.0;
      for (int k = 0; k <= i; ++k) s += L[(size_t)i * st.dim + k] * rng_normal(seed, b, st.offset + k, RNG_PRIOR, 0);
      v = pr.p0[j] + s;
    }
  }
  X[idx] = v;
}

thy_status sample_priors_dev(thy_model_s* m, int64_t B, uint64_t seed, double* X) {
  if (!m || !m->finalized) return fail(THY_ERR_STATE, "model is not finalized");
  if (B <= 0) return THY_OK;
  if (!X) return fail(THY_ERR_INVALID_ARGUMENT, "null device pointer");
  const long long total = (long long)B * m->d;
  k_sample_priors<<<(unsigned)ceil_div(total, 256), 256, 0, m->stream>>>(m->prior_view, m->pr_chol.ptr, m->pr_chol_off.ptr,
                                                                        B, seed, X);
  count_launch();
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

}  // namespace thy

using namespace thy;

extern "C" {

thy_status thy_sample_priors_dev(thy_model_t m, int64_t B, uint64_t rng_seed, double* X_dev) {
  thy::ErrScope _err_scope(m);
  THY_TRY(require_device());
  return sample_priors_dev(m, B, rng_seed, X_dev);
}

thy_status thy_sample_priors(thy_model_t m, int64_t B, uint64_t rng_seed, double* out_X) {
  thy::ErrScope _err_scope(m);
  THY_TRY(require_device());
  if (!m || !m->finalized) return fail(THY_ERR_STATE, "model is not finalized");
  if (B <= 0) return THY_OK;
  if (!out_X) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  const size_t nx = (size_t)B * m->d;
  THY_TRY(m->ws.X.reserve(nx));
  THY_TRY(sample_priors_dev(m, B, rng_seed, m->ws.X.ptr));
  THY_CUDA_CHECK(cudaMemcpyAsync(out_X, m->ws.X.ptr, nx * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  THY_CUDA_CHECK(cudaStreamSynchronize(m->stream));
  return THY_OK;
}

}  // extern "C"
