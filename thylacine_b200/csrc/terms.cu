// Prior terms of the posterior sum: one warp per point, warp-shuffle reductions.
// Reference arithmetic (paths relative to /root/reference/src/main/scala/thylacine/):
//   model/components/prior/Prior.scala:52-66               each prior reads its own labelled block
//   model/distributions/GaussianDistribution.scala:52-68   -1/2 r^T S^-1 r - lognorm ; S^-1 (mu - x)
//   model/distributions/UniformDistribution.scala:50-81    half-open box, -log V or -inf, zero gradient
//   model/distributions/CauchyDistribution.scala:51-79     logMult - (1+n)/2 log(1+q) ; +(1+n)/(1+q) S^-1 (x-mu) (Q3)
//   model/components/posterior/Posterior.scala:50-73       priors summed first, likelihoods added afterwards
// This kernel INITIALISES logp[b] and grad[b, :]; likelihood kernels add to them.
#include "kernels.h"

namespace thy {

__global__ void __launch_bounds__(256) k_prior(PriorDev pr, long long B, const double* __restrict__ X,
                                               double* __restrict__ logp, double* __restrict__ grad, double cauchy_sign) {
  const int lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int d = pr.d;
  const double* x = X + b * d;
  double* g = grad ? grad + b * d : nullptr;
  double acc = 0.0;
  bool inside = true;
  for (int j = lane; j < d; j += 32) {
    const int kind = pr.kind[j];
    const double xv = x[j];
    double gj = 0.0;
    if (kind == PK_GAUSS_DIAG) {
      const double dl = pr.p0[j] - xv;  // mean - x
      const double w = dl * pr.p1[j];
      acc += dl * w;
      gj = w;
    } else if (kind == PK_UNIFORM) {
      inside = inside && (xv >= pr.p0[j]) && (xv < pr.p1[j]);
    }
    if (g) g[j] = gj;  // Cauchy / full-covariance coordinates are overwritten below
  }
  acc = warp_sum(acc);
  inside = __all_sync(0xffffffffu, inside);
  double lp = -0.5 * acc + pr.log_const;
  for (int t = 0; t < pr.n_special; ++t) {
    const SpecialTermDev st = pr.special[t];
    if (st.kind == PK_CAUCHY) {
      double q = 0.0;
      for (int j = lane; j < st.dim; j += 32) {
        const double dl = x[st.offset + j] - pr.p0[st.offset + j];
        q += dl * dl * pr.p1[st.offset + j];
      }
      q = warp_sum(q);
      lp += st.log_mult - (1.0 + st.dim) / 2.0 * log(1.0 + q);
      if (g) {
        const double mult = cauchy_sign * (1.0 + st.dim) / (1.0 + q);
        for (int j = lane; j < st.dim; j += 32) {
          const double dl = x[st.offset + j] - pr.p0[st.offset + j];
          g[st.offset + j] = mult * (dl * pr.p1[st.offset + j]);
        }
      }
    } else {  // PK_GAUSS_FULL: v = P (mu - x)
      double quad = 0.0;
      for (int i = lane; i < st.dim; i += 32) {
        const double* prow = st.precision + (size_t)i * st.dim;
        double v = 0.0;
        for (int j = 0; j < st.dim; ++j) v += prow[j] * (pr.p0[st.offset + j] - x[st.offset + j]);
        quad += (pr.p0[st.offset + i] - x[st.offset + i]) * v;
        if (g) g[st.offset + i] = v;
      }
      quad = warp_sum(quad);
      lp += -0.5 * quad + st.log_mult;
    }
  }
  if (!inside) lp = -INFINITY;
  if (lane == 0) logp[b] = lp;
}

thy_status launch_prior(const PriorDev& pr, int64_t B, const double* X, double* logp, double* grad, double cauchy_sign,
                        cudaStream_t s) {
  if (B <= 0) return THY_OK;
  const int warps = 8;
  const long long blocks = (B + warps - 1) / warps;
  k_prior<<<(unsigned)blocks, warps * 32, 0, s>>>(pr, B, X, logp, grad, cauchy_sign);
  count_launch();
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

}  // namespace thy
