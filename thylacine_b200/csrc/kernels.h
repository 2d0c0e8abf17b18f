// Kernel launchers shared between translation units.
#pragma once
#include "model.h"

namespace thy {

// terms.cu -- priors (initialises logp / grad), warp-shuffle reductions
thy_status launch_prior(const PriorDev& pr, int64_t B, const double* X, double* logp, double* grad, double cauchy_sign,
                        cudaStream_t s);

// kernels_generic.cu -- any link / distribution / jacobian mode, any shape (tiled DFMA GEMMs, W materialised)
thy_status launch_generic_likelihood(thy_model_s* m, const LikDevOwner& lik, int64_t B, const double* X, double* logp,
                                     double* grad, double cauchy_sign, cudaStream_t s);

// One leapfrog step of B lock-step HMC chains folded into the output step of the gradient kernel
// (HmcmcEngine.scala:55-83): the kernel that owns a chain's complete gradient row also applies the momentum kicks and the
// drift (MID), or the last half kick, the Hamiltonian, the Metropolis test and the state update (LAST).  Positions are
// ping-ponged: the kernel reads X and writes x_next (stream-K tiles are read by several CTAs, so X is never updated in place).
enum HmcFuseMode : int { HMC_FUSE_NONE = 0, HMC_FUSE_MID = 1, HMC_FUSE_LAST = 2 };
struct HmcFuse {
  int mode = HMC_FUSE_NONE;
  double eps = 0.0;
  double* p = nullptr;            // [B][d] momenta (read-modify-write)
  double* x_next = nullptr;       // MID: [B][d] positions after the drift
  // LAST:
  double* x = nullptr;            // [B][d] chain state, overwritten on acceptance
  double* g = nullptr;            // [B][d] gradient at the chain state
  double* lp = nullptr;           // [B]
  const double* H0 = nullptr;     // [B]
  double* dH = nullptr;           // [B]
  long long* attempts = nullptr;
  long long* acceptances = nullptr;
  long long* block_acc = nullptr;
  uint64_t seed = 0;
  const unsigned long long* traj = nullptr;   // device trajectory counter (keys the acceptance draw)
};

// kernels_fused_dmma.cu -- linear forward model + Gaussian/Cauchy observations, fused forward/residual/adjoint
int fused_nt_for(int dl);  // 0 when the width is not covered by the fused kernel
bool fused_dmma_supported(const LikDev& lik, bool want_grad);
double fused_cost_model(const LikDev& lik, int64_t B);    // seconds, for auto-selection
// fuse_prior: the kernel also evaluates the per-coordinate priors and WRITES logp / grad (single term over the whole
// flat vector, no Cauchy / full-covariance prior blocks: fused_prior_supported)
bool fused_prior_supported(const thy_model_s* m, const LikDevOwner& lik);
// The launcher picks the resident variant (A held in shared memory for the CTA's lifetime) when resident_dmma_supported.
bool resident_dmma_supported(const LikDev& lik, int64_t B);
thy_status launch_fused_dmma(thy_model_s* m, const LikDevOwner& lik, int64_t B, const double* X, double* logp, double* grad,
                             double cauchy_sign, cudaStream_t s, bool fuse_prior = false, const HmcFuse* hmc = nullptr);

// kernels_stream.cu -- small-batch memory-bound pass (A streamed once per batch)
bool stream_supported(const LikDev& lik, int64_t B);
double stream_cost_model(const LikDev& lik, int64_t B);   // seconds, for auto-selection
thy_status launch_stream(thy_model_s* m, const LikDevOwner& lik, int64_t B, const double* X, double* logp, double* grad,
                         double cauchy_sign, cudaStream_t s);

// kernels_wide.cu -- any width / any link on the fp64 tensor cores (two tiled DMMA GEMMs, W materialised in tiled form)
bool wide_supported(const LikDev& lik);
thy_status launch_wide_likelihood(thy_model_s* m, LikDevOwner& lik, int64_t B, const double* X, double* logp, double* grad,
                                  double cauchy_sign, cudaStream_t s);

// slq_walk.cu -- SLQ replacement step (spawn + M constrained walk steps + commit) as one kernel for small linear models
bool slq_walk_supported(const thy_model_s* m);
// c_max sizes the grid; the walker count and the round number are read from device words (graph replay)
thy_status launch_slq_walk(thy_model_s* m, int c_max, const int* c_dev, long long Pl, int M, uint64_t seed,
                           const long long* round_dev, int rank, const int* sidx, const unsigned char* flags, double* x,
                           double* ll, double* lpr, const double* sigma, const double* scalars, int* accepted, cudaStream_t s,
                           const float* zpre = nullptr, const double* lacc = nullptr, int pre_groups = 0,
                           int expected = 0);   // walkers expected on this rank (sizes the CTAs; 0 = c_max)
// the round's proposal normals / acceptance thresholds for walker groups [0, groups), drawn ahead of the walk
size_t slq_walk_rng_floats_per_group(const thy_model_s* m, int M);
thy_status launch_slq_walk_rng(thy_model_s* m, int groups, int M, uint64_t seed, const long long* round_dev, int rank,
                               float* zpre, double* lacc, cudaStream_t s);

// link functions (device), shared by the generic and FD kernels
__device__ __forceinline__ double link_eval(int link, double z) {
  switch (link) {
    case THY_LINK_TANH: return tanh(z);
    case THY_LINK_EXP: return exp(z);
    case THY_LINK_SQUARE: return z * z;
    case THY_LINK_SOFTPLUS: return z > 0 ? z + log1p(exp(-z)) : log1p(exp(z));
    case THY_LINK_SIGMOID: return 1.0 / (1.0 + exp(-z));
    default: return z;
  }
}
__device__ __forceinline__ double link_deriv(int link, double z) {
  switch (link) {
    case THY_LINK_TANH: { double t = tanh(z); return 1.0 - t * t; }
    case THY_LINK_EXP: return exp(z);
    case THY_LINK_SQUARE: return 2.0 * z;
    case THY_LINK_SOFTPLUS: return 1.0 / (1.0 + exp(-z));
    case THY_LINK_SIGMOID: { double s = 1.0 / (1.0 + exp(-z)); return s * (1.0 - s); }
    default: return 1.0;
  }
}

// derivative of the link given both the pre-activation z and the value f = link(z) already in hand
// (saves the second transcendental: tanh' = 1 - f^2, exp' = f, sigmoid' = f (1 - f))
__device__ __forceinline__ double link_deriv_from(int link, double z, double f) {
  switch (link) {
    case THY_LINK_TANH: return 1.0 - f * f;
    case THY_LINK_EXP: return f;
    case THY_LINK_SQUARE: return 2.0 * z;
    case THY_LINK_SOFTPLUS: return 1.0 / (1.0 + exp(-z));
    case THY_LINK_SIGMOID: return f * (1.0 - f);
    default: return 1.0;
  }
}

}  // namespace thy
