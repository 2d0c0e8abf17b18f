// Host-side model description and the POD device views handed to kernels.
#pragma once
#include "common.cuh"
#include "user_link.h"
#include <map>
#include <memory>
#include <mutex>

namespace thy {

// ---- per-coordinate prior kinds (one prior owns each flat coordinate) -----------------------------
enum PriorKind : int { PK_GAUSS_DIAG = 0, PK_UNIFORM = 1, PK_CAUCHY = 2, PK_GAUSS_FULL = 3 };

struct PriorHost {
  std::string label;
  int dim = 0;
  int kind = PK_GAUSS_DIAG;
  std::vector<double> a;  // mean | upper bounds
  std::vector<double> b;  // variances | lower bounds | dense covariance (dim*dim)
  int offset = -1;        // position in the flat vector, assigned at finalize
};

struct LikelihoodHost {
  int distribution = 0, link = 0, jacobian = 0, n = 0;
  double differential = 0.0;
  std::vector<std::string> labels;
  std::vector<int> dims;
  std::vector<std::vector<double>> blocks;
  bool has_offset = false;
  std::vector<double> offset, meas, unc;
};

// Cauchy and full-covariance Gaussian priors need a per-block reduction / mat-vec.
struct SpecialTermDev {
  int kind;          // PK_CAUCHY or PK_GAUSS_FULL
  int offset, dim;
  double log_mult;   // Cauchy: logMultiplier; full Gaussian: -(dim/2 log 2pi + sum log diag chol)
  double grad_sign;  // Cauchy: +1 reference-faithful (Q3), -1 correct; full Gaussian: unused
  const double* precision;  // full Gaussian: dim*dim row-major inverse covariance (device)
};

struct PriorDev {
  int d;
  const int* kind;     // [d]
  const double* p0;    // [d] mean | lower
  const double* p1;    // [d] inverse variance | upper
  int n_special;
  const SpecialTermDev* special;
  double log_const;    // sum over diagonal-Gaussian normalisers and uniform -log V
};

struct LikDev {
  int distribution, link, jacobian;
  int n, dl;              // observations; compact domain width (columns this forward model reads)
  int n_pad, lds;         // rows padded to a multiple of 32; row pitch (doubles), lds % 16 in {4, 12}
  const double* A;        // [n_pad][lds], zero padded
  const double2* yiv;     // [n_pad] (y, 1/var); Uniform: (upper, lower); padded rows (0, 0)
  const double2* yiv_lin; // [n_pad] (y - offset, 1/var) for the linear fast paths
  const double* off;      // [n_pad] vectorOffset (0 when absent)
  const int* col;         // [dl] flat index of compact column k
  double log_const;       // Gaussian: -(n/2 log 2pi + sum log sqrt(var)); Cauchy: logMultiplier; Uniform: -log V
  double differential;
  double cauchy_sign;     // +1 reference-faithful, -1 correct
};

struct LikDevOwner {
  LikDev view{};
  DevBuf<double> A, off;
  DevBuf<double> AF, AG;    // wide path: A re-tiled for the forward / adjoint DMMA GEMMs (built lazily)
  DevBuf<double2> yiv, yiv_lin;
  DevBuf<int> col;
  std::vector<int> col_host;
  bool contiguous = false;  // columns are a contiguous flat range starting at col[0]
};

struct Workspace {  // grow-only scratch owned by the model, reused across evaluations
  DevBuf<double> X, logp, grad;         // staging for the host-pointer API
  DevBuf<double> W;                     // generic path: weighted residuals [B][n_pad]
  DevBuf<double> partG, partL;          // stream-K partial tiles
  DevBuf<double> Z;                     // generic path scratch
  DevBuf<double> XT;                    // wide path: packed / tiled chain vectors
  DevBuf<unsigned int> ticket;          // streaming kernel: last-CTA ticket, zero between launches
  DevBuf<unsigned int> fused_tickets;   // fused kernel: per (tile, warp) segment tickets, zero between launches
  // Number of times any of the buffers above has moved.  A CUDA graph captured over an evaluation holds their
  // addresses: its owner compares this epoch before every replay and re-captures when it has changed.
  uint64_t epoch = 0;
  Workspace() {
    X.realloc_counter = logp.realloc_counter = grad.realloc_counter = W.realloc_counter = partG.realloc_counter =
        partL.realloc_counter = Z.realloc_counter = XT.realloc_counter = &epoch;
    ticket.realloc_counter = fused_tickets.realloc_counter = &epoch;
  }
  Workspace(const Workspace&) = delete;
  Workspace& operator=(const Workspace&) = delete;
};

}  // namespace thy

struct thy_model_s {
  std::vector<thy::PriorHost> priors;
  std::vector<thy::LikelihoodHost> liks;
  std::string last_error;   // message of the last failing call made on this handle (or a sampler built on it)
  std::mutex err_mutex;
  bool finalized = false;
  bool poisoned = false;   // a finalize failed half-way: the handle only accepts thy_model_destroy
  int d = 0;
  int opt_cauchy_ref_sign = 1;
  int opt_fd_incremental = 0;
  int opt_kernel_path = 0;
  int opt_fuse_prior = 1;
  cudaStream_t stream = 0;
  // Default stream of the handle, created at finalize: a BLOCKING stream, so it keeps the implicit ordering with the
  // legacy default stream that callers of the *_dev entry points rely on, yet (unlike stream 0) it can be captured
  // into CUDA graphs.  thy_model_set_stream replaces it.
  cudaStream_t own_stream = nullptr;
  int device = -1;
  std::map<std::string, std::pair<int, int>> layout;  // label -> (offset, dim)
  std::vector<std::unique_ptr<thy::UserLink>> user_links;  // thy_model_register_link: link id = THY_LINK_USER_BASE + index
  // device state
  thy::PriorDev prior_view{};
  thy::DevBuf<int> pr_kind;
  thy::DevBuf<double> pr_p0, pr_p1;
  thy::DevBuf<thy::SpecialTermDev> pr_special;
  std::vector<std::unique_ptr<thy::DevBuf<double>>> pr_precisions;
  std::vector<std::unique_ptr<thy::LikDevOwner>> lik_dev;
  thy::Workspace ws;
  const char* last_path = "none";
  // host-pointer API: copy streams + events so the H2D / D2H copies of one chunk overlap the kernels of the next
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t ev_h2d[8] = {}, ev_done[8] = {}, ev_start = nullptr;
  // asynchronous host-buffer calls (thy_logpdf_grad_batch_async): two staging sets so that call k+1's H2D copy overlaps
  // call k's kernels and call k-1's D2H copy; per-set events order the reuse without blocking the host
  struct AsyncSet {
    thy::DevBuf<double> X, logp, grad;
    cudaEvent_t h2d_done = nullptr, compute_done = nullptr, d2h_done = nullptr;
    int64_t ticket = 0;   // last call that used this set
  } async_set[2];
  int64_t async_next_ticket = 1;
  // optional CUDA-event timing of the dominant kernel (bench.py's roofline leg)
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events;
  // host copies kept for prior sampling
  std::vector<int> h_kind;
  std::vector<double> h_p0, h_p1;
  std::vector<thy::SpecialTermDev> h_special;
  std::vector<std::vector<double>> h_chol;  // per special term: lower Cholesky factor (dim*dim), row-major
  thy::DevBuf<double> pr_chol;              // concatenated Cholesky factors of full-covariance priors
  std::vector<int> h_chol_off;
  thy::DevBuf<int> pr_chol_off;
};

namespace thy {
// Records a CUDA-event pair around the dominant kernel on the launching stream when timing is enabled.
struct KernelTimer {
  thy_model_s* m;
  cudaStream_t s;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  KernelTimer(thy_model_s* m_, cudaStream_t s_) : m(m_), s(s_) {
    if (m->timing && cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) cudaEventRecord(e0, s);
  }
  void stop() {
    if (e0 && e1) {
      cudaEventRecord(e1, s);
      m->timing_events.emplace_back(e0, e1);
      e0 = e1 = nullptr;
    }
  }
};
// eval.cu
thy_status eval_batch_dev(thy_model_s* m, int64_t B, const double* X, double* logp, double* grad, bool cauchy_correct_sign = false,
                          bool likelihoods_only = false);
struct HmcFuse;
bool hmc_step_fusable(thy_model_s* m, int64_t B);
thy_status eval_batch_hmc(thy_model_s* m, int64_t B, const double* X, double* logp, double* grad_scratch, const HmcFuse& hf);
thy_status launch_likelihood(thy_model_s* m, LikDevOwner& owner, int64_t B, const double* X, double* logp, double* grad,
                             double csign, cudaStream_t s, bool fuse_prior = false);
// model.cu
thy_status build_lik_dev(thy_model_s* m, int distribution, int link, int jacobian, double differential, int n,
                         const std::vector<int>& cols, const std::vector<const double*>& blocks,
                         const std::vector<int>& widths, const double* meas, const double* unc, const double* offset,
                         const double* log_const_override);
bool cholesky_lower(const std::vector<double>& S, int n, std::vector<double>& L);
void spd_inverse(const std::vector<double>& L, int n, std::vector<double>& P);
// sampling.cu
thy_status sample_priors_dev(thy_model_s* m, int64_t B, uint64_t seed, double* X);
}
