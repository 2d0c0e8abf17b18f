// Batched evaluation of the posterior sum: Posterior.logPdfAt / logPdfGradientAt for B points at once.
//   /root/reference/src/main/scala/thylacine/model/components/posterior/Posterior.scala:50-73
// Order: priors initialise logp / grad, then each likelihood term adds its contribution (the reference sums
// priors first, likelihoods after, left to right -- SURVEY Q7).
#include "kernels.h"
#include <algorithm>

namespace thy {

// Kernel path of one likelihood term for a batch of B points (0 = not resolvable: status set).
static int select_path(thy_model_s* m, const LikDev& lk, int64_t B, bool want_grad, thy_status* st) {
  *st = THY_OK;
  const bool linear = (lk.link == THY_LINK_IDENTITY && lk.jacobian == THY_JAC_CONSTANT &&
                       lk.distribution != THY_DIST_UNIFORM);
  int path = m->opt_kernel_path;
  if (lk.link >= THY_LINK_USER_BASE) return 4;   // NVRTC-compiled link: only the wide path has the hook
  if (path == 5) path = 2;   // 5 = fused DMMA, streaming variant only (launch_fused_dmma looks at the option itself)
  if (path == 2 && !(linear && fused_dmma_supported(lk, want_grad))) {
    *st = fail(THY_ERR_UNSUPPORTED, "fused DMMA kernel does not support this likelihood shape");
    return 0;
  }
  if (path == 3 && !(linear && stream_supported(lk, B))) {
    *st = fail(THY_ERR_UNSUPPORTED, "streaming kernel does not support this likelihood shape / batch");
    return 0;
  }
  if (path == 4 && !wide_supported(lk)) {
    *st = fail(THY_ERR_UNSUPPORTED, "wide DMMA path does not support this likelihood (uniform observations / FD Jacobian)");
    return 0;
  }
  if (path == 0) {
    if (linear && stream_supported(lk, B) &&
        (!fused_dmma_supported(lk, want_grad) || stream_cost_model(lk, B) <= fused_cost_model(lk, B)))
      path = 3;
    else if (linear && fused_dmma_supported(lk, want_grad))
      path = 2;
    else if (wide_supported(lk) && B >= 32 && (long long)lk.n * lk.dl >= 4096)
      path = 4;
    else
      path = 1;
  }
  return path;
}

// One likelihood term added into logp / grad (fuse_prior: the fused kernel also evaluates the priors and writes).
thy_status launch_likelihood(thy_model_s* m, LikDevOwner& owner, int64_t B, const double* X, double* logp, double* grad,
                             double csign, cudaStream_t s, bool fuse_prior) {
  thy_status st = THY_OK;
  const int path = select_path(m, owner.view, B, grad != nullptr, &st);
  THY_TRY(st);
  if (fuse_prior && path != 2) return fail(THY_ERR_STATE, "fused priors need the fused DMMA path");
  if (path == 2) {
    m->last_path = "fused_dmma";
    THY_TRY(launch_fused_dmma(m, owner, B, X, logp, grad, csign, s, fuse_prior));
  } else if (path == 3) {
    m->last_path = "stream";
    THY_TRY(launch_stream(m, owner, B, X, logp, grad, csign, s));
  } else if (path == 4) {
    m->last_path = "wide_dmma";
    THY_TRY(launch_wide_likelihood(m, owner, B, X, logp, grad, csign, s));
  } else {
    m->last_path = "generic";
    THY_TRY(launch_generic_likelihood(m, owner, B, X, logp, grad, csign, s));
  }
  return THY_OK;
}

// True when a whole evaluation of B points should be ONE kernel: the fused DMMA kernel with the priors folded into its
// output step.  Measured on B200: the folded output step costs the fused kernel ~3 % (config 2: 543 -> 561 us) and saves
// the prior kernel (a launch plus one pass over X and the outputs), so it pays for small terms -- the collapsed
// posterior gains 18-27 % -- and not for config 2.  opt_fuse_prior: 0 never, 1 by this cost model, 2 always.
// (Writing the outputs straight into pinned host memory from the same kernel was tried and dropped: its scattered
// 8-byte reads and writes across PCIe halved the end-to-end rate.)
static bool single_kernel_eval(thy_model_s* m, int64_t B, bool want_grad) {
  if (m->lik_dev.size() != 1 || !m->opt_fuse_prior) return false;
  LikDevOwner& own = *m->lik_dev[0];
  if (!fused_prior_supported(m, own)) return false;
  thy_status st = THY_OK;
  if (select_path(m, own.view, B, want_grad, &st) != 2) return false;
  if (m->opt_fuse_prior >= 2) return true;
  const double t_prior = 5e-6 + (double)B * m->d * 24.0 / 6e12;
  return 0.03 * fused_cost_model(own.view, B) < t_prior;
}

// A leapfrog step can be folded into the gradient kernel's output step when the whole posterior is ONE fused-DMMA launch:
// a single linear term over the whole flat vector with per-coordinate priors (fused_prior_supported).
bool hmc_step_fusable(thy_model_s* m, int64_t B) {
  if (!m || !m->finalized || m->lik_dev.size() != 1) return false;
  LikDevOwner& own = *m->lik_dev[0];
  if (!fused_prior_supported(m, own)) return false;
  thy_status st = THY_OK;
  const int path = select_path(m, own.view, B, true, &st);
  return st == THY_OK && path == 2;
}

// Gradient at X with the HMC step `hf` applied in the same launch (samplers always use the correct Cauchy sign).
thy_status eval_batch_hmc(thy_model_s* m, int64_t B, const double* X, double* logp, double* grad_scratch, const HmcFuse& hf) {
  if (!hmc_step_fusable(m, B)) return fail(THY_ERR_STATE, "the fused HMC step does not cover this model");
  m->last_path = "fused_dmma";
  return launch_fused_dmma(m, *m->lik_dev[0], B, X, logp, grad_scratch, -1.0, m->stream, true, &hf);
}

thy_status eval_batch_dev(thy_model_s* m, int64_t B, const double* X, double* logp, double* grad, bool cauchy_correct_sign,
                          bool likelihoods_only) {
  if (!m || !m->finalized) return fail(THY_ERR_STATE, "model is not finalized");
  if (B < 0) return fail(THY_ERR_INVALID_ARGUMENT, "negative batch size");
  if (B == 0) return THY_OK;
  if (!X || !logp) return fail(THY_ERR_INVALID_ARGUMENT, "null device pointer");
  const double csign = (cauchy_correct_sign || !m->opt_cauchy_ref_sign) ? -1.0 : 1.0;
  cudaStream_t s = m->stream;
  if (!likelihoods_only && single_kernel_eval(m, B, grad != nullptr))
    return launch_likelihood(m, *m->lik_dev[0], B, X, logp, grad, csign, s, true);
  // likelihoods_only: logp (and grad) are already initialised by the caller; only the likelihood terms are added
  if (!likelihoods_only) THY_TRY(launch_prior(m->prior_view, B, X, logp, grad, csign, s));
  for (auto& own : m->lik_dev) THY_TRY(launch_likelihood(m, *own, B, X, logp, grad, csign, s, false));
  if (m->lik_dev.empty()) m->last_path = "priors_only";
  return THY_OK;
}

static thy_status ensure_copy_streams(thy_model_s* m) {
  if (m->h2d_stream) return THY_OK;
  THY_CUDA_CHECK(cudaStreamCreateWithFlags(&m->h2d_stream, cudaStreamNonBlocking));
  THY_CUDA_CHECK(cudaStreamCreateWithFlags(&m->d2h_stream, cudaStreamNonBlocking));
  THY_CUDA_CHECK(cudaEventCreateWithFlags(&m->ev_start, cudaEventDisableTiming));
  for (int i = 0; i < 8; ++i) {
    THY_CUDA_CHECK(cudaEventCreateWithFlags(&m->ev_h2d[i], cudaEventDisableTiming));
    THY_CUDA_CHECK(cudaEventCreateWithFlags(&m->ev_done[i], cudaEventDisableTiming));
  }
  return THY_OK;
}

// Host-buffer entry point.  Large batches are cut into chunks: the H2D copy of chunk k+1 and the D2H copy of chunk k-1
// run on their own streams while the kernels of chunk k execute (pinned host buffers make the copies truly asynchronous;
// pageable buffers still work, just without overlap).  What cannot be hidden is the first chunk's H2D copy and the last
// chunk's D2H copy, so the plan is [small, large, small]: an eighth of the batch at either end, the rest in the middle
// (one large launch keeps the stream-K grid efficient).
static thy_status eval_host(thy_model_t m, int64_t B, const double* X, double* out_logp, double* out_grad) {
  if (!m || !m->finalized) return fail(THY_ERR_STATE, "model is not finalized");
  THY_TRY(require_device());
  if (B < 0) return fail(THY_ERR_INVALID_ARGUMENT, "negative batch size");
  if (B == 0) return THY_OK;
  if (!X || !out_logp) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  const size_t nx = (size_t)B * m->d;
  cudaStream_t s = m->stream;
  THY_TRY(m->ws.X.reserve(nx));
  THY_TRY(m->ws.logp.reserve((size_t)B));
  if (out_grad) THY_TRY(m->ws.grad.reserve(nx));
  static const int chunk_mode = [] {   // THY_HOST_CHUNKS: 1 = no overlap, 2..8 = that many equal chunks, unset = edge plan
    const char* e = std::getenv("THY_HOST_CHUNKS");
    const int v = e ? std::atoi(e) : 0;
    return v < 0 ? 0 : (v > 8 ? 8 : v);
  }();
  // chunk boundaries in whole 64-chain tiles
  int64_t bounds[9];
  int nch = 1;
  bounds[0] = 0;
  if (chunk_mode == 0) {
    if (B >= 2048) {
      const int64_t edge = std::max<int64_t>(256, (B / 8 + 63) / 64 * 64);
      bounds[1] = edge;
      bounds[2] = B - edge;
      nch = 3;
    }
  } else if (chunk_mode >= 2 && B / 1024 >= 2) {
    nch = (int)std::min<int64_t>(chunk_mode, B / 1024);
    const int64_t per = ((B + nch - 1) / nch + 63) / 64 * 64;
    for (int k = 1; k < nch; ++k) bounds[k] = std::min<int64_t>(B, per * k);
  }
  bounds[nch] = B;
  if (nch < 2) {
    THY_CUDA_CHECK(cudaMemcpyAsync(m->ws.X.ptr, X, nx * sizeof(double), cudaMemcpyHostToDevice, s));
    THY_TRY(eval_batch_dev(m, B, m->ws.X.ptr, m->ws.logp.ptr, out_grad ? m->ws.grad.ptr : nullptr, false));
    THY_CUDA_CHECK(cudaMemcpyAsync(out_logp, m->ws.logp.ptr, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (out_grad) THY_CUDA_CHECK(cudaMemcpyAsync(out_grad, m->ws.grad.ptr, nx * sizeof(double), cudaMemcpyDeviceToHost, s));
    THY_CUDA_CHECK(cudaStreamSynchronize(s));
    return THY_OK;
  }
  THY_TRY(ensure_copy_streams(m));
  // the copy streams must not start before earlier work on the model's stream has finished with the staging buffers
  THY_CUDA_CHECK(cudaEventRecord(m->ev_start, s));
  THY_CUDA_CHECK(cudaStreamWaitEvent(m->h2d_stream, m->ev_start, 0));
  // all input copies are queued up front (they serialise on the H2D stream), so the copy engine never waits for the host
  for (int k = 0; k < nch; ++k) {
    const int64_t b0 = bounds[k], nb = bounds[k + 1] - bounds[k];
    const size_t off = (size_t)b0 * m->d, cnt = (size_t)nb * m->d;
    THY_CUDA_CHECK(cudaMemcpyAsync(m->ws.X.ptr + off, X + off, cnt * sizeof(double), cudaMemcpyHostToDevice, m->h2d_stream));
    THY_CUDA_CHECK(cudaEventRecord(m->ev_h2d[k], m->h2d_stream));
  }
  for (int k = 0; k < nch; ++k) {
    const int64_t b0 = bounds[k], nb = bounds[k + 1] - bounds[k];
    const size_t off = (size_t)b0 * m->d, cnt = (size_t)nb * m->d;
    THY_CUDA_CHECK(cudaStreamWaitEvent(s, m->ev_h2d[k], 0));
    THY_TRY(eval_batch_dev(m, nb, m->ws.X.ptr + off, m->ws.logp.ptr + b0, out_grad ? m->ws.grad.ptr + off : nullptr, false));
    THY_CUDA_CHECK(cudaEventRecord(m->ev_done[k], s));
    THY_CUDA_CHECK(cudaStreamWaitEvent(m->d2h_stream, m->ev_done[k], 0));
    THY_CUDA_CHECK(cudaMemcpyAsync(out_logp + b0, m->ws.logp.ptr + b0, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost,
                                   m->d2h_stream));
    if (out_grad)
      THY_CUDA_CHECK(cudaMemcpyAsync(out_grad + off, m->ws.grad.ptr + off, cnt * sizeof(double), cudaMemcpyDeviceToHost,
                                     m->d2h_stream));
  }
  THY_CUDA_CHECK(cudaStreamSynchronize(m->d2h_stream));
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  return THY_OK;
}

// Pipelined host-buffer evaluation.  Call i uses staging set i % 2:
//   H2D stream:     wait until call i-2's kernels have finished with the set's X     -> copy X in
//   model stream:   wait for that copy and for call i-2's D2H copy of the set's outputs -> kernels
//   D2H stream:     wait for the kernels -> copy logp / grad out
// Nothing blocks the host; thy_model_wait(ticket) does.  With pinned buffers call i+1's input copy and call i-1's
// output copy run under call i's kernels (PCIe is full duplex), so back-to-back calls run at the device rate.
static thy_status eval_host_async(thy_model_t m, int64_t B, const double* X, double* out_logp, double* out_grad,
                                  int64_t* out_ticket) {
  if (!m || !m->finalized) return fail(THY_ERR_STATE, "model is not finalized");
  THY_TRY(require_device());
  if (B <= 0) return fail(THY_ERR_INVALID_ARGUMENT, "batch size must be positive");
  if (!X || !out_logp) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  THY_TRY(ensure_copy_streams(m));
  const int64_t ticket = m->async_next_ticket++;
  auto& set = m->async_set[ticket & 1];
  const size_t nx = (size_t)B * m->d;
  cudaStream_t s = m->stream;
  if (!set.h2d_done) {
    THY_CUDA_CHECK(cudaEventCreateWithFlags(&set.h2d_done, cudaEventDisableTiming));
    THY_CUDA_CHECK(cudaEventCreateWithFlags(&set.compute_done, cudaEventDisableTiming));
    THY_CUDA_CHECK(cudaEventCreateWithFlags(&set.d2h_done, cudaEventDisableTiming));
  }
  if (set.X.count < nx || set.logp.count < (size_t)B || (out_grad && set.grad.count < nx)) {
    // growing a staging buffer frees the old one: the set's previous call must have drained
    if (set.ticket) THY_CUDA_CHECK(cudaEventSynchronize(set.d2h_done));
    THY_TRY(set.X.reserve(nx));
    THY_TRY(set.logp.reserve((size_t)B));
    if (out_grad) THY_TRY(set.grad.reserve(nx));
  }
  if (set.ticket) THY_CUDA_CHECK(cudaStreamWaitEvent(m->h2d_stream, set.compute_done, 0));
  else {   // first use: order behind whatever the model stream is doing
    THY_CUDA_CHECK(cudaEventRecord(m->ev_start, s));
    THY_CUDA_CHECK(cudaStreamWaitEvent(m->h2d_stream, m->ev_start, 0));
  }
  THY_CUDA_CHECK(cudaMemcpyAsync(set.X.ptr, X, nx * sizeof(double), cudaMemcpyHostToDevice, m->h2d_stream));
  THY_CUDA_CHECK(cudaEventRecord(set.h2d_done, m->h2d_stream));
  THY_CUDA_CHECK(cudaStreamWaitEvent(s, set.h2d_done, 0));
  if (set.ticket) THY_CUDA_CHECK(cudaStreamWaitEvent(s, set.d2h_done, 0));
  THY_TRY(eval_batch_dev(m, B, set.X.ptr, set.logp.ptr, out_grad ? set.grad.ptr : nullptr, false));
  THY_CUDA_CHECK(cudaEventRecord(set.compute_done, s));
  THY_CUDA_CHECK(cudaStreamWaitEvent(m->d2h_stream, set.compute_done, 0));
  THY_CUDA_CHECK(cudaMemcpyAsync(out_logp, set.logp.ptr, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, m->d2h_stream));
  if (out_grad)
    THY_CUDA_CHECK(cudaMemcpyAsync(out_grad, set.grad.ptr, nx * sizeof(double), cudaMemcpyDeviceToHost, m->d2h_stream));
  THY_CUDA_CHECK(cudaEventRecord(set.d2h_done, m->d2h_stream));
  set.ticket = ticket;
  if (out_ticket) *out_ticket = ticket;
  return THY_OK;
}

}  // namespace thy

using namespace thy;

extern "C" {

thy_status thy_logpdf_grad_batch_async(thy_model_t m, int64_t B, const double* X, double* out_logp, double* out_grad,
                                       int64_t* out_ticket) {
  thy::ErrScope _err_scope(m);
  return eval_host_async(m, B, X, out_logp, out_grad, out_ticket);
}

thy_status thy_model_wait(thy_model_t m, int64_t ticket) {
  thy::ErrScope _err_scope(m);
  if (!m) return fail(THY_ERR_INVALID_ARGUMENT, "null model handle");
  for (auto& set : m->async_set) {
    if (!set.ticket || !set.d2h_done) continue;
    if (ticket <= 0 || set.ticket <= ticket) THY_CUDA_CHECK(cudaEventSynchronize(set.d2h_done));
  }
  return THY_OK;
}

thy_status thy_logpdf_batch(thy_model_t m, int64_t B, const double* X, double* out_logp) {
  thy::ErrScope _err_scope(m);
  return eval_host(m, B, X, out_logp, nullptr);
}

thy_status thy_logpdf_grad_batch(thy_model_t m, int64_t B, const double* X, double* out_logp, double* out_grad) {
  thy::ErrScope _err_scope(m);
  if (!out_grad) return fail(THY_ERR_INVALID_ARGUMENT, "null gradient pointer");
  return eval_host(m, B, X, out_logp, out_grad);
}

thy_status thy_logpdf_grad_batch_dev(thy_model_t m, int64_t B, const double* X_dev, double* logp_dev, double* grad_dev) {
  thy::ErrScope _err_scope(m);
  THY_TRY(require_device());
  return eval_batch_dev(m, B, X_dev, logp_dev, grad_dev, false);
}

thy_status thy_model_enable_timing(thy_model_t m, int enable) {
  thy::ErrScope _err_scope(m);
  if (!m) return fail(THY_ERR_INVALID_ARGUMENT, "null model handle");
  m->timing = enable != 0;
  return THY_OK;
}

thy_status thy_model_kernel_time(thy_model_t m, double* out_total_ms, int64_t* out_launches) {
  thy::ErrScope _err_scope(m);
  if (!m) return fail(THY_ERR_INVALID_ARGUMENT, "null model handle");
  double total = 0.0;
  int64_t n = 0;
  for (auto& ev : m->timing_events) {
    float ms = 0.f;
    THY_CUDA_CHECK(cudaEventSynchronize(ev.second));
    THY_CUDA_CHECK(cudaEventElapsedTime(&ms, ev.first, ev.second));
    total += ms;
    ++n;
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  m->timing_events.clear();
  if (out_total_ms) *out_total_ms = total;
  if (out_launches) *out_launches = n;
  return THY_OK;
}

}  // extern "C"
