// Leapfrog-MCMC: the reference's gradient-free ensemble sampler, re-expressed as batched half-pool sweeps.
// Reference: /root/reference/src/main/scala/thylacine/model/sampling/leapfrog/LeapfrogMcmcEngine.scala
//   getLeapfrogVector (:88-89)        x' = 2 x_j - x_i
//   generateProposal (:116-162)       i uniform; j ~ Categorical(d(i,.) / sum d(i,.)) by the CDF staircase
//                                     (util/MathOps.scala:76-91); q_rev = d(x',x_j) / sum_{k != i} d(x',x_k) (:91-111);
//                                     a = exp(l(x') - l_i) q_rev / q_fwd; accept iff a >= 1 || U < a (:133-138)
//   populateSamples / initialisePool  P prior draws (slot 0 <- seed), P logPdfs (:224-253)
//   burnInRecursion (:172-182)        until every slot has been proposed warmupStepCount times
// The reference is strictly one proposal at a time with a P x P distance table (34 GB at P = 65 536).  Here the pool is
// split into two halves; every member of the active half is updated in parallel against the FROZEN other half:
// partners j and both normalising sums range over the frozen half only, so each update is a Metropolis-Hastings kernel
// for pi(x_i) given the frozen set and the Hastings ratio stays exact.  Distance rows are recomputed on the fly
// (no P x P table).  One sweep = both halves = P proposals, one per pool member (systematic scan instead of random scan).
#include "kernels.h"
#include "philox.cuh"

struct thy_lfm_s {
  thy_model_s* model = nullptr;
  thy_leapfrog_mcmc_config cfg{};
  int distance = 0;
  int P = 0, d = 0;
  uint64_t seed = 0;
  thy::DevBuf<double> x, lp;          // pool [P][d], logPdfs [P]
  thy::DevBuf<double> xp, lpp;        // proposals for the active half
  thy::DevBuf<double> D;              // distance rows [nH][nS]
  thy::DevBuf<double> qf, qr;         // forward / reverse proposal probabilities
  thy::DevBuf<int> partner;
  thy::DevBuf<unsigned long long> counters;  // [0] proposals, [1] acceptances
  uint64_t sweep = 0;
  int emit_cursor = 0;                // next pool slot handed out by sample()
  int sweeps_since_emit = 0;
};

namespace thy {

constexpr int DT = 64, DK = 16;

__device__ __forceinline__ double dist_finish(int kind, double s) { return kind == 0 ? sqrt(s) : s; }

// D[i][k] = distance(Y_i, S_k); Y: nY x d, S: nS x d (row-major), tiled 64 x 64, direct difference form
// (no |a|^2 + |b|^2 - 2ab cancellation).
__global__ void __launch_bounds__(256) k_lfm_dist(int kind, int d, int nY, int nS, const double* __restrict__ Y,
                                                  const double* __restrict__ S, double* __restrict__ D) {
  __shared__ double Ys[DT][DK + 1];
  __shared__ double Ss[DT][DK + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * DT, k0 = blockIdx.x * DT;
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
  for (int j0 = 0; j0 < d; j0 += DK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = threadIdx.x + 256 * e;
      const int row = idx >> 4, jj = idx & 15;
      const int j = j0 + jj;
      Ys[row][jj] = (j < d && i0 + row < nY) ? Y[(size_t)(i0 + row) * d + j] : 0.0;
      Ss[row][jj] = (j < d && k0 + row < nS) ? S[(size_t)(k0 + row) * d + j] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < DK; ++jj) {
      double yr[4], sr[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) yr[r] = Ys[ty * 4 + r][jj];
#pragma unroll
      for (int c = 0; c < 4; ++c) sr[c] = Ss[tx + 16 * c][jj];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double df = yr[r] - sr[c];
          acc[r][c] = (kind == 2) ? acc[r][c] + fabs(df) : fma(df, df, acc[r][c]);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i >= nY) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int k = k0 + tx + 16 * c;
      if (k < nS) D[(size_t)i * nS + k] = dist_finish(kind, acc[r][c]);
    }
  }
}

__device__ __forceinline__ double block_sum_256(double v, double* sm) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double t = (threadIdx.x < 8) ? sm[threadIdx.x] : 0.0;
  if (w == 0) {
    t += __shfl_xor_sync(0xffffffffu, t, 4);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    if (lane == 0) sm[8] = t;
  }
  __syncthreads();
  return sm[8];
}

// One block per active member i: total = sum_k D[i][k]; partner j = first k with cdf_k > u * total (staircase
// lo <= u < hi); q_fwd = D[i][j] / total; proposal x' = 2 S_j - x_i.
__global__ void __launch_bounds__(256) k_lfm_pick(int d, int nH, int nS, uint64_t seed, uint64_t sweep, int base,
                                                  const double* __restrict__ D, const double* __restrict__ XH,
                                                  const double* __restrict__ S, int* __restrict__ partner,
                                                  double* __restrict__ qf, double* __restrict__ XP) {
  __shared__ double sm[16];
  __shared__ double chunk_sum[256];
  __shared__ int j_sh;
  const int i = blockIdx.x;
  const double* row = D + (size_t)i * nS;
  const int per = (nS + 255) / 256;
  const int k_lo = threadIdx.x * per, k_hi = min(nS, k_lo + per);
  double local = 0.0;
  for (int k = k_lo; k < k_hi; ++k) local += row[k];
  chunk_sum[threadIdx.x] = local;
  const double total = block_sum_256(local, sm);
  if (threadIdx.x == 0) j_sh = -1;
  __syncthreads();
  if (threadIdx.x == 0 && total > 0.0 && isfinite(total)) {
    const RngDraw u = rng_uniform2(seed, (uint64_t)(base + i), 0, RNG_LFM_PARTNER, sweep);
    const double target = u.u0 * total;
    double run = 0.0;
    int t = 0;
    for (; t < 256; ++t) {
      if (run + chunk_sum[t] > target) break;
      run += chunk_sum[t];
    }
    int j = -1;
    if (t < 256) {
      const int a = t * per, b = min(nS, a + per);
      for (int k = a; k < b; ++k) {
        run += row[k];
        if (run > target) {
          j = k;
          break;
        }
      }
    }
    if (j < 0) {  // rounding at the top of the staircase: last entry with positive width
      for (int k = nS - 1; k >= 0; --k)
        if (row[k] > 0.0) {
          j = k;
          break;
        }
    }
    j_sh = j;
    partner[i] = j;
    qf[i] = (j >= 0) ? row[j] / total : 0.0;
  } else if (threadIdx.x == 0) {
    partner[i] = -1;
    qf[i] = 0.0;
  }
  __syncthreads();
  const int j = j_sh;
  for (int c = threadIdx.x; c < d; c += 256) {
    const double xi = XH[(size_t)i * d + c];
    XP[(size_t)i * d + c] = (j >= 0) ? (2.0 * S[(size_t)j * d + c] - xi) : xi;
  }
}

// Reverse proposal probability from the distance rows of the proposals.
__global__ void __launch_bounds__(256) k_lfm_rev(int nH, int nS, const double* __restrict__ D, const int* __restrict__ partner,
                                                 double* __restrict__ qr) {
  __shared__ double sm[16];
  const int i = blockIdx.x;
  const double* row = D + (size_t)i * nS;
  double local = 0.0;
  for (int k = threadIdx.x; k < nS; k += 256) local += row[k];
  const double total = block_sum_256(local, sm);
  if (threadIdx.x == 0) {
    const int j = partner[i];
    qr[i] = (j >= 0 && total > 0.0) ? row[j] / total : 0.0;
  }
}

// Metropolis-Hastings test and in-place update of the active half.
__global__ void __launch_bounds__(256) k_lfm_accept(int d, int nH, uint64_t seed, uint64_t sweep, int base,
                                                    double* __restrict__ XH, double* __restrict__ lpH,
                                                    const double* __restrict__ XP, const double* __restrict__ lpP,
                                                    const int* __restrict__ partner, const double* __restrict__ qf,
                                                    const double* __restrict__ qr, unsigned long long* __restrict__ counters) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= nH) return;
  bool accept = false;
  if (partner[i] >= 0 && qf[i] > 0.0) {
    const double a = exp(lpP[i] - lpH[i]) * qr[i] / qf[i];
    const RngDraw u = rng_uniform2(seed, (uint64_t)(base + i), 0, RNG_LFM_ACCEPT, sweep);
    accept = (a >= 1.0) || (u.u0 < a);
  }
  if (accept) {
    for (int c = lane; c < d; c += 32) XH[(size_t)i * d + c] = XP[(size_t)i * d + c];
  }
  if (lane == 0) {
    if (accept) {
      lpH[i] = lpP[i];
      atomicAdd(&counters[1], 1ULL);
    }
    atomicAdd(&counters[0], 1ULL);
  }
}

static thy_status lfm_half_sweep(thy_lfm_s* h, int half) {
  thy_model_s* m = h->model;
  cudaStream_t s = m->stream;
  const int d = h->d;
  const int n0 = h->P / 2, n1 = h->P - n0;
  const int nH = half == 0 ? n0 : n1, nS = half == 0 ? n1 : n0;
  double* XH = h->x.ptr + (half == 0 ? 0 : (size_t)n0 * d);
  double* lpH = h->lp.ptr + (half == 0 ? 0 : n0);
  const double* S = h->x.ptr + (half == 0 ? (size_t)n0 * d : 0);
  dim3 gd((unsigned)ceil_div(nS, DT), (unsigned)ceil_div(nH, DT));
  k_lfm_dist<<<gd, 256, 0, s>>>(h->distance, d, nH, nS, XH, S, h->D.ptr);
  k_lfm_pick<<<nH, 256, 0, s>>>(d, nH, nS, h->seed, h->sweep, half == 0 ? 0 : n0, h->D.ptr, XH, S, h->partner.ptr, h->qf.ptr, h->xp.ptr);
  k_lfm_dist<<<gd, 256, 0, s>>>(h->distance, d, nH, nS, h->xp.ptr, S, h->D.ptr);
  k_lfm_rev<<<nH, 256, 0, s>>>(nH, nS, h->D.ptr, h->partner.ptr, h->qr.ptr);
  count_launch(4);
  THY_TRY(eval_batch_dev(m, nH, h->xp.ptr, h->lpp.ptr, nullptr, true));
  k_lfm_accept<<<(unsigned)ceil_div(nH, 8), 256, 0, s>>>(d, nH, h->seed, h->sweep, half == 0 ? 0 : n0, XH, lpH, h->xp.ptr, h->lpp.ptr,
                                                         h->partner.ptr, h->qf.ptr, h->qr.ptr, h->counters.ptr);
  count_launch();
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

static thy_status lfm_sweep(thy_lfm_s* h) {
  THY_TRY(lfm_half_sweep(h, 0));
  THY_TRY(lfm_half_sweep(h, 1));
  h->sweep += 1;
  return THY_OK;
}

}  // namespace thy

using namespace thy;

extern "C" {

thy_status thy_lfm_create(thy_model_t m, const thy_leapfrog_mcmc_config* cfg, int distance, uint64_t rng_seed,
                          const double* seed_point, thy_lfm_t* out) {
  thy::ErrScope _err_scope(m);
  THY_TRY(require_device());
  if (!m || !m->finalized) return fail(THY_ERR_STATE, "model is not finalized");
  if (!cfg || !out) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  if (cfg->samplePoolSize < 2) return fail(THY_ERR_INVALID_ARGUMENT, "samplePoolSize must be >= 2");
  if (cfg->stepsBetweenSamples < 1 || cfg->warmupStepCount < 0) return fail(THY_ERR_INVALID_ARGUMENT, "bad LeapfrogMcmcConfig");
  if (distance < THY_DISTANCE_EUCLIDEAN || distance > THY_DISTANCE_MANHATTAN)
    return fail(THY_ERR_INVALID_ARGUMENT, "unknown distance (a JVM closure cannot run on the device; see thy_distance)");
  auto* h = new (std::nothrow) thy_lfm_s();
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "allocation failed");
  h->model = m;
  h->cfg = *cfg;
  h->distance = distance;
  h->P = cfg->samplePoolSize;
  h->d = m->d;
  h->seed = rng_seed;
  const int P = h->P, d = h->d;
  const int nmax = P - P / 2;
  thy_status st = THY_OK;
  auto chk = [&](thy_status s2) { if (st == THY_OK) st = s2; };
  chk(h->x.reserve((size_t)P * d)); chk(h->lp.reserve(P)); chk(h->xp.reserve((size_t)nmax * d)); chk(h->lpp.reserve(nmax));
  chk(h->D.reserve((size_t)nmax * nmax)); chk(h->qf.reserve(nmax)); chk(h->qr.reserve(nmax)); chk(h->partner.reserve(nmax));
  chk(h->counters.reserve(2));
  cudaStream_t s = m->stream;
  if (st == THY_OK) chk(sample_priors_dev(m, P, rng_seed ^ 0xA5A5A5A5DEADBEEFull, h->x.ptr));   // populateSamples
  if (st == THY_OK && seed_point &&
      cudaMemcpyAsync(h->x.ptr, seed_point, d * sizeof(double), cudaMemcpyHostToDevice, s) != cudaSuccess)
    st = fail(THY_ERR_CUDA, "seed upload failed");
  if (st == THY_OK) cudaMemsetAsync(h->counters.ptr, 0, 2 * sizeof(unsigned long long), s);
  if (st == THY_OK) chk(eval_batch_dev(m, P, h->x.ptr, h->lp.ptr, nullptr, true));               // initialisePool
  for (int w = 0; st == THY_OK && w < cfg->warmupStepCount; ++w) chk(lfm_sweep(h));              // burn-in, as `.of` does
  if (st == THY_OK && cudaStreamSynchronize(s) != cudaSuccess) st = fail(THY_ERR_CUDA, "burn-in failed");
  if (st != THY_OK) {
    delete h;
    return st;
  }
  *out = h;
  return THY_OK;
}

thy_status thy_lfm_destroy(thy_lfm_t h) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  delete h;
  return THY_OK;
}

thy_status thy_lfm_run_sweeps(thy_lfm_t h, int n_sweeps) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  for (int i = 0; i < n_sweeps; ++i) THY_TRY(lfm_sweep(h));
  return THY_OK;
}

thy_status thy_lfm_sample(thy_lfm_t h, int n, double* out_samples) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (n <= 0 || !out_samples) return fail(THY_ERR_INVALID_ARGUMENT, "bad arguments");
  cudaStream_t s = h->model->stream;
  int done = 0;
  while (done < n) {
    // a pool member becomes a sample once it has been proposed stepsBetweenSamples times since it was last handed out
    if (h->emit_cursor == 0 || h->emit_cursor >= h->P) {
      for (int i = 0; i < h->cfg.stepsBetweenSamples; ++i) THY_TRY(lfm_sweep(h));
      h->emit_cursor = 0;
    }
    const int take = std::min(n - done, h->P - h->emit_cursor);
    THY_CUDA_CHECK(cudaMemcpyAsync(out_samples + (size_t)done * h->d, h->x.ptr + (size_t)h->emit_cursor * h->d,
                                   (size_t)take * h->d * sizeof(double), cudaMemcpyDeviceToHost, s));
    h->emit_cursor += take;
    done += take;
  }
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  return THY_OK;
}

thy_status thy_lfm_stats(thy_lfm_t h, uint64_t* out_proposals, uint64_t* out_acceptances) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  unsigned long long c[2];
  THY_CUDA_CHECK(cudaStreamSynchronize(h->model->stream));
  THY_CUDA_CHECK(cudaMemcpy(c, h->counters.ptr, sizeof(c), cudaMemcpyDeviceToHost));
  if (out_proposals) *out_proposals = c[0];
  if (out_acceptances) *out_acceptances = c[1];
  return THY_OK;
}

// Replace the whole pool (warm start near the mode, or a checkpointed pool): positions [samplePoolSize * d] from the
// host; their log-posteriors are re-evaluated on the device.
thy_status thy_lfm_set_pool(thy_lfm_t h, const double* x) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h || !x) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  cudaStream_t s = h->model->stream;
  THY_CUDA_CHECK(cudaMemcpyAsync(h->x.ptr, x, (size_t)h->P * h->d * sizeof(double), cudaMemcpyHostToDevice, s));
  THY_TRY(eval_batch_dev(h->model, h->P, h->x.ptr, h->lp.ptr, nullptr, true));
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  return THY_OK;
}

thy_status thy_lfm_get_pool(thy_lfm_t h, double* out_x, double* out_logp) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  THY_CUDA_CHECK(cudaStreamSynchronize(h->model->stream));
  if (out_x) THY_CUDA_CHECK(cudaMemcpy(out_x, h->x.ptr, (size_t)h->P * h->d * sizeof(double), cudaMemcpyDeviceToHost));
  if (out_logp) THY_CUDA_CHECK(cudaMemcpy(out_logp, h->lp.ptr, h->P * sizeof(double), cudaMemcpyDeviceToHost));
  return THY_OK;
}

}  // extern "C"
