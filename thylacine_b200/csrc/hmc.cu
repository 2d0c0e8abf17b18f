// Batched Hamiltonian MC: B independent chains in lock-step, state resident on the device.
// Reference: /root/reference/src/main/scala/thylacine/model/sampling/hmcmc/HmcmcEngine.scala
//   runLeapfrogAt (:55-80)      p -= eps/2 g ; x += eps p ; g = -grad logp(x) ; p -= eps/2 g        (L times)
//   getHamiltonianValue (:82-83) H = p.p/2 + E,  E = -logp
//   runDynamicSimulationFrom (:85-172)  fresh p ~ N(0,I) per trajectory; accept iff dH < 0 || U < exp(-dH) (:128);
//                                iterationCount <= maxIterations => maxIterations + 1 trajectories per block (:168, Q5);
//                                samples are a Set: a block in which nothing was accepted adds no sample (:149-152)
//   burnIn (:180-197)           warmupStepCount + 1 trajectories from the seed or a prior draw, re-run per sample() call
// The elementwise updates, the kinetic-energy reductions (warp shuffles) and the Metropolis test run in three small
// kernels fused around the gradient evaluation; all randomness is Philox4x32-10 keyed (seed, chain, item, trajectory),
// with the trajectory counter held in device memory so a whole trajectory can be replayed from a CUDA graph.
#include "kernels.h"
#include "philox.cuh"
#include <utility>

struct thy_hmc_s {
  thy_model_s* model = nullptr;
  thy_hmcmc_config cfg{};
  int64_t B = 0;
  int d = 0;
  uint64_t seed = 0;
  bool has_seed_points = false;
  thy::DevBuf<double> seed_points;                // [B][d] when provided
  thy::DevBuf<double> x, p, g, xn, gn;            // [B][d]; g, gn hold grad logp (not negated)
  thy::DevBuf<double> xn2;                        // second position buffer of the fused leapfrog step (ping-pong)
  int fused_steps = 1;                            // 0 never, 1 by cost model, 2 whenever the model allows it
  thy::DevBuf<double> lp, lpn, H0, dH;            // [B]
  thy::DevBuf<long long> attempts, acceptances, block_acc, emitted;  // [B]
  thy::DevBuf<unsigned long long> traj;           // [1] trajectory counter (device)
  thy::DevBuf<double> out;                        // sample buffer [B][n][d]
  thy::DevBuf<long long> min_emitted;             // [1]
  uint64_t init_epoch = 0;                        // distinguishes prior draws of successive sample() calls
  bool initialised = false;
  cudaGraphExec_t graph = nullptr;                // one trajectory
  uint64_t graph_ws_epoch = 0;                    // model workspace epoch and stream the graph was captured with
  cudaStream_t graph_stream = nullptr;
  uint64_t graph_launches = 0;                    // kernels inside the captured trajectory
  bool use_graph = true;
  ~thy_hmc_s() {
    if (graph) cudaGraphExecDestroy(graph);
  }
};

namespace thy {

// Trajectory start: p ~ N(0,I); H0 = p.p/2 - logp; first half kick and drift (into xn).
__global__ void __launch_bounds__(256) k_hmc_begin(long long B, int d, uint64_t seed, const unsigned long long* traj,
                                                   const double* __restrict__ x, const double* __restrict__ g,
                                                   const double* __restrict__ lp, double* __restrict__ p,
                                                   double* __restrict__ xn, double* __restrict__ H0, double eps) {
  const int lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const unsigned long long t = *traj;
  const double he = -eps / 2;
  double kin = 0.0;
  for (int j = lane; j < d; j += 32) {
    const double p0 = rng_normal(seed, (uint64_t)b, (uint32_t)j, RNG_HMC_MOMENTUM, t);
    kin += p0 * p0;
    // the reference works with -grad logp: pNew = p + gradNegLogPdf * (-eps/2) = p + (-g) * (-eps/2)
    const double pn = __dadd_rn(p0, __dmul_rn(-g[b * d + j], he));
    p[b * d + j] = pn;
    xn[b * d + j] = __dadd_rn(x[b * d + j], __dmul_rn(pn, eps));
  }
  kin = warp_sum(kin);
  if (lane == 0) H0[b] = kin / 2.0 + (-lp[b]);
}

// Between two gradient evaluations inside a trajectory: second half kick of step l, first half kick and drift of l+1.
__global__ void __launch_bounds__(256) k_hmc_mid(long long n, const double* __restrict__ gn, double* __restrict__ p,
                                                 double* __restrict__ xn, double eps) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double he = -eps / 2;
  const double kick = __dmul_rn(-gn[i], he);
  double pv = __dadd_rn(p[i], kick);
  pv = __dadd_rn(pv, kick);
  p[i] = pv;
  xn[i] = __dadd_rn(xn[i], __dmul_rn(pv, eps));
}

// Trajectory end: last half kick, H1, Metropolis test, state update, counters.
__global__ void __launch_bounds__(256) k_hmc_end(long long B, int d, uint64_t seed, unsigned long long* traj,
                                                 double* __restrict__ x, double* __restrict__ g, double* __restrict__ lp,
                                                 const double* __restrict__ p, const double* __restrict__ xn,
                                                 const double* __restrict__ gn, const double* __restrict__ lpn,
                                                 const double* __restrict__ H0, double* __restrict__ dH_out,
                                                 long long* __restrict__ attempts, long long* __restrict__ acceptances,
                                                 long long* __restrict__ block_acc, double eps) {
  const int lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const unsigned long long t = *traj;
  const double he = -eps / 2;
  double kin = 0.0;
  for (int j = lane; j < d; j += 32) {
    const double pv = __dadd_rn(p[b * d + j], __dmul_rn(-gn[b * d + j], he));
    kin += pv * pv;
  }
  kin = warp_sum(kin);
  const double H1 = kin / 2.0 + (-lpn[b]);
  const double dH = H1 - H0[b];
  const RngDraw u = rng_uniform2(seed, (uint64_t)b, 0, RNG_HMC_ACCEPT, t);
  const bool accept = (dH < 0) || (u.u0 < exp(-dH));   // NaN dH -> reject, as in the reference
  if (accept) {
    for (int j = lane; j < d; j += 32) {
      x[b * d + j] = xn[b * d + j];
      g[b * d + j] = gn[b * d + j];
    }
  }
  if (lane == 0) {
    if (accept) {
      lp[b] = lpn[b];
      acceptances[b] += 1;
      block_acc[b] += 1;
    }
    attempts[b] += 1;
    dH_out[b] = dH;
  }
}

__global__ void k_hmc_advance(unsigned long long* traj) { *traj += 1; }

// Sample emission with the reference's Set semantics: a chain emits after a block only if it moved (or has not
// emitted yet); chains that already hold n samples keep running but emit nothing.
__global__ void __launch_bounds__(256) k_hmc_emit(long long B, int d, long long n_per_chain, const double* __restrict__ x,
                                                  long long* __restrict__ block_acc, long long* __restrict__ emitted,
                                                  double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const long long k = emitted[b];
  const bool moved = block_acc[b] > 0 || k == 0;
  if (moved && k < n_per_chain) {
    double* dst = out + ((size_t)b * n_per_chain + k) * d;
    for (int j = lane; j < d; j += 32) dst[j] = x[b * d + j];
  }
  __syncwarp();
  if (lane == 0) {
    if (moved && k < n_per_chain) emitted[b] = k + 1;
    block_acc[b] = 0;
  }
}

__global__ void k_min_ll(long long B, const long long* __restrict__ v, long long* __restrict__ out) {
  // single block
  __shared__ long long sm[256];
  long long m = (1LL << 62);
  for (long long i = threadIdx.x; i < B; i += blockDim.x) m = v[i] < m ? v[i] : m;
  sm[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] = sm[threadIdx.x + o] < sm[threadIdx.x] ? sm[threadIdx.x + o] : sm[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sm[0];
}

static thy_status hmc_trajectory_launch(thy_hmc_s* h) {
  thy_model_s* m = h->model;
  cudaStream_t s = m->stream;
  const long long B = h->B;
  const int d = h->d;
  const int L = h->cfg.stepsInDynamicsSimulation;
  // no dynamics (L == 0): the reference returns (input, rawP) unchanged, so dH == 0 and the proposal is always accepted
  // (HmcmcEngine.scala:55-80 with an empty fold): no kick, no drift
  const double eps = L > 0 ? h->cfg.dynamicsSimulationStepSize : 0.0;
  const unsigned wblocks = (unsigned)ceil_div(B, 8);
  k_hmc_begin<<<wblocks, 256, 0, s>>>(B, d, h->seed, h->traj.ptr, h->x.ptr, h->g.ptr, h->lp.ptr, h->p.ptr, h->xn.ptr,
                                      h->H0.ptr, eps);
  count_launch();
  // Folding the step into the gradient kernel saves the prior pass and the elementwise kick / drift kernel (two launches,
  // ~64 bytes per coordinate of traffic) but makes the kernel's output step heavier, and the warps of a CTA reach that
  // step together: measured on B200 the fused kernel is ~2 % slower than the plain one (config 2: 558 vs 547 us per
  // step), so it wins while 2 % of the gradient kernel is less than the two small kernels it replaces (config 1 /
  // collapsed models: 17 vs 19 us per step) -- fused_steps == 1 applies that rule, 2 forces the single launch.
  bool fuse = L > 0 && h->fused_steps > 0 && hmc_step_fusable(m, B);
  if (fuse && h->fused_steps == 1) {
    const double t_small = 2 * 4e-6 + (double)B * d * 64.0 / 6e12;
    fuse = 0.02 * fused_cost_model(m->lik_dev[0]->view, B) < t_small;
  }
  if (fuse) {
    // ONE launch per leapfrog step: kicks / drift / Metropolis test live in the gradient kernel's output step
    THY_TRY(h->xn2.reserve((size_t)B * d));
    double* cur = h->xn.ptr;
    double* nxt = h->xn2.ptr;
    for (int l = 0; l < L; ++l) {
      HmcFuse hf;
      hf.mode = (l + 1 < L) ? HMC_FUSE_MID : HMC_FUSE_LAST;
      hf.eps = eps;
      hf.p = h->p.ptr;
      hf.x_next = nxt;
      hf.x = h->x.ptr;
      hf.g = h->g.ptr;
      hf.lp = h->lp.ptr;
      hf.H0 = h->H0.ptr;
      hf.dH = h->dH.ptr;
      hf.attempts = h->attempts.ptr;
      hf.acceptances = h->acceptances.ptr;
      hf.block_acc = h->block_acc.ptr;
      hf.seed = h->seed;
      hf.traj = h->traj.ptr;
      THY_TRY(eval_batch_hmc(m, B, cur, h->lpn.ptr, h->gn.ptr, hf));
      if (l + 1 < L) std::swap(cur, nxt);
    }
    k_hmc_advance<<<1, 1, 0, s>>>(h->traj.ptr);
    count_launch();
    THY_CUDA_CHECK(cudaGetLastError());
    return THY_OK;
  }
  if (L <= 0) {
    // the proposal is the current point; evaluate it so lpn / gn are defined
    THY_TRY(eval_batch_dev(m, B, h->x.ptr, h->lpn.ptr, h->gn.ptr, true));
    THY_CUDA_CHECK(cudaMemcpyAsync(h->xn.ptr, h->x.ptr, (size_t)B * d * sizeof(double), cudaMemcpyDeviceToDevice, s));
  }
  for (int l = 0; l < L; ++l) {
    THY_TRY(eval_batch_dev(m, B, h->xn.ptr, h->lpn.ptr, h->gn.ptr, true));
    if (l + 1 < L) {
      k_hmc_mid<<<(unsigned)ceil_div(B * d, 256), 256, 0, s>>>(B * d, h->gn.ptr, h->p.ptr, h->xn.ptr, eps);
      count_launch();
    }
  }
  k_hmc_end<<<wblocks, 256, 0, s>>>(B, d, h->seed, h->traj.ptr, h->x.ptr, h->g.ptr, h->lp.ptr, h->p.ptr, h->xn.ptr, h->gn.ptr,
                                    h->lpn.ptr, h->H0.ptr, h->dH.ptr, h->attempts.ptr, h->acceptances.ptr, h->block_acc.ptr,
                                    eps);
  k_hmc_advance<<<1, 1, 0, s>>>(h->traj.ptr);
  count_launch(2);
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

// One trajectory; captured once into a CUDA graph (launch-bound for small models) and replayed afterwards.
static thy_status hmc_trajectory(thy_hmc_s* h) {
  cudaStream_t s = h->model->stream;
  if (!h->use_graph || s == 0 /* legacy stream cannot be captured */) return hmc_trajectory_launch(h);
  if (h->graph && (h->graph_ws_epoch != h->model->ws.epoch || h->graph_stream != s)) {
    // another evaluation on this model grew a workspace buffer (or the stream changed): the captured addresses are stale
    cudaGraphExecDestroy(h->graph);
    h->graph = nullptr;
  }
  if (!h->graph) {
    // warm every code path (workspace allocation, cudaFuncSetAttribute) outside capture
    THY_TRY(hmc_trajectory_launch(h));
    THY_CUDA_CHECK(cudaStreamSynchronize(s));
    cudaGraph_t graph = nullptr;
    const uint64_t before = g_kernel_launches.load();
    THY_CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    thy_status st = hmc_trajectory_launch(h);
    cudaError_t ce = cudaStreamEndCapture(s, &graph);
    h->graph_launches = g_kernel_launches.load() - before;
    g_kernel_launches.store(before);  // captured, not executed
    if (st != THY_OK || ce != cudaSuccess || !graph) {
      cudaGetLastError();
      h->use_graph = false;
      if (graph) cudaGraphDestroy(graph);
      return st != THY_OK ? st : hmc_trajectory_launch(h);
    }
    ce = cudaGraphInstantiate(&h->graph, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) {
      cudaGetLastError();
      h->graph = nullptr;
      h->use_graph = false;
      return hmc_trajectory_launch(h);
    }
    h->graph_ws_epoch = h->model->ws.epoch;
    h->graph_stream = s;
    return THY_OK;  // the warm-up launch above was this call's trajectory
  }
  THY_CUDA_CHECK(cudaGraphLaunch(h->graph, s));
  count_launch(h->graph_launches);
  return THY_OK;
}

static thy_status hmc_block(thy_hmc_s* h, int max_iterations) {
  for (int it = 0; it <= max_iterations; ++it) THY_TRY(hmc_trajectory(h));  // iterationCount <= maxIterations (Q5)
  return THY_OK;
}

static thy_status hmc_initialise(thy_hmc_s* h) {
  thy_model_s* m = h->model;
  cudaStream_t s = m->stream;
  const size_t nx = (size_t)h->B * h->d;
  if (h->has_seed_points) {
    THY_CUDA_CHECK(cudaMemcpyAsync(h->x.ptr, h->seed_points.ptr, nx * sizeof(double), cudaMemcpyDeviceToDevice, s));
  } else {
    THY_TRY(sample_priors_dev(m, h->B, h->seed ^ (0x9E3779B97F4A7C15ull * (h->init_epoch + 1)), h->x.ptr));
  }
  h->init_epoch += 1;
  THY_TRY(eval_batch_dev(m, h->B, h->x.ptr, h->lp.ptr, h->g.ptr, true));
  THY_CUDA_CHECK(cudaMemsetAsync(h->block_acc.ptr, 0, h->B * sizeof(long long), s));
  h->initialised = true;
  return THY_OK;
}

}  // namespace thy

using namespace thy;

extern "C" {

thy_status thy_hmc_create(thy_model_t m, const thy_hmcmc_config* cfg, int64_t n_chains, uint64_t rng_seed,
                          const double* seed_points, thy_hmc_t* out) {
  thy::ErrScope _err_scope(m);
  THY_TRY(require_device());
  if (!m || !m->finalized) return fail(THY_ERR_STATE, "model is not finalized");
  if (!cfg || !out) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  if (n_chains <= 0) return fail(THY_ERR_INVALID_ARGUMENT, "n_chains must be positive");
  if (cfg->stepsBetweenSamples < 0 || cfg->stepsInDynamicsSimulation < 0 || cfg->warmupStepCount < 0)
    return fail(THY_ERR_INVALID_ARGUMENT, "negative HmcmcConfig field");
  auto* h = new (std::nothrow) thy_hmc_s();
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "allocation failed");
  h->model = m;
  h->cfg = *cfg;
  h->B = n_chains;
  h->d = m->d;
  h->seed = rng_seed;
  const size_t nx = (size_t)n_chains * m->d;
  thy_status st = THY_OK;
  auto chk = [&](thy_status s2) { if (st == THY_OK) st = s2; };
  chk(h->x.reserve(nx)); chk(h->p.reserve(nx)); chk(h->g.reserve(nx)); chk(h->xn.reserve(nx)); chk(h->gn.reserve(nx));
  chk(h->lp.reserve(n_chains)); chk(h->lpn.reserve(n_chains)); chk(h->H0.reserve(n_chains)); chk(h->dH.reserve(n_chains));
  chk(h->attempts.reserve(n_chains)); chk(h->acceptances.reserve(n_chains)); chk(h->block_acc.reserve(n_chains));
  chk(h->emitted.reserve(n_chains)); chk(h->traj.reserve(1)); chk(h->min_emitted.reserve(1));
  if (st == THY_OK && seed_points) {
    h->has_seed_points = true;
    chk(h->seed_points.reserve(nx));
    if (st == THY_OK && cudaMemcpy(h->seed_points.ptr, seed_points, nx * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
      st = fail(THY_ERR_CUDA, "seed point upload failed");
  }
  if (st == THY_OK) {
    cudaMemset(h->attempts.ptr, 0, n_chains * sizeof(long long));
    cudaMemset(h->acceptances.ptr, 0, n_chains * sizeof(long long));
    cudaMemset(h->block_acc.ptr, 0, n_chains * sizeof(long long));
    cudaMemset(h->emitted.ptr, 0, n_chains * sizeof(long long));
    cudaMemset(h->traj.ptr, 0, sizeof(unsigned long long));
    cudaMemset(h->dH.ptr, 0, n_chains * sizeof(double));
  }
  if (st != THY_OK) {
    delete h;
    return st;
  }
  *out = h;
  return THY_OK;
}

thy_status thy_hmc_destroy(thy_hmc_t h) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  delete h;
  return THY_OK;
}

thy_status thy_hmc_set_fused_steps(thy_hmc_t h, int enable) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  h->fused_steps = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
  if (h->graph) {   // the captured trajectory belongs to the other mode
    cudaGraphExecDestroy(h->graph);
    h->graph = nullptr;
  }
  return THY_OK;
}

thy_status thy_hmc_set_state(thy_hmc_t h, const double* x, uint64_t trajectory_counter) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h || !x) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  thy_model_s* m = h->model;
  cudaStream_t s = m->stream;
  const unsigned long long t = trajectory_counter;
  THY_CUDA_CHECK(cudaMemcpyAsync(h->x.ptr, x, (size_t)h->B * h->d * sizeof(double), cudaMemcpyHostToDevice, s));
  THY_CUDA_CHECK(cudaMemcpyAsync(h->traj.ptr, &t, sizeof(t), cudaMemcpyHostToDevice, s));
  THY_TRY(eval_batch_dev(m, h->B, h->x.ptr, h->lp.ptr, h->g.ptr, true));
  THY_CUDA_CHECK(cudaMemsetAsync(h->block_acc.ptr, 0, h->B * sizeof(long long), s));
  THY_CUDA_CHECK(cudaStreamSynchronize(s));   // x and t are host stack / caller memory
  h->initialised = true;
  return THY_OK;
}

thy_status thy_hmc_get_trajectory_counter(thy_hmc_t h, uint64_t* out_counter) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h || !out_counter) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  unsigned long long t = 0;
  THY_CUDA_CHECK(cudaStreamSynchronize(h->model->stream));
  THY_CUDA_CHECK(cudaMemcpy(&t, h->traj.ptr, sizeof(t), cudaMemcpyDeviceToHost));
  *out_counter = t;
  return THY_OK;
}

thy_status thy_hmc_set_use_graph(thy_hmc_t h, int enable) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  h->use_graph = enable != 0;
  return THY_OK;
}

thy_status thy_hmc_burn_in(thy_hmc_t h) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  THY_TRY(hmc_initialise(h));
  THY_TRY(hmc_block(h, h->cfg.warmupStepCount));
  THY_CUDA_CHECK(cudaMemsetAsync(h->block_acc.ptr, 0, h->B * sizeof(long long), h->model->stream));
  return THY_OK;
}

thy_status thy_hmc_run_trajectories(thy_hmc_t h, int n_trajectories) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (!h->initialised) THY_TRY(hmc_initialise(h));
  for (int i = 0; i < n_trajectories; ++i) THY_TRY(hmc_trajectory(h));
  return THY_OK;
}

static thy_status hmc_sample_impl(thy_hmc_t h, int n_per_chain, double* out_host, double* out_dev, int rerun_burn_in) {
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (n_per_chain <= 0) return fail(THY_ERR_INVALID_ARGUMENT, "n_per_chain must be positive");
  if (!out_host && !out_dev) return fail(THY_ERR_INVALID_ARGUMENT, "null output pointer");
  cudaStream_t s = h->model->stream;
  if (rerun_burn_in || !h->initialised) THY_TRY(thy_hmc_burn_in(h));
  const size_t total = (size_t)h->B * n_per_chain * h->d;
  double* dst = out_dev;
  if (!dst) {
    THY_TRY(h->out.reserve(total));
    dst = h->out.ptr;
  }
  THY_CUDA_CHECK(cudaMemsetAsync(h->emitted.ptr, 0, h->B * sizeof(long long), s));
  const unsigned wblocks = (unsigned)ceil_div(h->B, 8);
  long long guard = 0;
  const long long max_blocks = 1000LL * n_per_chain + 1000;
  while (true) {
    THY_TRY(hmc_block(h, h->cfg.stepsBetweenSamples));
    k_hmc_emit<<<wblocks, 256, 0, s>>>(h->B, h->d, n_per_chain, h->x.ptr, h->block_acc.ptr, h->emitted.ptr, dst);
    count_launch();
    ++guard;
    if (guard >= n_per_chain) {
      long long mn = 0;
      k_min_ll<<<1, 256, 0, s>>>(h->B, h->emitted.ptr, h->min_emitted.ptr);
      count_launch();
      THY_CUDA_CHECK(cudaMemcpyAsync(&mn, h->min_emitted.ptr, sizeof(long long), cudaMemcpyDeviceToHost, s));
      THY_CUDA_CHECK(cudaStreamSynchronize(s));
      if (mn >= n_per_chain) break;
      if (guard > max_blocks) return fail(THY_ERR_STATE, "HMC: a chain never accepts; check dynamicsSimulationStepSize");
    }
  }
  if (out_host) {
    THY_CUDA_CHECK(cudaMemcpyAsync(out_host, dst, total * sizeof(double), cudaMemcpyDeviceToHost, s));
    THY_CUDA_CHECK(cudaStreamSynchronize(s));
  }
  return THY_OK;
}

thy_status thy_hmc_sample(thy_hmc_t h, int n_per_chain, double* out_samples, int rerun_burn_in) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  return hmc_sample_impl(h, n_per_chain, out_samples, nullptr, rerun_burn_in);
}

thy_status thy_hmc_sample_dev(thy_hmc_t h, int n_per_chain, double* out_samples_dev, int rerun_burn_in) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  return hmc_sample_impl(h, n_per_chain, nullptr, out_samples_dev, rerun_burn_in);
}

thy_status thy_hmc_stats(thy_hmc_t h, int64_t* out_attempts, int64_t* out_acceptances, double* out_last_dH) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  cudaStream_t s = h->model->stream;
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  static_assert(sizeof(long long) == sizeof(int64_t), "int64");
  if (out_attempts) THY_CUDA_CHECK(cudaMemcpy(out_attempts, h->attempts.ptr, h->B * sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (out_acceptances)
    THY_CUDA_CHECK(cudaMemcpy(out_acceptances, h->acceptances.ptr, h->B * sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (out_last_dH) THY_CUDA_CHECK(cudaMemcpy(out_last_dH, h->dH.ptr, h->B * sizeof(double), cudaMemcpyDeviceToHost));
  return THY_OK;
}

thy_status thy_hmc_get_state(thy_hmc_t h, double* out_x, double* out_logp) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  THY_CUDA_CHECK(cudaStreamSynchronize(h->model->stream));
  if (out_x) THY_CUDA_CHECK(cudaMemcpy(out_x, h->x.ptr, (size_t)h->B * h->d * sizeof(double), cudaMemcpyDeviceToHost));
  if (out_logp) THY_CUDA_CHECK(cudaMemcpy(out_logp, h->lp.ptr, h->B * sizeof(double), cudaMemcpyDeviceToHost));
  return THY_OK;
}

}  // extern "C"
