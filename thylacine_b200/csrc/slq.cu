// SLQ (stochastic Lebesgue quadrature = nested sampling) with the live-point threshold step on the device.
// Reference: /root/reference/src/main/scala/thylacine/model/integration/slq/
//   SlqEngine.scala:120-197        live pool keyed by logPdf; discard the global minimum; replace it by a new draw
//                                  accepted only above the threshold; per-acceptance abscissa extension
//   SlqEngine.scala:249-277        convergence: iterations > minIterationCount && iterations >= 10 P H, or >= maxIterationCount
//   SlqEngine.scala:171-180        drainSamplePool: remaining live points appended in ascending order
//   QuadratureAbscissa.scala:33-73 prior-mass simulation: pool of P uniforms, max moved to the abscissa list (first = 1.0),
//                                  trapezoid weights 1/2 (X_{i-1} - X_{i+1}), ends half the adjacent gap
//   QuadratureIntegrator.scala:29-83  Z_a, H_a and getIntegrationStats per abscissa
//   SamplingSimulation.scala:41-101   posterior resampling of the retired points, one simulated abscissa per draw
// Re-design for 10^6 live points (DESIGN.md "SLQ"): the K lowest live points are retired per round (K = sampleParallelism,
// the reference's number of concurrent sampling fibres).  A round is
//     spread of the pool -> [all-gather of the ranks' log-likelihoods] -> radix SELECT of the K-th smallest (slq_select.cu,
//     no sort) -> retire flags / slots -> fused constrained walk (slq_walk.cu) -> scale adaptation,
// with the quadrature update (sort of the K retired values, prior-mass simulation, log-sum-exp; slq_evidence.cu) on a
// side stream beside the walk.  Nothing in a round needs the host: the retirement counter, the round number and the
// convergence verdict live in device memory, so a full round is ONE CUDA-graph launch and the host runs two rounds
// ahead of the device; it learns about convergence from a write-once device word with a fixed lag, which makes the
// stopping round identical on every rank (they all compute the same evidence from the same gathered keys).
// Measure / integrand: points are drawn from the PRIOR and ranked by log-LIKELIHOOD, so ln Z is the calibrated
// evidence (for uniform priors this coincides with the reference's uniform-domain / full-density formulation, Q6).
#include "slq_internal.h"
#include "comm.h"
#include <cub/cub.cuh>
#include <cmath>
#include <climits>
#include <algorithm>
#include <chrono>

struct thy_slq_s {
  thy_model_s* model = nullptr;
  thy_comm_s* comm = nullptr;
  thy_slq_config cfg{};
  uint64_t seed = 0;
  int world = 1, rank = 0;
  long long P = 0, Pl = 0;   // global / local live points
  int d = 0, A = 1, K = 1, M = 20;
  thy::DevBuf<double> x, ll, lpr;                        // live pool
  thy::DevBuf<double> keys_all;                          // [P] log-likelihoods of all ranks (world > 1)
  // threshold selection (slq_select.cu)
  thy::DevBuf<thy::SelectState> sel_state;
  thy::DevBuf<unsigned int> sel_hist, sel_ticket, sel_cursor;
  thy::DevBuf<int2> sel_cnt;
  thy::DevBuf<unsigned char> flags;
  thy::DevBuf<int> sidx, count_dev;
  thy::DevBuf<double> vals_raw, vals_sorted;             // retired values of the round, unordered / ascending
  thy::DevBuf<int> pair_in, pair_out;                    // slot ids beside the keys (drain / keepRetiredPoints)
  thy::DevBuf<unsigned char> cub_tmp, cub_tmp2;          // sort scratch: quadrature stream / main stream (pairs)
  thy::DevBuf<unsigned long long> bs_minmax;              // bucket sort of a round's retired values (slq_sort.cu)
  thy::DevBuf<unsigned int> bs_count, bs_cursor, bs_offset;
  thy::DevBuf<double> bs_tmp;
  thy::DevBuf<double> Y, Yp, yl, ypr, tl, tpr;           // walkers of the per-step path
  thy::DevBuf<int> slot, accepted;
  thy::DevBuf<double> sigma, sig_part;                   // per-dimension spread of the live pool
  thy::DevBuf<float> zpre;                               // pre-drawn walk normals (sharded pools: few walkers per rank)
  thy::DevBuf<double> lacc;                              // pre-drawn acceptance thresholds
  int pre_groups = 0;                                    // walker groups covered by zpre / lacc (0: drawn in the walk)
  // NVLink peer-memory exchange of new keys (slq_p2p.cu); off => one NCCL all-gather of all keys per round
  bool p2p = false;
  thy::DevBuf<unsigned char> p2p_buf;
  thy::DevBuf<unsigned long long> p2p_bases;
  std::vector<void*> p2p_opened;
  thy::P2pView p2p_view{};
  thy::DevBuf<double> scalars;                           // [0] walk scale, [1] threshold, [4] last rate, [5] accepted, [6] proposed
  // quadrature (slq_evidence.cu)
  thy::DevBuf<double> ev_lt, ev_segsum, ev_part, abs_state, abs_out;
  thy::DevBuf<unsigned int> ev_tickets;
  thy::DevBuf<double> ret_ll;                            // retired log-likelihoods (global order)
  thy::DevBuf<double> ret_x;                             // retired points (optional, world == 1)
  thy::DevBuf<long long> ctl;                            // [0] iterations, [1] round, [2] first converged round (-1: none)
  long long iterations = 0, rounds = 0, evals = 0;       // host mirrors of ctl[0], ctl[1]
  long long live_iterations = -1;                        // set by the drain
  long long ret_capacity = 0;
  bool fused_walk = false;             // slq_walk.cu covers this model: one kernel per round instead of 5 per walk step
  cudaStream_t side = nullptr;         // quadrature update runs beside the walk
  cudaStream_t side2 = nullptr;        // pool spread runs beside the all-gather and the selection
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_fork2 = nullptr, ev_sigma = nullptr, ev_rng = nullptr;
  cudaEvent_t ev_round[3] = {};        // completion of round r is ev_round[r % 3]
  double* host_ao = nullptr;           // pinned [A][2] copy of abs_out
  int* host_c = nullptr;               // pinned copy of the local retire count (per-step path)
  long long* host_conv = nullptr;      // pinned copy of ctl[2]
  cudaStream_t graph_stream = nullptr;
  cudaGraphExec_t graph = nullptr;     // one full round (Kr == K)
  uint64_t graph_launches = 0;
  bool use_graph = true;
  bool sort_sized = false;             // cub scratch sized (first round runs outside the graph)
  bool cub_round_sort = false;         // THY_SLQ_CUB_SORT=1: a round's retired values through cub (measurement knob)
  // phase trace (thy_slq_set_trace / THY_SLQ_TRACE=1): CUDA-event times of the round's phases, plain launches
  bool trace = false;
  cudaEvent_t tev[8] = {};             // [0..5] phase boundaries on the main stream, [6], [7] around the pool spread
  double tsum[7] = {0, 0, 0, 0, 0, 0, 0};
  long long tcount = 0;
  long long ref_rebuilds = 0;          // reference-faithful mode: cover rebuilds
  bool finished = false;
  double host_wait_s = 0.0, host_enqueue_s = 0.0;   // THY_SLQ_TIMING: host seconds waiting for round r-2 / enqueuing rounds
  thy_slq_stats stats{};
  ~thy_slq_s() {
    if (graph) cudaGraphExecDestroy(graph);
    if (side) cudaStreamDestroy(side);
    if (side2) cudaStreamDestroy(side2);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (ev_fork2) cudaEventDestroy(ev_fork2);
    if (ev_sigma) cudaEventDestroy(ev_sigma);
    if (ev_rng) cudaEventDestroy(ev_rng);
    for (auto& e : ev_round) if (e) cudaEventDestroy(e);
    for (auto& e : tev) if (e) cudaEventDestroy(e);
    if (host_ao) cudaFreeHost(host_ao);
    if (host_c) cudaFreeHost(host_c);
    if (host_conv) cudaFreeHost(host_conv);
    for (void* q : p2p_opened) cudaIpcCloseMemHandle(q);
  }
};

namespace thy {

// [0..4]: consecutive phases of the main stream (they add up to the round); [5]: the pool-spread kernels, which run on
// their own stream beside phases 0-1; [6]: unused
static const char* const kPhaseNames[7] = {"all-gather of keys", "radix select + flags + retired values", "walk + adapt",
                                           "wait for quadrature", "round control", "pool spread (side stream)", "-"};

__global__ void k_iota(int n, int* v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

// retired values recorded in global order at ret_ll[iterations ...] (iterations read from the device-side control
// word so that the launch can live inside a CUDA graph)
__global__ void __launch_bounds__(256) k_slq_copy_retired(const double* __restrict__ sorted, int Kr,
                                                          const long long* __restrict__ ctl, double* __restrict__ ret_ll) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Kr) ret_ll[ctl[0] + i] = sorted[i];
}

// Convergence test of the reference (SlqEngine.scala:249-262: iterations > minIterationCount && iterations >= 10 P H) on
// the device: the first round at which it holds is written ONCE to ctl[2].
__global__ void k_slq_converge(long long* __restrict__ ctl, const double* __restrict__ abs_out, int A, int Kr, double P,
                               long long min_iterations) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double hmax = -INFINITY;
  for (int a = 0; a < A; ++a) hmax = fmax(hmax, abs_out[a * 2 + 1]);
  const long long it = ctl[0] + Kr;
  if (ctl[2] < 0 && it > min_iterations && isfinite(hmax) && (double)it >= 10.0 * P * hmax) ctl[2] = ctl[1];
}
__global__ void k_slq_advance(long long* __restrict__ ctl, int Kr) {
  ctl[0] += Kr;
  ctl[1] += 1;
}
// key of every retired slot (for the index sort of keepRetiredPoints)
__global__ void __launch_bounds__(256) k_slq_slot_keys(int c, const int* __restrict__ sidx, const double* __restrict__ ll,
                                                       double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c) out[i] = ll[sidx[i]];
}

// Per-dimension standard deviation of the live pool: deterministic two-stage reduction.
// Pl = number of SAMPLED rows; sampled row r is pool row r * row_stride (a strided subsample of at most 65 536 points
// estimates the spread to well under a percent; it only sets the proposal scale, which any positive value keeps valid).
__global__ void __launch_bounds__(256) k_slq_sigma_part(long long Pl, long long row_stride, int d, const double* __restrict__ x,
                                                        double* __restrict__ part, int nslab) {
  // block = one slab of rows, read as one contiguous stream (fully coalesced); thread t always sees the same column
  // phase because the stride 256 * k is re-based per row group: element e of the slab has column (e0 + e) % d
  __shared__ double ss[256], ss2[256];
  const int slab = blockIdx.x;
  const long long r0 = (Pl * slab) / nslab, r1 = (Pl * (slab + 1)) / nslab;
  // rows handled in groups of RG rows so that RG * d is a multiple of 256-thread passes with a fixed column per thread
  // (simple scheme: thread t owns column t % d of rows t / d, t / d + 256 / d, ... ; threads >= (256/d)*d idle)
  const int rows_per_pass = 256 / d > 0 ? 256 / d : 1;
  double s = 0.0, s2 = 0.0;
  if (d <= 256) {
    const int rr = threadIdx.x / d, j = threadIdx.x % d;
    if (rr < rows_per_pass) {
      for (long long r = r0 + rr; r < r1; r += rows_per_pass) {
        const double v = x[r * row_stride * d + j];
        s += v;
        s2 = fma(v, v, s2);
      }
    }
    ss[threadIdx.x] = (rr < rows_per_pass) ? s : 0.0;
    ss2[threadIdx.x] = (rr < rows_per_pass) ? s2 : 0.0;
    __syncthreads();
    if (threadIdx.x < d) {
      double a = 0.0, a2 = 0.0;
      for (int q = 0; q < rows_per_pass; ++q) {   // fixed order: deterministic
        a += ss[q * d + threadIdx.x];
        a2 += ss2[q * d + threadIdx.x];
      }
      part[((size_t)slab * 2 + 0) * d + threadIdx.x] = a;
      part[((size_t)slab * 2 + 1) * d + threadIdx.x] = a2;
    }
  } else {
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
      s = 0.0;
      s2 = 0.0;
      for (long long r = r0; r < r1; ++r) {
        const double v = x[r * row_stride * d + j];
        s += v;
        s2 = fma(v, v, s2);
      }
      part[((size_t)slab * 2 + 0) * d + j] = s;
      part[((size_t)slab * 2 + 1) * d + j] = s2;
    }
  }
}
__global__ void __launch_bounds__(256) k_slq_sigma_final(long long Pl, int d, const double* __restrict__ part, int nslab,
                                                         double* __restrict__ sigma) {
  // one warp per dimension; lanes stride the slabs in a fixed order, then a shuffle tree: deterministic
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= d) return;
  double s = 0.0, s2 = 0.0;
  for (int b = lane; b < nslab; b += 32) {
    s += part[((size_t)b * 2 + 0) * d + j];
    s2 += part[((size_t)b * 2 + 1) * d + j];
  }
  s = warp_sum(s);
  s2 = warp_sum(s2);
  if (lane != 0) return;
  const double mean = s / Pl;
  const double var = fmax(s2 / Pl - mean * mean, 0.0);
  sigma[j] = sqrt(var);
}

// Walkers start from uniformly chosen survivors (slq_pick_survivor: the same draw as the fused walk kernel).
__global__ void __launch_bounds__(256) k_slq_spawn(int c, long long Pl, int d, uint64_t seed, uint64_t round, int rank,
                                                   const int* __restrict__ sidx, const unsigned char* __restrict__ flags,
                                                   const double* __restrict__ x, const double* __restrict__ ll,
                                                   const double* __restrict__ lpr, double* __restrict__ Y,
                                                   double* __restrict__ yl, double* __restrict__ ypr, int* __restrict__ slot) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= c) return;
  const long long src = slq_pick_survivor(seed, ((uint64_t)rank << 40) | (uint64_t)w, round, Pl, flags);
  for (int j = lane; j < d; j += 32) Y[(size_t)w * d + j] = x[(size_t)src * d + j];
  if (lane == 0) {
    yl[w] = ll[src];
    ypr[w] = lpr[src];
    slot[w] = sidx[w];
  }
}

__global__ void __launch_bounds__(256) k_slq_propose(int c, int d, uint64_t seed, uint64_t step, int rank,
                                                     const double* __restrict__ Y, const double* __restrict__ sigma,
                                                     const double* __restrict__ scalars, double* __restrict__ Yp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)c * d) return;
  const int w = (int)(i / d), j = (int)(i % d);
  double z0, z1;
  double z[4];
  rng_normal_quad_fast(seed, ((uint64_t)rank << 40) | (uint64_t)w, slq_move_item(j), RNG_SLQ_MOVE, step, z);
  const int br = slq_move_branch(j);
  Yp[i] = Y[i] + scalars[0] * sigma[j] * (br == 0 ? z[0] : br == 1 ? z[1] : br == 2 ? z[2] : z[3]);
}

// Metropolis test w.r.t. the prior, hard likelihood constraint L' > L*.
__global__ void __launch_bounds__(256) k_slq_accept(int c, int d, uint64_t seed, uint64_t step, int rank,
                                                    double* __restrict__ Y, double* __restrict__ yl, double* __restrict__ ypr,
                                                    const double* __restrict__ Yp, const double* __restrict__ tl,
                                                    const double* __restrict__ tpr, const double* __restrict__ scalars,
                                                    int* __restrict__ accepted) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= c) return;
  const double lprior_new = tpr[w];
  const double lnew = tl[w];
  const RngDraw u = rng_uniform2(seed, ((uint64_t)rank << 40) | (uint64_t)w, 0, RNG_SLQ_ACCEPT, step);
  const bool ok = isfinite(lprior_new) && isfinite(lnew) && (lnew > scalars[1]) &&
                  (lprior_new - ypr[w] >= log(1.0 - u.u0));
  if (ok) {
    for (int j = lane; j < d; j += 32) Y[(size_t)w * d + j] = Yp[(size_t)w * d + j];
  }
  if (lane == 0) {
    if (ok) {
      yl[w] = lnew;
      ypr[w] = lprior_new;
      accepted[w] += 1;   // per-walker tally, summed once per round by k_slq_adapt (no atomics)
    }
  }
}

__global__ void __launch_bounds__(256) k_slq_commit(int c, int d, const int* __restrict__ slot, const double* __restrict__ Y,
                                                    const double* __restrict__ yl, const double* __restrict__ ypr,
                                                    double* __restrict__ x, double* __restrict__ ll, double* __restrict__ lpr) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= c) return;
  const int s = slot[w];
  for (int j = lane; j < d; j += 32) x[(size_t)s * d + j] = Y[(size_t)w * d + j];
  if (lane == 0) {
    ll[s] = yl[w];
    lpr[s] = ypr[w];
  }
}

// scale *= (1 + inc) when the round's acceptance beat the target, else /= (1 + inc)   (SlqConfig.domainScalingIncrement,
// targetAcceptanceProbability play the same role for the reference's cube scaling, QuadratureDomainTelemetry.scala)
// counter != 0: the fused walk kernel has already summed the round's acceptances into accepted[0] (one integer atomic per
// warp), so one warp suffices; otherwise accepted[w] holds per-walker tallies of the per-step kernels.
__global__ void __launch_bounds__(1024) k_slq_adapt(double* scalars, double inc, double target, const int* __restrict__ c_dev,
                                                    int steps, int* __restrict__ accepted, int counter) {
  __shared__ double sh[32];
  const int c = *c_dev;
  double a = 0.0;
  if (counter) {
    if (threadIdx.x == 0) {
      a = (double)accepted[0];
      accepted[0] = 0;
    }
  } else {
    for (int w = threadIdx.x; w < c; w += blockDim.x) {
      a += (double)accepted[w];
      accepted[w] = 0;
    }
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x != 0) return;
  double acc = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) acc += sh[i];
  const double prop = (double)c * (double)steps;
  if (prop > 0.0) {
    const double rate = acc / prop;
    scalars[4] = rate;
    if (rate > target) scalars[0] *= (1.0 + inc); else scalars[0] /= (1.0 + inc);
    scalars[0] = fmin(fmax(scalars[0], 1e-6), 10.0);
  }
  scalars[5] += acc;
  scalars[6] += prop;
  scalars[2] = fmax(scalars[2], (double)c);   // load balance across ranks: largest / total local retire count so far
  scalars[3] += (double)c;
}

// Copy retired points in the order of sidx (world == 1 with keepRetiredPoints: slots sorted by log-likelihood, i.e. the
// global retirement order).
__global__ void __launch_bounds__(256) k_slq_store(int c, int d, const int* __restrict__ sidx, const double* __restrict__ x,
                                                   double* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= c) return;
  const int s = sidx[w];
  for (int j = lane; j < d; j += 32) dst[(size_t)w * d + j] = x[(size_t)s * d + j];
}

static SelectBuffers sel_buffers(thy_slq_s* h) {
  SelectBuffers b{};
  b.state = h->sel_state.ptr;
  b.hist = h->sel_hist.ptr;
  b.ticket = h->sel_ticket.ptr;
  b.cursor = h->sel_cursor.ptr;
  b.cnt = h->sel_cnt.ptr;
  b.flags = h->flags.ptr;
  b.sidx = h->sidx.ptr;
  b.count = h->count_dev.ptr;
  b.scalars = h->scalars.ptr;
  return b;
}
static EvidenceBuffers ev_buffers(thy_slq_s* h) {
  EvidenceBuffers e{};
  e.lt = h->ev_lt.ptr;
  e.segsum = h->ev_segsum.ptr;
  e.part = h->ev_part.ptr;
  e.tickets = h->ev_tickets.ptr;
  e.state = h->abs_state.ptr;
  e.out = h->abs_out.ptr;
  e.stride = std::max<long long>(h->K, h->P);
  e.nseg_max = (int)((e.stride + 2047) / 2048);
  return e;
}

// Ascending sort of n values: a round's retired log-likelihoods go through the bucket sort of slq_sort.cu (no library
// call in a round); the whole pool, once, at the drain, through cub's radix sort.
static thy_status slq_sort_values(thy_slq_s* h, const double* in, double* out, int n, cudaStream_t s) {
  if (n <= h->K && h->bs_tmp.ptr && !h->cub_round_sort) {
    const BucketSortBuffers b{h->bs_minmax.ptr, h->bs_count.ptr, h->bs_cursor.ptr, h->bs_offset.ptr, h->bs_tmp.ptr};
    return slq_bucket_sort(b, in, out, n, s);
  }
  size_t bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, bytes, in, out, n, 0, 64, s);
  if (bytes > h->cub_tmp.count) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cs);
    if (cs != cudaStreamCaptureStatusNone) return fail(THY_ERR_STATE, "SLQ: sort scratch must be sized before graph capture");
    THY_CUDA_CHECK(cudaStreamSynchronize(s));
    THY_TRY(h->cub_tmp.reserve(bytes));
  }
  THY_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(h->cub_tmp.ptr, bytes, in, out, n, 0, 64, s));
  count_launch(3);
  return THY_OK;
}
static thy_status slq_sort_pairs(thy_slq_s* h, const double* kin, double* kout, const int* vin, int* vout, int n,
                                 cudaStream_t s) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, 64, s);
  if (bytes > h->cub_tmp2.count) {
    THY_CUDA_CHECK(cudaStreamSynchronize(s));
    THY_CUDA_CHECK(cudaStreamSynchronize(h->side));
    THY_TRY(h->cub_tmp2.reserve(bytes));
  }
  THY_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(h->cub_tmp2.ptr, bytes, kin, kout, vin, vout, n, 0, 64, s));
  count_launch(3);
  return THY_OK;
}

// log-prior and log-likelihood of a batch, each evaluated once: the prior kernel writes `prior`, the likelihood terms
// are accumulated into a zeroed `loglik` (no total - prior subtraction, so no cancellation either)
static thy_status slq_eval(thy_slq_s* h, long long n, const double* X, double* loglik, double* prior) {
  thy_model_s* m = h->model;
  THY_TRY(launch_prior(m->prior_view, n, X, prior, nullptr, -1.0, m->stream));
  THY_CUDA_CHECK(cudaMemsetAsync(loglik, 0, (size_t)n * sizeof(double), m->stream));
  THY_TRY(eval_batch_dev(m, n, X, loglik, nullptr, true, true));
  return THY_OK;
}

// One full round, enqueued on s (and the side stream, forked and joined): graph-capturable when the fused walk covers
// the model.  tracing: CUDA events between the phases.
static thy_status slq_round_enqueue(thy_slq_s* h, int Kr, cudaStream_t s, bool tracing) {
  thy_model_s* m = h->model;
  const int d = h->d;
  const SelectBuffers sb = sel_buffers(h);
  auto mark = [&](int i) { if (tracing) cudaEventRecord(h->tev[i], s); };
  mark(0);
  // ---- proposal spread of the live pool (strided subsample), beside the all-gather and the selection ----
  THY_CUDA_CHECK(cudaEventRecord(h->ev_fork2, s));
  THY_CUDA_CHECK(cudaStreamWaitEvent(h->side2, h->ev_fork2, 0));
  if (tracing) cudaEventRecord(h->tev[6], h->side2);
  const long long nrows = std::min<long long>(h->Pl, 65536), row_stride = h->Pl / nrows;
  const int nslab = (int)std::min<long long>(296, nrows);
  k_slq_sigma_part<<<nslab, 256, 0, h->side2>>>(nrows, row_stride, d, h->x.ptr, h->sig_part.ptr, nslab);
  k_slq_sigma_final<<<(unsigned)ceil_div(d, 8), 256, 0, h->side2>>>(nrows, d, h->sig_part.ptr, nslab, h->sigma.ptr);
  count_launch(2);
  if (tracing) cudaEventRecord(h->tev[7], h->side2);
  THY_CUDA_CHECK(cudaEventRecord(h->ev_sigma, h->side2));
  // the walk's randomness for this round, drawn while the main stream gathers and selects (sharded pools only), on the
  // quadrature stream, which is idle until the selection has finished: beside the pool spread, not behind it
  if (h->fused_walk && h->pre_groups > 0) {
    THY_CUDA_CHECK(cudaStreamWaitEvent(h->side, h->ev_fork2, 0));
    THY_TRY(launch_slq_walk_rng(m, h->pre_groups, h->M, h->seed, h->ctl.ptr + 1, h->rank, h->zpre.ptr, h->lacc.ptr, h->side));
    THY_CUDA_CHECK(cudaEventRecord(h->ev_rng, h->side));
  }
  // ---- every rank sees every log-likelihood: ONE collective per round ----
  const double* keys = h->ll.ptr;
  if (h->world > 1) {
    if (h->p2p) {
      // no collective: the peers' walks staged their new keys in this rank's memory over NVLink; wait for their flags
      // and fold the pairs into the local copy of all keys
      THY_TRY(slq_p2p_wait_and_apply(h->p2p_view, h->p2p_buf.ptr, h->ctl.ptr, h->scalars.ptr, h->keys_all.ptr, s));
    } else {
      THY_TRY(comm_all_gather_f64(h->comm, h->ll.ptr, h->keys_all.ptr, (size_t)h->Pl, s));
    }
    keys = h->keys_all.ptr;
  }
  mark(1);
  // ---- threshold = K-th smallest; this rank's retirees; the retired values (before the walk overwrites their slots) ----
  THY_TRY(slq_select_launch(sb, keys, h->Pl, h->world, h->rank, Kr, h->vals_raw.ptr, s));
  mark(2);
  // ---- quadrature update beside the walk ----
  THY_CUDA_CHECK(cudaEventRecord(h->ev_fork, s));
  THY_CUDA_CHECK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
  THY_TRY(slq_sort_values(h, h->vals_raw.ptr, h->vals_sorted.ptr, Kr, h->side));
  k_slq_copy_retired<<<(unsigned)ceil_div(Kr, 256), 256, 0, h->side>>>(h->vals_sorted.ptr, Kr, h->ctl.ptr, h->ret_ll.ptr);
  THY_TRY(slq_evidence_launch(ev_buffers(h), h->A, Kr, h->P, h->K, LLONG_MAX, 0, h->ctl.ptr, h->seed, h->vals_sorted.ptr, 0,
                              h->side));
  k_slq_converge<<<1, 32, 0, h->side>>>(h->ctl.ptr, h->abs_out.ptr, h->A, Kr, (double)h->P, (long long)h->cfg.minIterationCount);
  count_launch(2);
  THY_CUDA_CHECK(cudaMemcpyAsync(h->host_ao, h->abs_out.ptr, (size_t)h->A * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->side));
  THY_CUDA_CHECK(cudaMemcpyAsync(h->host_conv, h->ctl.ptr + 2, sizeof(long long), cudaMemcpyDeviceToHost, h->side));
  THY_CUDA_CHECK(cudaEventRecord(h->ev_join, h->side));
  // ---- retired positions in retirement order (single GPU, keepRetiredPoints): all Kr retirees are local ----
  if (h->cfg.keepRetiredPoints && h->world == 1) {
    k_slq_slot_keys<<<(unsigned)ceil_div(Kr, 256), 256, 0, s>>>(Kr, h->sidx.ptr, h->ll.ptr, h->Yp.ptr);
    THY_TRY(slq_sort_pairs(h, h->Yp.ptr, h->tl.ptr, h->sidx.ptr, h->pair_out.ptr, Kr, s));
    k_slq_store<<<(unsigned)ceil_div(Kr, 8), 256, 0, s>>>(Kr, d, h->pair_out.ptr, h->x.ptr, h->ret_x.ptr + (size_t)h->iterations * d);
    count_launch(2);
  }
  // ---- replacement: constrained walks from survivors ----
  THY_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_sigma, 0));
  if (h->fused_walk && h->pre_groups > 0) THY_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_rng, 0));
  if (h->fused_walk) {
    THY_TRY(launch_slq_walk(m, Kr, h->count_dev.ptr, h->Pl, h->M, h->seed, h->ctl.ptr + 1, h->rank, h->sidx.ptr, h->flags.ptr,
                            h->x.ptr, h->ll.ptr, h->lpr.ptr, h->sigma.ptr, h->scalars.ptr, h->accepted.ptr, s,
                            h->pre_groups > 0 ? h->zpre.ptr : nullptr, h->lacc.ptr, h->pre_groups,
                            (int)std::min<long long>(Kr, ((long long)Kr + h->world - 1) / h->world)));
  } else {
    int c = Kr;   // single GPU: every retiree is local
    if (h->world > 1) {
      THY_CUDA_CHECK(cudaMemcpyAsync(h->host_c, h->count_dev.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
      THY_CUDA_CHECK(cudaStreamSynchronize(s));
      c = *h->host_c;
    }
    if (c > h->K) return fail(THY_ERR_STATE, "SLQ: internal error, retire count exceeds the walker buffers");
    if (c > 0) {
      k_slq_spawn<<<(unsigned)ceil_div(c, 8), 256, 0, s>>>(c, h->Pl, d, h->seed, (uint64_t)h->rounds, h->rank, h->sidx.ptr,
                                                          h->flags.ptr, h->x.ptr, h->ll.ptr, h->lpr.ptr, h->Y.ptr, h->yl.ptr,
                                                          h->ypr.ptr, h->slot.ptr);
      count_launch();
      for (int step = 0; step < h->M; ++step) {
        const uint64_t gstep = (uint64_t)h->rounds * (uint64_t)h->M + step;
        k_slq_propose<<<(unsigned)ceil_div((long long)c * d, 256), 256, 0, s>>>(c, d, h->seed, gstep, h->rank, h->Y.ptr,
                                                                                 h->sigma.ptr, h->scalars.ptr, h->Yp.ptr);
        THY_TRY(slq_eval(h, c, h->Yp.ptr, h->tl.ptr, h->tpr.ptr));
        k_slq_accept<<<(unsigned)ceil_div(c, 8), 256, 0, s>>>(c, d, h->seed, gstep, h->rank, h->Y.ptr, h->yl.ptr, h->ypr.ptr,
                                                             h->Yp.ptr, h->tl.ptr, h->tpr.ptr, h->scalars.ptr, h->accepted.ptr);
        count_launch(2);
      }
      k_slq_commit<<<(unsigned)ceil_div(c, 8), 256, 0, s>>>(c, d, h->slot.ptr, h->Y.ptr, h->yl.ptr, h->ypr.ptr, h->x.ptr,
                                                           h->ll.ptr, h->lpr.ptr);
      count_launch();
    }
  }
  k_slq_adapt<<<1, h->fused_walk ? 32 : 1024, 0, s>>>(h->scalars.ptr, h->cfg.domainScalingIncrement,
                                                      h->cfg.targetAcceptanceProbability, h->count_dev.ptr, h->M,
                                                      h->accepted.ptr, h->fused_walk ? 1 : 0);
  count_launch();
  // this rank's new keys go to every rank's staging area over NVLink; then the flag says so
  if (h->p2p) THY_TRY(slq_p2p_signal(h->p2p_view, h->Pl, h->sidx.ptr, h->ll.ptr, h->count_dev.ptr, h->ctl.ptr, s));
  mark(3);
  THY_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_join, 0));
  mark(4);
  k_slq_advance<<<1, 1, 0, s>>>(h->ctl.ptr, Kr);
  count_launch();
  mark(5);
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

static thy_status slq_round(thy_slq_s* h, int Kr) {
  cudaStream_t s = h->model->stream;
  if (h->iterations + Kr > h->ret_capacity) return fail(THY_ERR_STATE, "SLQ: retired buffer exhausted");
  const bool graphable = h->use_graph && !h->trace && s != 0 && Kr == h->K && h->fused_walk && h->sort_sized &&
                         !(h->cfg.keepRetiredPoints && h->world == 1);
  if (h->graph && h->graph_stream != s) {   // the model was moved to another stream: re-capture
    cudaGraphExecDestroy(h->graph);
    h->graph = nullptr;
  }
  if (graphable && !h->graph) {
    cudaGraph_t g = nullptr;
    const uint64_t before = g_kernel_launches.load();
    THY_CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const thy_status st = slq_round_enqueue(h, Kr, s, false);
    const cudaError_t ce = cudaStreamEndCapture(s, &g);
    h->graph_launches = g_kernel_launches.load() - before;
    g_kernel_launches.store(before);   // captured, not executed
    if (st != THY_OK || ce != cudaSuccess || !g) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      h->use_graph = false;   // capture unsupported here (e.g. a collective library without graph support): plain launches
    } else {
      const cudaError_t ci = cudaGraphInstantiate(&h->graph, g, 0);
      cudaGraphDestroy(g);
      if (ci != cudaSuccess) {
        cudaGetLastError();
        h->graph = nullptr;
        h->use_graph = false;
      } else {
        h->graph_stream = s;
      }
    }
  }
  if (graphable && h->graph) {
    THY_CUDA_CHECK(cudaGraphLaunch(h->graph, s));
    count_launch(h->graph_launches);
  } else {
    THY_TRY(slq_round_enqueue(h, Kr, s, h->trace));
    h->sort_sized = true;
  }
  THY_CUDA_CHECK(cudaEventRecord(h->ev_round[h->rounds % 3], s));
  if (h->trace) {
    THY_CUDA_CHECK(cudaEventSynchronize(h->tev[5]));
    for (int i = 0; i < 5; ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, h->tev[i], h->tev[i + 1]) == cudaSuccess) h->tsum[i] += ms;
    }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->tev[6], h->tev[7]) == cudaSuccess) h->tsum[5] += ms;
    h->tcount += 1;
  }
  h->iterations += Kr;
  h->rounds += 1;
  h->evals += (long long)Kr * h->M;
  return THY_OK;
}

// Live phase: rounds of "retire the K lowest, replace them above the threshold" until the reference's convergence test
// (SlqEngine.scala:249-262) or maxIterationCount.  The host stays at most two rounds ahead of the device: before it
// enqueues round r it waits for round r - 2 and reads the device's write-once "first converged round" word.  Stopping
// once that word is <= r - 2 makes every rank run exactly (first converged round + 2) rounds, whatever the host timing.
static thy_status slq_live_phase(thy_slq_s* h, long long max_rounds, bool* converged) {
  *converged = false;
  for (long long r = 0; r < max_rounds; ++r) {
    const long long remaining = (long long)h->cfg.maxIterationCount - h->iterations;
    if (remaining <= 0) {
      *converged = true;
      break;
    }
    const auto hw0 = std::chrono::steady_clock::now();
    if (h->rounds >= 2) {
      THY_CUDA_CHECK(cudaEventSynchronize(h->ev_round[(h->rounds - 2) % 3]));
      const long long fc = *(volatile long long*)h->host_conv;
      if (fc >= 0 && fc <= h->rounds - 2) {
        *converged = true;
        break;
      }
    }
    const auto hw1 = std::chrono::steady_clock::now();
    THY_TRY(slq_round(h, (int)std::min<long long>(h->K, remaining)));
    h->host_wait_s += std::chrono::duration<double>(hw1 - hw0).count();
    h->host_enqueue_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - hw1).count();
  }
  return THY_OK;
}

__global__ void __launch_bounds__(256) k_slq_gather_rows(int n, int d, const long long* __restrict__ idx,
                                                         const double* __restrict__ src, double* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n) return;
  const long long s = idx[w];
  for (int j = lane; j < d; j += 32) dst[(size_t)w * d + j] = src[(size_t)s * d + j];
}

// ln X of abscissa a for the retired points [0, n) into host memory (chunked through the quadrature scratch).
static thy_status slq_export_logx(thy_slq_s* h, int a, long long n, double* out_host) {
  cudaStream_t s = h->model->stream;
  const EvidenceBuffers e = ev_buffers(h);
  const long long live = h->live_iterations >= 0 ? h->live_iterations : LLONG_MAX;
  const int chunk = (int)std::min<long long>(e.stride, 1 << 20);
  DevBuf<double> carry, buf;
  THY_TRY(carry.reserve(1));
  THY_TRY(buf.reserve((size_t)chunk));
  THY_CUDA_CHECK(cudaMemsetAsync(carry.ptr, 0, sizeof(double), s));
  for (long long i0 = 0; i0 < n; i0 += chunk) {
    const int cnt = (int)std::min<long long>(chunk, n - i0);
    // retirements per round of the prior-mass simulation: K in the default mode, 1 with the reference's proposals
    THY_TRY(slq_logx_launch(e, a, i0, cnt, h->P, h->cfg.referenceProposals ? 1 : h->K, live, h->seed, carry.ptr, buf.ptr, s));
    THY_CUDA_CHECK(cudaMemcpyAsync(out_host + i0, buf.ptr, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, s));
    THY_CUDA_CHECK(cudaStreamSynchronize(s));
  }
  return THY_OK;
}


// ---- reference-faithful proposals (thy_slq_config.referenceProposals): the cubical cover -------------------------------
// Reference: integration/slq/PointInInterval.scala:40-109, PointInCube.scala:30-117, PointInCubeCollection.scala:60-111,
// QuadratureDomainTelemetry.scala:27-110 (rescaling disabled, its default), SlqEngine.scala:120-197,216-247,309-345.
// The reference proposes ONE point at a time from a cover of the live pool by pairwise-disjoint, point-centred cubes
// (cube chosen with probability proportional to the SUM of its side lengths, PointInCube.scala:30-33, sic), replaces the
// pool minimum when the proposal's logPdf is above it, and rebuilds the cover from the pool after 1000 rejections in a row
// (given at least one acceptance since the last rebuild).  The make-disjoint fold is a strictly sequential O(P^2) chain and
// the pool update is one point at a time, so this mode is for SMALL pools: the geometry and the pool bookkeeping run on
// the host in the library, every logPdf -- the hot path -- is evaluated on the device in batches of sampleParallelism
// (the reference's fibres in flight: proposals of a batch come from the same cover, and what is left of a batch when the
// cover is rebuilt is discarded), and the quadrature runs on the device as in the default mode, with one retirement per
// step.  Ranking key = the full posterior logPdf, as in the reference (its measure quirk Q6 included).  Every random
// number is a counter-based draw (stream, proposal index), so oracle/slq_cubical_cover.py replays the run draw for draw.
namespace {

struct CubicalCover {
  int P = 0, d = 0;
  std::vector<double> pt, lo, up, vol;   // [P][d] points / bounds; [P] "volumes"
  void build() {
    lo.assign((size_t)P * d, 0.0);
    up.assign((size_t)P * d, 0.0);
    // initialise (PointInCubeCollection.scala:60-85): half-width = largest distance to any other point, per dimension
    for (int i = 0; i < P; ++i)
      for (int k = 0; k < d; ++k) {
        double h = 0.0;
        for (int j = 0; j < P; ++j) {
          const double a = std::fabs(pt[(size_t)j * d + k] - pt[(size_t)i * d + k]);
          if (a > h) h = a;
        }
        lo[(size_t)i * d + k] = pt[(size_t)i * d + k] - h;
        up[(size_t)i * d + k] = pt[(size_t)i * d + k] + h;
      }
    // makeDisjoint (PointInCube.scala:84-115): the fold, accumulator order included
    std::vector<int> acc, work;
    for (int c = 0; c < P; ++c) {
      work.assign(1, c);
      for (size_t a = 0; a < acc.size(); ++a) {
        const int cur = acc[a], head = work[0];
        bool inter = true;
        for (int k = 0; k < d && inter; ++k) {
          const double l1 = lo[(size_t)head * d + k], u1 = up[(size_t)head * d + k];
          const double l2 = lo[(size_t)cur * d + k], u2 = up[(size_t)cur * d + k];
          inter = ((u1 > l2) && (l1 < u2)) || ((u2 > l1) && (l2 < u1));
        }
        if (inter) {
          int kbest = 0;
          double best = -1.0;
          for (int k = 0; k < d; ++k) {   // dimensionOfLargestSeparation: first maximum
            const double df = pt[(size_t)head * d + k] - pt[(size_t)cur * d + k];
            if (df * df > best) { best = df * df; kbest = k; }
          }
          const size_t ih = (size_t)head * d + kbest, ic = (size_t)cur * d + kbest;
          const double p1 = pt[ih], p2 = pt[ic];
          const bool one_larger = p1 > p2;
          const double larger_lo = one_larger ? lo[ih] : lo[ic], smaller_up = one_larger ? up[ic] : up[ih];
          const double avg = (p1 + p2) / 2.0;
          double boundary;
          if (larger_lo <= avg && smaller_up > avg) boundary = avg;
          else if (larger_lo <= avg) boundary = smaller_up;
          else boundary = larger_lo;   // (the reference throws when neither holds; unreachable for intersecting cubes)
          if (one_larger) { lo[ih] = boundary; up[ic] = boundary; } else { up[ih] = boundary; lo[ic] = boundary; }
        }
        work.insert(work.begin() + 1, cur);   // Vector(new, current) ++ tail
      }
      acc = work;
    }
    // symmetrize (PointInInterval.scala:51-61) and volumes (PointInCube.scala:30-33: the SUM of the side lengths)
    vol.assign(P, 0.0);
    for (int i = 0; i < P; ++i) {
      double v = 0.0;
      for (int k = 0; k < d; ++k) {
        const size_t q = (size_t)i * d + k;
        const double l1 = up[q] - pt[q], l2 = pt[q] - lo[q];
        if (l1 > l2) up[q] = pt[q] + l2; else if (l1 < l2) lo[q] = pt[q] - l1;
        v += std::fabs(up[q] - lo[q]);
      }
      vol[i] = v;
    }
  }
  int choose(double u, std::vector<double>& cdf) const {   // MathOps.cdfStaircase + half-open lookup
    cdf.resize(P);
    double run = 0.0;
    for (int i = 0; i < P; ++i) { run += vol[i]; cdf[i] = run; }
    for (int i = 0; i < P; ++i)
      if (cdf[i] / run > u) return i;
    return P - 1;
  }
};

constexpr uint32_t RNG_SLQ_REFERENCE = 12;
inline double ref_uniform(uint64_t seed, int stream, long long n) {
  return rng_uniform2(seed, (uint64_t)stream, 0, RNG_SLQ_REFERENCE, (uint64_t)n).u0;
}

}  // namespace

static thy_status slq_run_reference(thy_slq_s* h) {
  thy_model_s* m = h->model;
  cudaStream_t s = m->stream;
  const int P = (int)h->P, d = h->d, B = h->K;
  const long long max_it = h->cfg.maxIterationCount;
  const int rebuild_streak = 1000, exhausted_streak = 100000;   // QuadratureDomainTelemetry.scala:43,48-49
  CubicalCover cov;
  cov.P = P;
  cov.d = d;
  cov.pt.resize((size_t)P * d);
  std::vector<double> key(P);
  // ranking key = posterior logPdf of the pool (prior + likelihood), evaluated on the device
  THY_TRY(eval_batch_dev(m, P, h->x.ptr, h->ll.ptr, nullptr, true));
  THY_CUDA_CHECK(cudaMemcpyAsync(cov.pt.data(), h->x.ptr, (size_t)P * d * sizeof(double), cudaMemcpyDeviceToHost, s));
  THY_CUDA_CHECK(cudaMemcpyAsync(key.data(), h->ll.ptr, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, s));
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  cov.build();
  DevBuf<double> Xp, Lp;
  THY_TRY(Xp.reserve((size_t)B * d));
  THY_TRY(Lp.reserve((size_t)B));
  std::vector<double> props((size_t)B * d), vals(B), cdf, retired, retired_x;
  long long n = 0, processed = 0;
  int acceptances_since_rebuild = 0, streak = 0;
  long long acceptances = 0, rejections = 0, rebuilds = 0;
  const EvidenceBuffers ev = ev_buffers(h);
  const int chunk = (int)std::min<long long>(ev.stride, P);
  bool converged = false;
  // quadrature on the device, one retirement per step (K = 1): n_live = P throughout the live phase
  auto flush_evidence = [&](long long upto) -> thy_status {
    while (upto - processed >= chunk || (upto == (long long)retired.size() && upto > processed && converged)) {
      const int cnt = (int)std::min<long long>(chunk, upto - processed);
      if (cnt <= 0) break;
      if (processed + cnt > h->ret_capacity) return fail(THY_ERR_STATE, "SLQ: retired buffer exhausted");
      THY_CUDA_CHECK(cudaMemcpyAsync(h->ret_ll.ptr + processed, retired.data() + processed, (size_t)cnt * sizeof(double),
                                     cudaMemcpyHostToDevice, s));
      THY_TRY(slq_evidence_launch(ev, h->A, cnt, h->P, 1, LLONG_MAX, processed, nullptr, h->seed, h->ret_ll.ptr + processed, 0, s));
      THY_CUDA_CHECK(cudaMemcpyAsync(h->host_ao, h->abs_out.ptr, (size_t)h->A * 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
      THY_CUDA_CHECK(cudaStreamSynchronize(s));
      processed += cnt;
    }
    return THY_OK;
  };
  while ((long long)retired.size() < max_it && streak < exhausted_streak && !converged) {
    for (int b = 0; b < B; ++b) {
      const int i = cov.choose(ref_uniform(h->seed, 0, n + b), cdf);
      for (int k = 0; k < d; ++k) {
        const size_t q = (size_t)i * d + k;
        // PointInInterval.getSample (:76-77) at scale 1
        props[(size_t)b * d + k] = (ref_uniform(h->seed, 1 + k, n + b) - 0.5) * 1.0 * (cov.up[q] - cov.lo[q]) + cov.pt[q];
      }
    }
    THY_CUDA_CHECK(cudaMemcpyAsync(Xp.ptr, props.data(), (size_t)B * d * sizeof(double), cudaMemcpyHostToDevice, s));
    THY_TRY(eval_batch_dev(m, B, Xp.ptr, Lp.ptr, nullptr, true));
    THY_CUDA_CHECK(cudaMemcpyAsync(vals.data(), Lp.ptr, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, s));
    THY_CUDA_CHECK(cudaStreamSynchronize(s));
    h->evals += B;
    for (int b = 0; b < B; ++b) {
      n += 1;
      if ((long long)retired.size() >= max_it) break;
      int imin = 0;
      for (int j = 1; j < P; ++j)
        if (key[j] < key[imin]) imin = j;
      bool rebuilt = false;
      if (key[imin] < vals[b]) {   // updateStateWithSampleCalculation (:216-233): strictly above the current minimum
        retired.push_back(key[imin]);
        if (h->cfg.keepRetiredPoints) retired_x.insert(retired_x.end(), cov.pt.begin() + (size_t)imin * d, cov.pt.begin() + (size_t)(imin + 1) * d);
        std::copy(props.begin() + (size_t)b * d, props.begin() + (size_t)(b + 1) * d, cov.pt.begin() + (size_t)imin * d);
        key[imin] = vals[b];
        ++acceptances; ++acceptances_since_rebuild; streak = 0;
      } else {
        ++rejections; ++streak;
      }
      if (streak >= rebuild_streak && acceptances_since_rebuild >= 1) {   // initiateRebuild -> rebuildDomain (:309-345)
        cov.build();
        acceptances = rejections = 0;
        acceptances_since_rebuild = 0;   // resetForRebuild keeps the streak
        ++rebuilds;
        rebuilt = true;
      }
      if (rebuilt || streak >= exhausted_streak) {
        n += B - 1 - b;   // the rest of the batch was drawn from the old cover: discarded, its draws skipped
        break;
      }
    }
    // convergence test of the reference (SlqEngine.scala:249-262), here once per P retirements instead of per acceptance
    THY_TRY(flush_evidence((long long)retired.size()));
    if (processed > 0) {
      double hmax = -INFINITY;
      for (int a = 0; a < h->A; ++a) hmax = std::max(hmax, h->host_ao[(size_t)a * 2 + 1]);
      if (processed > h->cfg.minIterationCount && std::isfinite(hmax) && (double)processed >= 10.0 * (double)P * hmax) converged = true;
    }
  }
  converged = true;
  THY_TRY(flush_evidence((long long)retired.size()));
  // drain (SlqEngine.scala:171-180): the remaining pool in ascending order, one at a time
  const long long live = (long long)retired.size();
  std::vector<int> order(P);
  for (int i = 0; i < P; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b2) { return key[a] < key[b2]; });
  std::vector<double> tail(P);
  for (int i = 0; i < P; ++i) tail[i] = key[order[i]];
  if (live + P > h->ret_capacity) return fail(THY_ERR_STATE, "SLQ: retired buffer exhausted");
  THY_CUDA_CHECK(cudaMemcpyAsync(h->ret_ll.ptr + live, tail.data(), (size_t)P * sizeof(double), cudaMemcpyHostToDevice, s));
  THY_TRY(slq_evidence_launch(ev, h->A, P, h->P, 1, live, live, nullptr, h->seed, h->ret_ll.ptr + live, 1, s));
  if (h->cfg.keepRetiredPoints) {
    for (int i = 0; i < P; ++i)
      retired_x.insert(retired_x.end(), cov.pt.begin() + (size_t)order[i] * d, cov.pt.begin() + (size_t)(order[i] + 1) * d);
    THY_CUDA_CHECK(cudaMemcpyAsync(h->ret_x.ptr, retired_x.data(), retired_x.size() * sizeof(double), cudaMemcpyHostToDevice, s));
  }
  // the pool as it ended, back on the device
  THY_CUDA_CHECK(cudaMemcpyAsync(h->x.ptr, cov.pt.data(), (size_t)P * d * sizeof(double), cudaMemcpyHostToDevice, s));
  THY_CUDA_CHECK(cudaMemcpyAsync(h->ll.ptr, key.data(), (size_t)P * sizeof(double), cudaMemcpyHostToDevice, s));
  std::vector<double> ao((size_t)h->A * 2);
  THY_CUDA_CHECK(cudaMemcpyAsync(ao.data(), h->abs_out.ptr, ao.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  h->live_iterations = live;
  h->iterations = live + P;
  h->rounds = n;   // proposals made
  double mean = 0.0, m2 = 0.0, hm = 0.0;
  for (int a = 0; a < h->A; ++a) { mean += ao[(size_t)a * 2]; hm += ao[(size_t)a * 2 + 1]; }
  mean /= h->A;
  hm /= h->A;
  for (int a = 0; a < h->A; ++a) m2 += (ao[(size_t)a * 2] - mean) * (ao[(size_t)a * 2] - mean);
  h->stats.iterations = live;
  h->stats.total_retired = h->iterations;
  h->stats.rounds = n;
  h->stats.log_evidence_mean = mean;
  h->stats.log_evidence_std = h->A > 1 ? std::sqrt(m2 / (h->A - 1)) : 0.0;
  h->stats.neg_entropy_mean = hm;
  h->stats.min_logpdf = tail[0];
  h->stats.walk_scale = 1.0;                       // QuadratureDomainTelemetry.currentScaleFactor (rescaling disabled)
  h->stats.walk_acceptance = n > 0 ? (double)live / (double)n : 0.0;
  h->stats.likelihood_evaluations = h->evals;
  h->ref_rebuilds = rebuilds;
  h->finished = true;
  return THY_OK;
}

}  // namespace thy

using namespace thy;

extern "C" {

thy_status thy_slq_create(thy_model_t m, const thy_slq_config* cfg, uint64_t rng_seed, const double* seeds, int n_seeds,
                          thy_comm_t comm, thy_slq_t* out) {
  thy::ErrScope _err_scope(m);
  THY_TRY(require_device());
  if (!m || !m->finalized) return fail(THY_ERR_STATE, "model is not finalized");
  if (!cfg || !out) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  if (m->lik_dev.empty()) return fail(THY_ERR_INVALID_ARGUMENT, "SLQ needs at least one likelihood");
  const int world = comm ? comm->world : 1;
  if (cfg->poolSize < 4 * world) return fail(THY_ERR_INVALID_ARGUMENT, "poolSize too small");
  if (cfg->poolSize % world != 0) return fail(THY_ERR_INVALID_ARGUMENT, "poolSize must be divisible by the number of ranks");
  if (cfg->abscissaNumber < 1 || cfg->abscissaNumber > 1024) return fail(THY_ERR_INVALID_ARGUMENT, "abscissaNumber out of range");
  if (cfg->maxIterationCount < 1 || cfg->minIterationCount < 0) return fail(THY_ERR_INVALID_ARGUMENT, "bad iteration counts");
  if (!(cfg->domainScalingIncrement > 0) || !(cfg->targetAcceptanceProbability > 0 && cfg->targetAcceptanceProbability < 1))
    return fail(THY_ERR_INVALID_ARGUMENT, "bad scaling increment / target acceptance");
  if (n_seeds < 0 || (n_seeds > 0 && !seeds)) return fail(THY_ERR_INVALID_ARGUMENT, "bad seeds");
  if (cfg->referenceProposals) {
    if (world != 1) return fail(THY_ERR_UNSUPPORTED, "referenceProposals (cubical cover) is a single-GPU, small-pool mode");
    if (cfg->poolSize > 8192) return fail(THY_ERR_UNSUPPORTED, "referenceProposals: the cover's make-disjoint fold is O(P^2) and sequential; poolSize <= 8192");
  }
  auto* h = new (std::nothrow) thy_slq_s();
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "allocation failed");
  h->model = m;
  h->comm = comm;
  h->cfg = *cfg;
  h->seed = rng_seed;
  h->world = world;
  h->rank = comm ? comm->rank : 0;
  h->P = cfg->poolSize;
  h->Pl = h->P / world;
  h->d = m->d;
  h->A = cfg->abscissaNumber;
  h->M = cfg->constrainedWalkSteps > 0 ? cfg->constrainedWalkSteps : 20;
  int K = cfg->sampleParallelism > 0 ? cfg->sampleParallelism : 1;
  // the K globally lowest may all sit on one rank, which must keep survivors to spawn from: K <= half of a rank's pool
  K = (int)std::min<long long>(K, std::max<long long>(1, h->Pl / 2));
  h->K = K;
  h->ret_capacity = (long long)cfg->maxIterationCount + K + h->P;
  const int d = h->d;
  const long long Pl = h->Pl;
  const long long KP = std::max<long long>(K, h->P);
  const int spr = slq_select_slices(Pl);
  const int nseg_max = (int)((KP + 2047) / 2048);
  thy_status st = THY_OK;
  auto chk = [&](thy_status s2) { if (st == THY_OK) st = s2; };
  chk(h->x.reserve((size_t)Pl * d)); chk(h->ll.reserve(Pl)); chk(h->lpr.reserve(Pl));
  if (world > 1) chk(h->keys_all.reserve((size_t)h->P));
  chk(h->sel_state.reserve(1)); chk(h->sel_hist.reserve(6 * 2048)); chk(h->sel_ticket.reserve(1)); chk(h->sel_cursor.reserve(1));
  chk(h->sel_cnt.reserve((size_t)spr * world)); chk(h->flags.reserve(Pl)); chk(h->sidx.reserve(K)); chk(h->count_dev.reserve(1));
  chk(h->vals_raw.reserve((size_t)KP)); chk(h->vals_sorted.reserve((size_t)KP));
  chk(h->pair_in.reserve((size_t)std::max<long long>(K, Pl))); chk(h->pair_out.reserve((size_t)std::max<long long>(K, Pl)));
  chk(h->Y.reserve((size_t)K * d)); chk(h->Yp.reserve((size_t)K * std::max(d, 1))); chk(h->yl.reserve(K)); chk(h->ypr.reserve(K));
  chk(h->tl.reserve(K)); chk(h->tpr.reserve(K)); chk(h->slot.reserve(K)); chk(h->accepted.reserve(K));
  chk(h->bs_minmax.reserve(2)); chk(h->bs_count.reserve(SLQ_SORT_MAX_BUCKETS)); chk(h->bs_cursor.reserve(SLQ_SORT_MAX_BUCKETS));
  chk(h->bs_offset.reserve(SLQ_SORT_MAX_BUCKETS + 1)); chk(h->bs_tmp.reserve(K));
  chk(h->sigma.reserve(d)); chk(h->sig_part.reserve((size_t)296 * 2 * d));
  chk(h->scalars.reserve(8));
  chk(h->ev_lt.reserve((size_t)h->A * KP)); chk(h->ev_segsum.reserve((size_t)h->A * nseg_max));
  chk(h->ev_part.reserve((size_t)h->A * nseg_max * 3)); chk(h->ev_tickets.reserve(h->A));
  chk(h->abs_state.reserve((size_t)h->A * EV_STATE)); chk(h->abs_out.reserve((size_t)h->A * 2));
  chk(h->ret_ll.reserve((size_t)h->ret_capacity));
  chk(h->ctl.reserve(3));
  if (cfg->keepRetiredPoints) {
    if (world != 1) chk(fail(THY_ERR_UNSUPPORTED, "keepRetiredPoints is single-GPU only in this version"));
    chk(h->ret_x.reserve((size_t)h->ret_capacity * d));
  }
  cudaStream_t s = m->stream;
  if (st == THY_OK) {
    bool ok = cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h->side2, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_fork2, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_sigma, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_rng, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) == cudaSuccess &&
              cudaMallocHost(&h->host_ao, (size_t)h->A * 2 * sizeof(double)) == cudaSuccess &&
              cudaMallocHost(&h->host_c, sizeof(int)) == cudaSuccess &&
              cudaMallocHost(&h->host_conv, sizeof(long long)) == cudaSuccess;
    for (auto& e : h->ev_round) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) st = fail(THY_ERR_CUDA, "SLQ stream / event setup failed");
    else *h->host_conv = -1;
    h->fused_walk = !cfg->disableFusedWalk && slq_walk_supported(m);
    // THY_SLQ_PREDRAW=1 draws a round's walk randomness ahead on a side stream (k_slq_walk_rng) for 1.25 K / world + 64
    // walkers (the rest, if a rank ever has more, draw in place).  Round 2 made it the default for sharded pools; since
    // the walk kernel draws four normals per Philox call and its in-place instantiation runs 16 warps per SM the
    // in-place form is faster at every size measured (125 000-point shard: 241 vs 281 us per round; 10^6: 1089 vs
    // 1202 us), so it is now an option only (profiles/r02_walk_rewrite.md).
    const char* pre_env = std::getenv("THY_SLQ_PREDRAW");
    const bool predraw = pre_env && pre_env[0] == '1';
    if (h->fused_walk && predraw && st == THY_OK) {
      const long long walkers = std::min<long long>(K, (5LL * K) / (4LL * world) + 64);
      h->pre_groups = (int)((walkers + 7) / 8);
      chk(h->zpre.reserve((size_t)h->pre_groups * slq_walk_rng_floats_per_group(m, h->M)));
      chk(h->lacc.reserve((size_t)h->pre_groups * h->M * 8));
    }
    const char* cs = std::getenv("THY_SLQ_CUB_SORT");
    h->cub_round_sort = cs && cs[0] == '1';
    if (st == THY_OK) {
      const BucketSortBuffers b{h->bs_minmax.ptr, h->bs_count.ptr, h->bs_cursor.ptr, h->bs_offset.ptr, h->bs_tmp.ptr};
      chk(slq_bucket_sort_init(b, s));
    }
    const char* ng = std::getenv("THY_SLQ_GRAPH");
    h->use_graph = !(ng && ng[0] == '0');
    const char* tr = std::getenv("THY_SLQ_TRACE");
    if (tr && tr[0] == '1') chk(thy_slq_set_trace(h, 1));
  }
  if (st == THY_OK) {
    // initial pool: prior draws (+ seeds in the first slots, SlqEngine.scala:384-389,409)
    chk(sample_priors_dev(m, Pl, rng_seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(h->rank + 1)), h->x.ptr));
    if (st == THY_OK && n_seeds > 0 && h->rank == 0) {
      const long long ns = std::min<long long>(n_seeds, Pl);
      if (cudaMemcpyAsync(h->x.ptr, seeds, (size_t)ns * d * sizeof(double), cudaMemcpyHostToDevice, s) != cudaSuccess)
        st = fail(THY_ERR_CUDA, "seed upload failed");
    }
  }
  if (st == THY_OK) {
    chk(slq_eval(h, Pl, h->x.ptr, h->ll.ptr, h->lpr.ptr));
    std::vector<double> sc = {0.5, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    bool ok = cudaMemsetAsync(h->accepted.ptr, 0, (size_t)K * sizeof(int), s) == cudaSuccess &&
              cudaMemsetAsync(h->sel_hist.ptr, 0, 6 * 2048 * sizeof(unsigned int), s) == cudaSuccess &&
              cudaMemsetAsync(h->sel_ticket.ptr, 0, sizeof(unsigned int), s) == cudaSuccess &&
              cudaMemsetAsync(h->sel_cursor.ptr, 0, sizeof(unsigned int), s) == cudaSuccess &&
              cudaMemsetAsync(h->sel_state.ptr, 0, sizeof(SelectState), s) == cudaSuccess &&
              cudaMemsetAsync(h->flags.ptr, 0, (size_t)Pl, s) == cudaSuccess &&
              cudaMemsetAsync(h->count_dev.ptr, 0, sizeof(int), s) == cudaSuccess &&
              cudaMemsetAsync(h->ev_tickets.ptr, 0, (size_t)h->A * sizeof(unsigned int), s) == cudaSuccess;
    if (!ok) st = fail(THY_ERR_CUDA, "SLQ initialisation failed");
    chk(h->scalars.upload(sc, s));
    std::vector<double> stt((size_t)h->A * EV_STATE, 0.0);
    for (int a = 0; a < h->A; ++a) stt[(size_t)a * EV_STATE + 5] = -INFINITY;
    chk(h->abs_state.upload(stt, s));
    std::vector<long long> ctl = {0, 0, -1};
    chk(h->ctl.upload(ctl, s));
    // First collective on the communicator here, not inside the first round: NCCL sets its channels up lazily at the
    // first call (hundreds of milliseconds), which would otherwise be charged to the simulation.
    if (st == THY_OK && world > 1) chk(comm_all_gather_f64(comm, h->ll.ptr, h->keys_all.ptr, (size_t)Pl, s));
    // Peer-memory exchange (THY_SLQ_P2P=0 keeps the NCCL all-gather): every rank maps every rank's
    // exchange buffer.  The verdict is collective, so either all ranks use it or none does.
    const char* p2p_env = std::getenv("THY_SLQ_P2P");
    if (st == THY_OK && world > 1 && !(p2p_env && p2p_env[0] == '0')) {
      const long long cap = K;
      chk(h->p2p_buf.reserve(p2p_bytes(world, cap)));
      if (st == THY_OK && cudaMemsetAsync(h->p2p_buf.ptr, 0, p2p_bytes(world, cap), s) != cudaSuccess) st = fail(THY_ERR_CUDA, "memset");
      if (st == THY_OK && cudaStreamSynchronize(s) != cudaSuccess) st = fail(THY_ERR_CUDA, "SLQ initialisation failed");
      std::vector<void*> ptrs;
      bool ok = false;
      if (st == THY_OK) chk(comm_map_peers(comm, h->p2p_buf.ptr, ptrs, h->p2p_opened, &ok, s));
      if (st == THY_OK && ok) {
        std::vector<unsigned long long> bases(world);
        for (int r = 0; r < world; ++r) bases[r] = (unsigned long long)ptrs[r];
        chk(h->p2p_bases.upload(bases, s));
        h->p2p_view.peer_base = h->p2p_bases.ptr;
        h->p2p_view.world = world;
        h->p2p_view.rank = h->rank;
        h->p2p_view.cap = cap;
        h->p2p = st == THY_OK;
      }
    }
    if (cudaStreamSynchronize(s) != cudaSuccess) st = fail(THY_ERR_CUDA, "SLQ initialisation failed");
  }
  if (st != THY_OK) {
    delete h;
    return st;
  }
  *out = h;
  return THY_OK;
}

thy_status thy_slq_destroy(thy_slq_t h) {
  if (h && h->p2p && h->world > 1 && comm_is_live(h->comm)) {
    // peers hold mappings of this rank's exchange buffer and may still be writing into it: destroy is collective
    cudaStreamSynchronize(h->model->stream);
    comm_all_reduce_f64(h->comm, h->scalars.ptr + 3, 1, 0, h->model->stream);
    cudaStreamSynchronize(h->model->stream);
  }
  if (h && std::getenv("THY_SLQ_TIMING") && h->rounds > 0)
    std::fprintf(stderr, "[thy slq timing] rank %d at destroy: %lld rounds, graph %s; host: %.3f s waiting for round r-2, %.3f s "
                 "enqueuing\n", h->rank, h->rounds, h->graph ? "on" : "off", h->host_wait_s, h->host_enqueue_s);
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (h && h->trace && h->tcount > 0 && std::getenv("THY_SLQ_TRACE")) {
    std::fprintf(stderr, "[thy slq trace] rank %d: %lld rounds; us/round:", h->rank, h->tcount);
    for (int i = 0; i < 6; ++i) std::fprintf(stderr, " %s %.1f%s", kPhaseNames[i], 1e3 * h->tsum[i] / h->tcount, i < 5 ? "," : "\n");
  }
  delete h;
  return THY_OK;
}

thy_status thy_slq_set_trace(thy_slq_t h, int enable) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (enable && !h->tev[0])
    for (auto& e : h->tev)
      if (cudaEventCreate(&e) != cudaSuccess) return fail(THY_ERR_CUDA, "SLQ trace events");
  h->trace = enable != 0;
  return THY_OK;
}

thy_status thy_slq_phase_times(thy_slq_t h, double out_us[7], int64_t* out_rounds) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  for (int i = 0; i < 7; ++i)
    if (out_us) out_us[i] = h->tcount > 0 ? 1e3 * h->tsum[i] / h->tcount : 0.0;
  // bit 0: untraced full rounds replay from a CUDA graph; bit 1: peer-memory exchange instead of the NCCL all-gather
  if (out_us) out_us[6] = (h->graph ? 1.0 : 0.0) + (h->p2p ? 2.0 : 0.0);
  if (out_rounds) *out_rounds = h->tcount;
  return THY_OK;
}

thy_status thy_slq_load_balance(thy_slq_t h, double* out_mean_local, double* out_max_local) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  double sc[8];
  THY_CUDA_CHECK(cudaStreamSynchronize(h->model->stream));
  THY_CUDA_CHECK(cudaMemcpy(sc, h->scalars.ptr, sizeof(sc), cudaMemcpyDeviceToHost));
  if (out_mean_local) *out_mean_local = h->rounds > 0 ? sc[3] / (double)h->rounds : 0.0;
  if (out_max_local) *out_max_local = sc[2];
  return THY_OK;
}

thy_status thy_slq_run_rounds(thy_slq_t h, int n_rounds, int* out_converged) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (h->finished) return fail(THY_ERR_STATE, "the simulation has already been drained");
  if (h->cfg.referenceProposals) return fail(THY_ERR_UNSUPPORTED, "referenceProposals runs as a whole (thy_slq_run)");
  bool conv = false;
  THY_TRY(slq_live_phase(h, n_rounds, &conv));
  if (out_converged) *out_converged = conv ? 1 : 0;
  return THY_OK;
}

thy_status thy_slq_run(thy_slq_t h, thy_slq_stats* out_stats) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (h->finished) {
    if (out_stats) *out_stats = h->stats;
    return THY_OK;
  }
  if (h->cfg.referenceProposals) {
    THY_TRY(slq_run_reference(h));
    if (out_stats) *out_stats = h->stats;
    return THY_OK;
  }
  cudaStream_t s = h->model->stream;
  std::vector<double> ao((size_t)h->A * 2);
  bool conv = false;
  const bool timing = std::getenv("THY_SLQ_TIMING") != nullptr;   // host wall-clock split of the run on stderr
  const auto t_live0 = std::chrono::steady_clock::now();
  THY_TRY(slq_live_phase(h, (long long)1 << 60, &conv));
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  const auto t_live1 = std::chrono::steady_clock::now();
  // drainSamplePool (SlqEngine.scala:171-180): the remaining live points in ascending order
  const double* keys = h->ll.ptr;
  if (h->world > 1) {
    THY_TRY(comm_all_gather_f64(h->comm, h->ll.ptr, h->keys_all.ptr, (size_t)h->Pl, s));
    keys = h->keys_all.ptr;
  }
  if (h->iterations + h->P > h->ret_capacity) return fail(THY_ERR_STATE, "SLQ: retired buffer exhausted");
  if (h->cfg.keepRetiredPoints && h->world == 1) {
    k_iota<<<(unsigned)ceil_div(h->Pl, 256), 256, 0, s>>>((int)h->Pl, h->pair_in.ptr);
    THY_TRY(slq_sort_pairs(h, keys, h->vals_sorted.ptr, h->pair_in.ptr, h->pair_out.ptr, (int)h->P, s));
    k_slq_store<<<(unsigned)ceil_div(h->Pl, 8), 256, 0, s>>>((int)h->Pl, h->d, h->pair_out.ptr, h->x.ptr,
                                                            h->ret_x.ptr + (size_t)h->iterations * h->d);
    count_launch(2);
  } else {
    THY_TRY(slq_sort_values(h, keys, h->vals_sorted.ptr, (int)h->P, s));
  }
  THY_CUDA_CHECK(cudaMemcpyAsync(h->ret_ll.ptr + h->iterations, h->vals_sorted.ptr, (size_t)h->P * sizeof(double),
                                 cudaMemcpyDeviceToDevice, s));
  h->live_iterations = h->iterations;
  THY_TRY(slq_evidence_launch(ev_buffers(h), h->A, (int)h->P, h->P, h->K, h->live_iterations, h->iterations, nullptr, h->seed,
                              h->vals_sorted.ptr, 1, s));
  h->iterations += h->P;
  std::vector<double> sc(8);
  THY_CUDA_CHECK(cudaMemcpyAsync(ao.data(), h->abs_out.ptr, ao.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
  THY_CUDA_CHECK(cudaMemcpyAsync(sc.data(), h->scalars.ptr, 8 * sizeof(double), cudaMemcpyDeviceToHost, s));
  double minl = 0.0;
  THY_CUDA_CHECK(cudaMemcpyAsync(&minl, h->ret_ll.ptr + h->live_iterations, sizeof(double), cudaMemcpyDeviceToHost, s));
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  if (sc[7] != 0.0) return fail(THY_ERR_STATE, "SLQ: a peer rank did not signal within 5 s (peer-memory exchange); results are invalid");
  double mean = 0.0, m2 = 0.0, hm = 0.0;
  for (int a = 0; a < h->A; ++a) {
    mean += ao[(size_t)a * 2];
    hm += ao[(size_t)a * 2 + 1];
  }
  mean /= h->A;
  hm /= h->A;
  for (int a = 0; a < h->A; ++a) m2 += (ao[(size_t)a * 2] - mean) * (ao[(size_t)a * 2] - mean);
  h->stats.iterations = h->live_iterations;
  h->stats.total_retired = h->iterations;
  h->stats.rounds = h->rounds;
  h->stats.log_evidence_mean = mean;
  h->stats.log_evidence_std = h->A > 1 ? std::sqrt(m2 / (h->A - 1)) : 0.0;
  h->stats.neg_entropy_mean = hm;
  h->stats.min_logpdf = minl;
  h->stats.walk_scale = sc[0];
  h->stats.walk_acceptance = sc[6] > 0 ? sc[5] / sc[6] : 0.0;
  h->stats.likelihood_evaluations = h->evals;
  if (timing)
    std::fprintf(stderr, "[thy slq timing] rank %d: live phase %.3f s (%lld rounds, graph %s; host: %.3f s waiting for round r-2, "
                 "%.3f s enqueuing), drain + statistics %.3f s\n", h->rank,
                 std::chrono::duration<double>(t_live1 - t_live0).count(), h->rounds, h->graph ? "on" : "off", h->host_wait_s,
                 h->host_enqueue_s, std::chrono::duration<double>(std::chrono::steady_clock::now() - t_live1).count());
  h->finished = true;
  if (out_stats) *out_stats = h->stats;
  return THY_OK;
}

/* Retired log-likelihoods in retirement order (live phase followed by the drained pool). */
thy_status thy_slq_retired(thy_slq_t h, int64_t max_count, double* out_loglik, int64_t* out_count) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  const long long n = std::min<long long>(max_count, h->iterations);
  if (out_loglik && n > 0) {
    THY_CUDA_CHECK(cudaStreamSynchronize(h->model->stream));
    THY_CUDA_CHECK(cudaMemcpy(out_loglik, h->ret_ll.ptr, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (out_count) *out_count = h->iterations;
  return THY_OK;
}

/* ln X_i of one simulated abscissa for the retired points, in retirement order (QuadratureAbscissa.scala:33-73): the
 * prior-mass simulation is replayed on the device from the counter RNG, so nothing had to be stored during the run. */
thy_status thy_slq_retired_log_x(thy_slq_t h, int abscissa, int64_t max_count, double* out_log_x, int64_t* out_count) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (abscissa < 0 || abscissa >= h->A) return fail(THY_ERR_INVALID_ARGUMENT, "abscissa index out of range");
  const long long n = std::min<long long>(max_count, h->iterations);
  if (out_log_x && n > 0) THY_TRY(slq_export_logx(h, abscissa, n, out_log_x));
  if (out_count) *out_count = h->iterations;
  return THY_OK;
}

/* The rank's live points and their log-likelihoods (the reference's `samples` map, SlqEngine.scala:120-143). */
thy_status thy_slq_live_points(thy_slq_t h, int64_t* out_count, double* out_x, double* out_loglik) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (out_count) *out_count = h->Pl;
  THY_CUDA_CHECK(cudaStreamSynchronize(h->model->stream));
  if (out_x) THY_CUDA_CHECK(cudaMemcpy(out_x, h->x.ptr, (size_t)h->Pl * h->d * sizeof(double), cudaMemcpyDeviceToHost));
  if (out_loglik) THY_CUDA_CHECK(cudaMemcpy(out_loglik, h->ll.ptr, (size_t)h->Pl * sizeof(double), cudaMemcpyDeviceToHost));
  return THY_OK;
}

/* Per-abscissa results of the quadrature so far: ln Z_a and H_a (QuadratureIntegrator.scala:29-62). */
thy_status thy_slq_abscissa_results(thy_slq_t h, double* out_log_evidence, double* out_neg_entropy) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  std::vector<double> ao((size_t)h->A * 2);
  THY_CUDA_CHECK(cudaStreamSynchronize(h->model->stream));
  THY_CUDA_CHECK(cudaStreamSynchronize(h->side));
  THY_CUDA_CHECK(cudaMemcpy(ao.data(), h->abs_out.ptr, ao.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int a = 0; a < h->A; ++a) {
    if (out_log_evidence) out_log_evidence[a] = ao[(size_t)a * 2];
    if (out_neg_entropy) out_neg_entropy[a] = ao[(size_t)a * 2 + 1];
  }
  return THY_OK;
}

/* integrate (SlqEngine.scala:460-466 -> QuadratureIntegrator.getIntegrationStats, QuadratureIntegrator.scala:64-83):
 * per simulated abscissa a,  I_a = sum_i w_{a,i} f(exp(l_i - shift)),  Z_a = sum_i w_{a,i} exp(l_i - shift)  with the
 * trapezoid weights of that abscissa; the reference returns the mean over a of  I_a / Z_a - ln Z_a  (literally what
 * getIntegrationStats computes) -> out_reference_statistic; the plain quadrature mean over a of I_a -> out_integral.
 * log_shift: NaN = the reference's (max l + min l) / 2 (its BigDecimal arithmetic cannot overflow; in fp64 that shift can,
 * so callers may pass e.g. max l).  The integrand is a host callback, as the reference's is a JVM closure. */
thy_status thy_slq_integrate(thy_slq_t h, thy_slq_integrand integrand, void* user, double log_shift,
                             double* out_reference_statistic, double* out_integral, double* out_per_abscissa_integral) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (!integrand) return fail(THY_ERR_INVALID_ARGUMENT, "null integrand");
  if (!h->finished) return fail(THY_ERR_STATE, "run the simulation first (rebuildSampleSimulation)");
  const long long N = h->iterations;
  std::vector<double> l((size_t)N), lx((size_t)N);
  THY_CUDA_CHECK(cudaMemcpy(l.data(), h->ret_ll.ptr, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost));
  double lmax = -INFINITY, lmin = INFINITY;
  for (double v : l) {
    if (v > lmax) lmax = v;
    if (v < lmin && v > -INFINITY) lmin = v;
  }
  const double shift = std::isnan(log_shift) ? 0.5 * (lmax + lmin) : log_shift;
  std::vector<double> fv((size_t)N), ev((size_t)N);
  for (long long i = 0; i < N; ++i) {
    ev[(size_t)i] = std::exp(l[(size_t)i] - shift);
    fv[(size_t)i] = integrand(ev[(size_t)i], user);
  }
  long double stat = 0.0L, integ = 0.0L;
  for (int a = 0; a < h->A; ++a) {
    THY_TRY(slq_export_logx(h, a, N, lx.data()));
    long double Ia = 0.0L, Za = 0.0L;
    for (long long i = 0; i < N; ++i) {
      const double xm = std::exp(lx[(size_t)(i == 0 ? 0 : i - 1)]), xp = std::exp(lx[(size_t)(i + 1 < N ? i + 1 : i)]);
      const double w = 0.5 * (xm - xp);
      Ia += (long double)w * fv[(size_t)i];
      Za += (long double)w * ev[(size_t)i];
    }
    if (out_per_abscissa_integral) out_per_abscissa_integral[a] = (double)Ia;
    stat += Ia / Za - logl(Za);
    integ += Ia;
  }
  if (out_reference_statistic) *out_reference_statistic = (double)(stat / h->A);
  if (out_integral) *out_integral = (double)(integ / h->A);
  return THY_OK;
}

/* sample(n) (SlqEngine.scala:452-469 -> SamplingSimulation.scala:41-101): every draw picks one of the simulated
 * abscissas uniformly, then a retired point with probability proportional to that abscissa's trapezoid area
 * 1/2 (L_i + L_next)(X_i - X_next) (the innermost point: L X).  The staircases are built on the host from the exported
 * ln X; the chosen points are gathered on the device and copied back in ONE transfer.  Requires keepRetiredPoints. */
thy_status thy_slq_sample(thy_slq_t h, int n, uint64_t rng_seed, double* out_samples) {
  thy::ErrScope _err_scope(h ? h->model : nullptr);
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (!h->finished) return fail(THY_ERR_STATE, "run the simulation first (rebuildSampleSimulation)");
  if (!h->cfg.keepRetiredPoints || h->world != 1) return fail(THY_ERR_UNSUPPORTED, "retired points were not kept");
  if (n <= 0 || !out_samples) return fail(THY_ERR_INVALID_ARGUMENT, "bad arguments");
  const long long N = h->iterations;
  std::vector<double> l((size_t)N), lx((size_t)N), cdf((size_t)N);
  THY_CUDA_CHECK(cudaMemcpy(l.data(), h->ret_ll.ptr, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost));
  // draws grouped by abscissa
  std::vector<int> which((size_t)n);
  std::vector<double> uu((size_t)n);
  for (int k = 0; k < n; ++k) {
    const RngDraw u = rng_uniform2(rng_seed, (uint64_t)k, 0, RNG_SLQ_PICK, 0xFFFFFFFFull);
    int a = (int)(u.u1 * h->A);
    which[(size_t)k] = a >= h->A ? h->A - 1 : a;
    uu[(size_t)k] = u.u0;
  }
  std::vector<long long> idx((size_t)n, 0);
  for (int a = 0; a < h->A; ++a) {
    bool used = false;
    for (int k = 0; k < n && !used; ++k) used = which[(size_t)k] == a;
    if (!used) continue;
    THY_TRY(slq_export_logx(h, a, N, lx.data()));
    // log weight of point i: trapezoid between X_i and X_{i+1} (X_N = 0)
    double mx = -INFINITY;
    for (long long i = 0; i < N; ++i) {
      double lw;
      if (i + 1 < N) {
        const double la = l[(size_t)i], lb = l[(size_t)i + 1];
        const double lsum = lb + std::log1p(std::exp(la - lb));   // ln(L_i + L_{i+1}), l ascending
        lw = std::log(0.5) + lsum + lx[(size_t)i] + std::log(-std::expm1(lx[(size_t)i + 1] - lx[(size_t)i]));
      } else {
        lw = l[(size_t)i] + lx[(size_t)i];
      }
      cdf[(size_t)i] = lw;
      if (lw > mx) mx = lw;
    }
    double run = 0.0;
    for (long long i = 0; i < N; ++i) {
      run += std::exp(cdf[(size_t)i] - mx);
      cdf[(size_t)i] = run;
    }
    for (int k = 0; k < n; ++k) {
      if (which[(size_t)k] != a) continue;
      const double target = uu[(size_t)k] * run;
      idx[(size_t)k] = std::min<long long>(N - 1, std::upper_bound(cdf.begin(), cdf.end(), target) - cdf.begin());
    }
  }
  cudaStream_t s = h->model->stream;
  DevBuf<long long> idx_dev;
  DevBuf<double> out_dev;
  THY_TRY(idx_dev.upload(idx, s));
  THY_TRY(out_dev.reserve((size_t)n * h->d));
  k_slq_gather_rows<<<(unsigned)ceil_div(n, 8), 256, 0, s>>>(n, h->d, idx_dev.ptr, h->ret_x.ptr, out_dev.ptr);
  count_launch();
  THY_CUDA_CHECK(cudaMemcpyAsync(out_samples, out_dev.ptr, (size_t)n * h->d * sizeof(double), cudaMemcpyDeviceToHost, s));
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  return THY_OK;
}

}  // extern "C"
