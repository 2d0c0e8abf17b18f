// SLQ (stochastic Lebesgue quadrature = nested sampling) with the live-point threshold step on the device.
// Reference: /root/reference/src/main/scala/thylacine/model/integration/slq/
//   SlqEngine.scala:120-197        live pool keyed by logPdf; discard the global minimum; replace it by a new draw
//                                  accepted only above the threshold; per-acceptance abscissa extension
//   SlqEngine.scala:249-277        convergence: iterations > minIterationCount && iterations >= 10 P H, or >= maxIterationCount
//   SlqEngine.scala:171-180        drainSamplePool: remaining live points appended in ascending order
//   QuadratureAbscissa.scala:33-73 prior-mass simulation: pool of P uniforms, max moved to the abscissa list (first = 1.0),
//                                  trapezoid weights 1/2 (X_{i-1} - X_{i+1}), ends half the adjacent gap
//   QuadratureIntegrator.scala:29-62  Z_a = sum w_i exp(l_i - s), H_a = sum w_i e^{l_i - s}(l_i - s)/Z_a - ln Z_a
// Re-design for 10^6 live points (DESIGN.md "SLQ"): the K lowest live points are retired per round (K = sampleParallelism,
// the reference's number of concurrent sampling fibres); thresholds come from a device radix sort plus, across GPUs, an
// NCCL all-gather of each rank's K smallest; replacements are constrained random walks started from surviving live
// points (the reference's O(P^2 d) cubical cover cannot exist at this size); prior-mass shrinkage is sampled directly as
// t = u^{1/n_live} per abscissa in log space and the evidence is a streaming log-sum-exp (no BigDecimal needed).
// Measure / integrand: points are drawn from the PRIOR and ranked by log-LIKELIHOOD, so ln Z is the calibrated
// evidence (for uniform priors this coincides with the reference's uniform-domain / full-density formulation, Q6).
#include "kernels.h"
#include "comm.h"
#include "philox.cuh"
#include <cub/cub.cuh>
#include <cmath>
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
  thy::DevBuf<double> keys_in, keys_out;                 // sort scratch
  thy::DevBuf<int> idx_in, idx_out;
  thy::DevBuf<unsigned char> cub_tmp;
  thy::DevBuf<double> cand_all, cand_sorted;             // gathered candidates
  thy::DevBuf<double> Y, Yp, yl, ypr, tl, tpr;           // walkers
  thy::DevBuf<int> slot, accepted;
  thy::DevBuf<double> sigma, sig_part;                   // per-dimension spread of the live pool
  thy::DevBuf<double> scalars;                           // [0] walk scale, [1] threshold, [2] accepted, [3] proposed
  thy::DevBuf<int> count_dev;                            // [0] local retire count
  thy::DevBuf<double> tmp_logx;                          // [A][Kmax]
  thy::DevBuf<double> abs_state;                         // [A][8]: logX_cur, Xm2, Xm1, pend_l, has_pend(0/1/2), m, s, s_l
  thy::DevBuf<double> abs_out;                           // [A][2]: logZ, H
  thy::DevBuf<double> ret_ll;                            // retired log-likelihoods (global order)
  thy::DevBuf<double> ret_x;                             // retired points (optional, world == 1)
  thy::DevBuf<double> ret_logw;                          // log posterior weight under the expected abscissa
  long long iterations = 0, rounds = 0, evals = 0;
  long long ret_capacity = 0;
  bool fused_walk = false;             // slq_walk.cu covers this model: one kernel per round instead of 5 per walk step
  cudaStream_t side = nullptr;         // evidence accumulation runs beside the walk (it only needs the retired values)
  cudaEvent_t ev_retired = nullptr, ev_evidence = nullptr;
  double* host_ao = nullptr;           // pinned [A][2] copy of abs_out for the convergence test
  int* host_c = nullptr;               // pinned copy of the local retire count
  long long* host_it = nullptr;        // pinned staging of `iterations` for the device-side round control
  thy::DevBuf<long long> ctl;          // [0] iterations at the start of the round (read by graph-captured kernels)
  cudaStream_t select_graph_stream = nullptr;
  cudaGraphExec_t select_graph = nullptr;   // the selection phase of a full round (sigma, sort, gather, merge, threshold)
  bool use_graph = true;
  // THY_SLQ_TRACE=1: CUDA-event phase times of the round (sort / gather / merge+threshold / replace), printed at destroy
  bool trace = false;
  cudaEvent_t tev[6] = {};
  double tsum[5] = {0, 0, 0, 0, 0};
  long long tcount = 0;
  // THY_SLQ_TRACE=2: host wall-clock split of the round (graph replay stays on): enqueue / wait for the retire count /
  // enqueue of the replacement step / wait for the evidence update
  bool htrace = false;
  double hsum[4] = {0, 0, 0, 0};
  long long hcount = 0;
  bool finished = false;
  ~thy_slq_s() {
    if (side) cudaStreamDestroy(side);
    if (ev_retired) cudaEventDestroy(ev_retired);
    if (ev_evidence) cudaEventDestroy(ev_evidence);
    if (host_ao) cudaFreeHost(host_ao);
    if (host_c) cudaFreeHost(host_c);
    if (host_it) cudaFreeHost(host_it);
    if (select_graph) cudaGraphExecDestroy(select_graph);
    if (htrace && hcount > 0)
      std::fprintf(stderr, "[thy slq host trace] rank %d: %lld rounds, graph %s; host us/round: enqueue selection %.1f, wait count %.1f, "
                   "enqueue replacement %.1f, wait evidence %.1f\n", rank, hcount, select_graph ? "on" : "off",
                   1e6 * hsum[0] / hcount, 1e6 * hsum[1] / hcount, 1e6 * hsum[2] / hcount, 1e6 * hsum[3] / hcount);
    if (trace) {
      if (tcount > 0)
        std::fprintf(stderr, "[thy slq trace] rank %d: %lld rounds; us/round: sigma %.1f, local sort %.1f, all-gather %.1f, "
                     "merge+threshold %.1f, replace+adapt %.1f\n", rank, tcount, 1e3 * tsum[0] / tcount, 1e3 * tsum[1] / tcount,
                     1e3 * tsum[2] / tcount, 1e3 * tsum[3] / tcount, 1e3 * tsum[4] / tcount);
      for (auto& e : tev) if (e) cudaEventDestroy(e);
    }
  }
  thy_slq_stats stats{};
};

namespace thy {

__global__ void k_iota(int n, int* v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

__global__ void k_fill(long long n, double* v, double val) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = val;
}

// threshold = merged[Kr-1].  Exactly Kr points are retired globally even when log-likelihoods tie (a walker that never
// moved is an exact copy of a survivor; the reference nudges duplicates by one ulp instead, SlqEngine.scala:107-118):
// every value < threshold goes, and the ties at the threshold are taken in rank order, lowest sorted index first.
__device__ __forceinline__ long long lower_bound_dev(const double* v, long long n, double x) {  // first index with v >= x
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (v[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ long long upper_bound_dev(const double* v, long long n, double x) {  // first index with v > x
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (v[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// Merge of the ranks' sorted candidate lists: element (r, i) computes its position in the union -- values from lower
// ranks win ties, then the position inside the list -- and the Kr smallest land in `merged` in ascending order.
// One launch instead of a radix sort of Kr * world keys.
__global__ void __launch_bounds__(256) k_slq_merge(const double* __restrict__ cand_all, int Kr, int world,
                                                   double* __restrict__ merged) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)Kr * world) return;
  const int r = (int)(e / Kr), i = (int)(e % Kr);
  const double v = cand_all[e];
  long long pos = i;
  for (int q = 0; q < world; ++q) {
    if (q == r) continue;
    const double* lst = cand_all + (size_t)q * Kr;
    // the smallest value of a list that is already beyond v contributes nothing; skip its search
    if (lst[0] > v) continue;
    pos += (q < r) ? upper_bound_dev(lst, Kr, v) : lower_bound_dev(lst, Kr, v);
    if (pos >= Kr) return;
  }
  if (pos < Kr) merged[pos] = v;
}

// retired values recorded in global order at ret_ll[iterations ...] (iterations read from the device-side control
// word so that the launch can live inside a CUDA graph)
__global__ void __launch_bounds__(256) k_slq_copy_retired(const double* __restrict__ merged, int Kr,
                                                          const long long* __restrict__ ctl, double* __restrict__ ret_ll) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Kr) ret_ll[ctl[0] + i] = merged[i];
}

__global__ void k_slq_threshold(const double* __restrict__ merged, int Kr, const double* __restrict__ local_sorted,
                                const double* __restrict__ cand_all, int world, int rank,
                                double* __restrict__ scalars, int* __restrict__ count) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double thr = merged[Kr - 1];
  scalars[1] = thr;
  if (world == 1) {
    count[0] = Kr;
    return;
  }
  const long long below = lower_bound_dev(merged, (long long)Kr * world, thr);
  long long need = Kr - below;  // ties still to take, >= 1
  long long mine = 0;
  for (int r = 0; r <= rank; ++r) {
    const double* cr = cand_all + (size_t)r * Kr;
    const long long ties = upper_bound_dev(cr, Kr, thr) - lower_bound_dev(cr, Kr, thr);
    const long long take = ties < need ? ties : need;
    if (r == rank) mine = take;
    need -= take;
  }
  count[0] = (int)(lower_bound_dev(local_sorted, Kr, thr) + mine);
}

// Per-dimension standard deviation of the live pool: deterministic two-stage reduction.
__global__ void __launch_bounds__(256) k_slq_sigma_part(long long Pl, int d, const double* __restrict__ x,
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
        const double v = x[r * d + j];
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
        const double v = x[r * d + j];
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

// Walkers start from uniformly chosen survivors.
__global__ void __launch_bounds__(256) k_slq_spawn(int c, long long Pl, int d, uint64_t seed, uint64_t round, int rank,
                                                   const int* __restrict__ sidx, const double* __restrict__ x,
                                                   const double* __restrict__ ll, const double* __restrict__ lpr,
                                                   double* __restrict__ Y, double* __restrict__ yl, double* __restrict__ ypr,
                                                   int* __restrict__ slot) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= c) return;
  const RngDraw u = rng_uniform2(seed, ((uint64_t)rank << 40) | (uint64_t)w, 0, RNG_SLQ_PICK, round);
  long long pick = c + (long long)(u.u0 * (double)(Pl - c));
  if (pick >= Pl) pick = Pl - 1;
  const int src = sidx[pick];
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
  rng_normal_pair_fast(seed, ((uint64_t)rank << 40) | (uint64_t)w, slq_move_item(j), RNG_SLQ_MOVE, step, z0, z1);
  Yp[i] = Y[i] + scalars[0] * sigma[j] * (slq_move_branch(j) ? z1 : z0);
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
__global__ void __launch_bounds__(1024) k_slq_adapt(double* scalars, double inc, double target, int c, int steps,
                                                    int* __restrict__ accepted) {
  __shared__ double sh[32];
  double a = 0.0;
  for (int w = threadIdx.x; w < c; w += blockDim.x) {
    a += (double)accepted[w];
    accepted[w] = 0;
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
}

// Copy retired points (local sorted order == global order when world == 1).
__global__ void __launch_bounds__(256) k_slq_store(int c, int d, const int* __restrict__ sidx, const double* __restrict__ x,
                                                   double* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= c) return;
  const int s = sidx[w];
  for (int j = lane; j < d; j += 32) dst[(size_t)w * d + j] = x[(size_t)s * d + j];
}

// ---- evidence accumulation -------------------------------------------------------------------------------------
// One block per abscissa.  r[0..Kr) = retired log-likelihoods of this round in ascending order; the j-th of them was
// retired with n0 - j points alive, so the prior mass shrinks by t = u^{1/(n0-j)} AFTER it has been assigned its X
// (the reference assigns X = 1.0 to the very first retired point, QuadratureAbscissa.scala:64-73).
// state: [0] logX assigned to the next point, [1] logX_{i-2}, [2] logX_{i-1}, [3] l_{i-1}, [4] pending count (0,1,2+),
//        [5] running max m, [6] sum exp(term - m), [7] sum exp(term - m) * l
__device__ __forceinline__ void lse_add(double& m, double& s, double& sl, double term, double l) {
  if (!(term > -INFINITY)) return;
  if (term > m) {
    const double f = exp(m - term);
    s = s * f + 1.0;
    sl = sl * f + l;
    m = term;
  } else {
    const double e = exp(term - m);
    s += e;
    sl += e * l;
  }
}
__device__ __forceinline__ double log_half_gap(double lx_hi, double lx_lo) {
  // ln( 1/2 (X_hi - X_lo) ),  X_hi >= X_lo given as logs
  return log(0.5) + lx_hi + log(-expm1(lx_lo - lx_hi));
}

constexpr int EVT = 1024;  // threads of the evidence kernel

__global__ void __launch_bounds__(EVT) k_slq_evidence(int Kr, long long n0, long long it0, uint64_t seed,
                                                      const double* __restrict__ r, double* __restrict__ tmp_logx,
                                                      long long tmp_stride, double* __restrict__ state, int finalize,
                                                      double* __restrict__ out) {
  __shared__ double wsum[EVT / 32];
  __shared__ double carry;
  __shared__ double red_m[EVT], red_s[EVT], red_sl[EVT];
  const int a = blockIdx.x;
  double* st = state + (size_t)a * 8;
  double* lx = tmp_logx + (size_t)a * tmp_stride;
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  // pass 1: logX_j = logX_cur + sum_{k<j} ln t_k.  Each thread owns a contiguous segment: serial sum, block scan of
  // the segment totals (warp shuffles), serial write-back.
  {
    const int S = (Kr + EVT - 1) / EVT;
    const int j0 = tid * S < Kr ? tid * S : Kr, j1 = (j0 + S < Kr) ? j0 + S : Kr;
    double local = 0.0;
    for (int j = j0; j < j1; ++j) {
      const RngDraw u = rng_uniform2(seed, (uint64_t)a, 0, RNG_SLQ_SHRINK, (uint64_t)(it0 + j));
      const double lt = log(1.0 - u.u0) / (double)(n0 - j);   // ln u^{1/n}, u in (0,1]
      lx[j] = lt;
      local += lt;
    }
    double incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[wrp] = incl;
    __syncthreads();
    double base = st[0];
    for (int w = 0; w < wrp; ++w) base += wsum[w];
    double run = base + (incl - local);
    for (int j = j0; j < j1; ++j) {
      const double lt = lx[j];
      lx[j] = run;
      run += lt;
    }
    if (tid == EVT - 1) carry = run;   // logX assigned to the first point of the next round
    __syncthreads();
  }
  // pass 2: finalise point j-1 once X_j is known: w_{j-1} = 1/2 (X_{j-2} - X_j)  (first ever point: 1/2 (X_0 - X_1))
  double m = -INFINITY, s = 0.0, sl = 0.0;
  const double pend = st[4];
  for (int j = tid; j < Kr; j += EVT) {
    double l_prev, lx_m2;
    bool have_prev = true, first_ever = false;
    if (j == 0) {
      if (pend < 0.5) have_prev = false;
      l_prev = st[3];
      lx_m2 = st[1];
      first_ever = pend < 1.5;   // pending point is the very first retired point: no X_{i-2}
      if (first_ever) lx_m2 = st[2];  // use X_{i-1} itself: 1/2 (X_0 - X_1)
    } else if (j == 1) {
      l_prev = r[0];
      lx_m2 = st[2];
      if (pend < 0.5) { first_ever = true; lx_m2 = lx[0]; }
    } else {
      l_prev = r[j - 1];
      lx_m2 = lx[j - 2];
    }
    if (have_prev) {
      const double term = l_prev + log_half_gap(lx_m2, lx[j]);
      lse_add(m, s, sl, term, l_prev);
    }
  }
  red_m[tid] = m; red_s[tid] = s; red_sl[tid] = sl;
  __syncthreads();
  for (int o = EVT / 2; o > 0; o >>= 1) {
    if (tid < o) {
      double m2 = red_m[tid + o], s2 = red_s[tid + o], sl2 = red_sl[tid + o];
      double m1 = red_m[tid], s1 = red_s[tid], sl1 = red_sl[tid];
      if (m2 > m1) { const double f = exp(m1 - m2); s1 = s1 * f + s2; sl1 = sl1 * f + sl2; m1 = m2; }
      else if (m2 > -INFINITY) { const double f = exp(m2 - m1); s1 += s2 * f; sl1 += sl2 * f; }
      red_m[tid] = m1; red_s[tid] = s1; red_sl[tid] = sl1;
    }
    __syncthreads();
  }
  if (tid == 0) {
    double M = st[5], S = st[6], SL = st[7];
    if (red_m[0] > -INFINITY) {
      if (red_m[0] > M) { const double f = exp(M - red_m[0]); S = S * f + red_s[0]; SL = SL * f + red_sl[0]; M = red_m[0]; }
      else { const double f = exp(red_m[0] - M); S += red_s[0] * f; SL += red_sl[0] * f; }
    }
    if (Kr > 0) {
      // new pending point = last of this round
      const double new_m1 = lx[Kr - 1];
      const double new_m2 = (Kr >= 2) ? lx[Kr - 2] : st[2];
      const double newpend = (Kr >= 2) ? 2.0 : (pend + 1.0);
      st[1] = new_m2;
      st[2] = new_m1;
      st[3] = r[Kr - 1];
      st[4] = newpend > 2.0 ? 2.0 : newpend;
      st[0] = carry;
    }
    if (finalize && st[4] > 1.5) {
      // last point: 1/2 (X_{N-2} - X_{N-1})
      const double term = st[3] + log_half_gap(st[1], st[2]);
      lse_add(M, S, SL, term, st[3]);
      st[4] = 0.0;
    }
    st[5] = M; st[6] = S; st[7] = SL;
    const double logZ = (S > 0.0) ? M + log(S) : -INFINITY;
    out[a * 2 + 0] = logZ;
    out[a * 2 + 1] = (S > 0.0) ? SL / S - logZ : 0.0;
  }
}

static thy_status slq_sort_local(thy_slq_s* h) {
  cudaStream_t s = h->model->stream;
  THY_CUDA_CHECK(cudaMemcpyAsync(h->keys_in.ptr, h->ll.ptr, h->Pl * sizeof(double), cudaMemcpyDeviceToDevice, s));
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, h->keys_in.ptr, h->keys_out.ptr, h->idx_in.ptr, h->idx_out.ptr, (int)h->Pl, 0, 64, s);
  THY_TRY(h->cub_tmp.reserve(bytes));
  THY_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(h->cub_tmp.ptr, bytes, h->keys_in.ptr, h->keys_out.ptr, h->idx_in.ptr,
                                                 h->idx_out.ptr, (int)h->Pl, 0, 64, s));
  return THY_OK;
}

static thy_status slq_sort_keys(thy_slq_s* h, const double* in, double* out, int n) {
  cudaStream_t s = h->model->stream;
  size_t bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, bytes, in, out, n, 0, 64, s);
  THY_TRY(h->cub_tmp.reserve(bytes));
  THY_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(h->cub_tmp.ptr, bytes, in, out, n, 0, 64, s));
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
  h->evals += n;
  return THY_OK;
}

static thy_status slq_evidence(thy_slq_s* h, const double* retired_sorted, int Kr, long long n0, int finalize,
                               cudaStream_t s) {
  const long long stride = std::max<long long>(h->K, h->P);
  k_slq_evidence<<<h->A, EVT, 0, s>>>(Kr, n0, h->iterations, h->seed, retired_sorted, h->tmp_logx.ptr, stride,
                                      h->abs_state.ptr, finalize, h->abs_out.ptr);
  count_launch();
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

// Selection phase of a round: proposal spread of the live pool, local sort, exchange of each rank's Kr smallest, merge,
// threshold and local retire count, retired values appended.  Every launch has round-independent arguments (the
// append position comes from h->ctl), so a full round's selection phase is captured once as a CUDA graph and replayed:
// ~25 launches become one, which is what keeps 8 ranks from being host-launch bound.
static thy_status slq_select_phase(thy_slq_s* h, int Kr, cudaStream_t s, bool tracing) {
  const int d = h->d;
  if (tracing) cudaEventRecord(h->tev[0], s);
  const int nslab = 1184;   // 8 slabs per SM
  k_slq_sigma_part<<<nslab, 256, 0, s>>>(h->Pl, d, h->x.ptr, h->sig_part.ptr, nslab);
  k_slq_sigma_final<<<(unsigned)ceil_div(d, 8), 256, 0, s>>>(h->Pl, d, h->sig_part.ptr, nslab, h->sigma.ptr);
  if (tracing) cudaEventRecord(h->tev[1], s);
  THY_TRY(slq_sort_local(h));
  if (tracing) cudaEventRecord(h->tev[2], s);
  // each rank contributes its Kr smallest; the Kr globally smallest are retired
  const double* merged = h->keys_out.ptr;
  if (h->world > 1) {
    THY_TRY(comm_all_gather_f64(h->comm, h->keys_out.ptr, h->cand_all.ptr, (size_t)Kr, s));
    if (tracing) cudaEventRecord(h->tev[3], s);
    k_slq_merge<<<(unsigned)ceil_div((long long)Kr * h->world, 256), 256, 0, s>>>(h->cand_all.ptr, Kr, h->world,
                                                                                 h->cand_sorted.ptr);
    merged = h->cand_sorted.ptr;
  } else if (tracing) {
    cudaEventRecord(h->tev[3], s);
  }
  k_slq_threshold<<<1, 32, 0, s>>>(merged, Kr, h->keys_out.ptr, h->cand_all.ptr, h->world, h->rank, h->scalars.ptr,
                                   h->count_dev.ptr);
  THY_CUDA_CHECK(cudaMemcpyAsync(h->host_c, h->count_dev.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
  k_slq_copy_retired<<<(unsigned)ceil_div(Kr, 256), 256, 0, s>>>(merged, Kr, h->ctl.ptr, h->ret_ll.ptr);
  if (tracing) cudaEventRecord(h->tev[4], s);
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

static thy_status slq_round(thy_slq_s* h, int Kr, int* out_local_count) {
  thy_model_s* m = h->model;
  cudaStream_t s = m->stream;
  const int d = h->d;
  if (h->trace && h->rounds > 0) {   // the previous round's events have all completed or will by the sync below
    cudaEventSynchronize(h->tev[5]);
    for (int i = 0; i < 5; ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, h->tev[i], h->tev[i + 1]) == cudaSuccess) h->tsum[i] += ms;
    }
    h->tcount += 1;
  }
  if (h->iterations + Kr > h->ret_capacity) return fail(THY_ERR_STATE, "SLQ: retired buffer exhausted");
  const auto ht0 = std::chrono::steady_clock::now();
  *h->host_it = h->iterations;   // the previous round's copy of this word completed before its sync returned
  THY_CUDA_CHECK(cudaMemcpyAsync(h->ctl.ptr, h->host_it, sizeof(long long), cudaMemcpyHostToDevice, s));
  const bool graphable = h->use_graph && !h->trace && s != 0 && Kr == h->K && h->rounds > 0;   // round 0 sizes the sort scratch
  if (h->select_graph && h->select_graph_stream != s) {   // the model was moved to another stream: re-capture
    cudaGraphExecDestroy(h->select_graph);
    h->select_graph = nullptr;
  }
  if (graphable && !h->select_graph) {
    h->select_graph_stream = s;
    cudaGraph_t g = nullptr;
    THY_CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    thy_status st = slq_select_phase(h, Kr, s, false);
    cudaError_t ce = cudaStreamEndCapture(s, &g);
    if (st != THY_OK || ce != cudaSuccess || !g) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      h->use_graph = false;   // capture unsupported here (e.g. a collective library without graph support): plain launches
    } else {
      ce = cudaGraphInstantiate(&h->select_graph, g, 0);
      cudaGraphDestroy(g);
      if (ce != cudaSuccess) {
        cudaGetLastError();
        h->select_graph = nullptr;
        h->use_graph = false;
      }
    }
  }
  const int select_launches = 2 + 11 + (h->world > 1 ? 2 : 0) + 2;
  if (graphable && h->select_graph) {
    THY_CUDA_CHECK(cudaGraphLaunch(h->select_graph, s));
  } else {
    THY_TRY(slq_select_phase(h, Kr, s, h->trace));
  }
  count_launch(select_launches);
  // the evidence update only needs the retired values: it runs on the side stream while the walk replaces the points
  THY_CUDA_CHECK(cudaEventRecord(h->ev_retired, s));
  THY_CUDA_CHECK(cudaStreamWaitEvent(h->side, h->ev_retired, 0));
  THY_TRY(slq_evidence(h, h->ret_ll.ptr + h->iterations, Kr, h->P, 0, h->side));
  THY_CUDA_CHECK(cudaMemcpyAsync(h->host_ao, h->abs_out.ptr, (size_t)h->A * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->side));
  THY_CUDA_CHECK(cudaEventRecord(h->ev_evidence, h->side));
  const auto ht1 = std::chrono::steady_clock::now();
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  const auto ht2 = std::chrono::steady_clock::now();
  const int c = *h->host_c;
  if (c > h->K) return fail(THY_ERR_STATE, "SLQ: internal error, retire count exceeds the walker buffers");
  if (c >= h->Pl) return fail(THY_ERR_STATE, "SLQ: a rank lost all of its live points in one round; lower sampleParallelism");
  if (h->cfg.keepRetiredPoints && h->world == 1 && c > 0) {
    k_slq_store<<<(unsigned)ceil_div(c, 8), 256, 0, s>>>(c, d, h->idx_out.ptr, h->x.ptr, h->ret_x.ptr + (size_t)h->iterations * d);
    count_launch();
  }
  if (c > 0) {
    if (h->fused_walk) {
      THY_TRY(launch_slq_walk(m, c, h->Pl, h->M, h->seed, (uint64_t)h->rounds, h->rank, h->idx_out.ptr, h->x.ptr, h->ll.ptr,
                              h->lpr.ptr, h->sigma.ptr, h->scalars.ptr, h->accepted.ptr, s));
      h->evals += (long long)c * h->M;
    } else {
    k_slq_spawn<<<(unsigned)ceil_div(c, 8), 256, 0, s>>>(c, h->Pl, d, h->seed, (uint64_t)h->rounds, h->rank, h->idx_out.ptr,
                                                        h->x.ptr, h->ll.ptr, h->lpr.ptr, h->Y.ptr, h->yl.ptr, h->ypr.ptr,
                                                        h->slot.ptr);
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
    k_slq_adapt<<<1, 1024, 0, s>>>(h->scalars.ptr, h->cfg.domainScalingIncrement, h->cfg.targetAcceptanceProbability, c, h->M,
                                   h->accepted.ptr);
    count_launch();
  }
  if (h->trace) cudaEventRecord(h->tev[5], s);
  if (h->htrace) {
    const auto ht3 = std::chrono::steady_clock::now();
    h->hsum[0] += std::chrono::duration<double>(ht1 - ht0).count();
    h->hsum[1] += std::chrono::duration<double>(ht2 - ht1).count();
    h->hsum[2] += std::chrono::duration<double>(ht3 - ht2).count();
    h->hcount += 1;
  }
  THY_CUDA_CHECK(cudaGetLastError());
  h->iterations += Kr;
  h->rounds += 1;
  if (out_local_count) *out_local_count = c;
  return THY_OK;
}

}  // namespace thy

using namespace thy;

extern "C" {

thy_status thy_slq_create(thy_model_t m, const thy_slq_config* cfg, uint64_t rng_seed, const double* seeds, int n_seeds,
                          thy_comm_t comm, thy_slq_t* out) {
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
  // every rank contributes its K smallest and must keep survivors to spawn from: K <= half of a rank's pool (a rank
  // holds ~K / world of a round's retirees; one that would lose its whole pool is reported as an error)
  K = (int)std::min<long long>(K, std::max<long long>(1, h->Pl / 2));
  h->K = K;
  h->ret_capacity = (long long)cfg->maxIterationCount + K + h->P;
  const int d = h->d;
  const long long Pl = h->Pl;
  thy_status st = THY_OK;
  auto chk = [&](thy_status s2) { if (st == THY_OK) st = s2; };
  chk(h->x.reserve((size_t)Pl * d)); chk(h->ll.reserve(Pl)); chk(h->lpr.reserve(Pl));
  chk(h->keys_in.reserve(Pl)); chk(h->keys_out.reserve(Pl)); chk(h->idx_in.reserve(Pl)); chk(h->idx_out.reserve(Pl));
  chk(h->cand_all.reserve((size_t)std::max<long long>((long long)K * world, h->P)));
  chk(h->cand_sorted.reserve((size_t)std::max<long long>((long long)K * world, h->P)));
  chk(h->Y.reserve((size_t)K * d)); chk(h->Yp.reserve((size_t)K * d)); chk(h->yl.reserve(K)); chk(h->ypr.reserve(K));
  chk(h->tl.reserve(K)); chk(h->tpr.reserve(K)); chk(h->slot.reserve(K)); chk(h->accepted.reserve(K));
  chk(h->sigma.reserve(d)); chk(h->sig_part.reserve((size_t)1184 * 2 * d));
  chk(h->scalars.reserve(8)); chk(h->count_dev.reserve(1));
  chk(h->tmp_logx.reserve((size_t)h->A * std::max<long long>(K, h->P)));
  chk(h->abs_state.reserve((size_t)h->A * 8)); chk(h->abs_out.reserve((size_t)h->A * 2));
  chk(h->ret_ll.reserve((size_t)h->ret_capacity));
  if (cfg->keepRetiredPoints) {
    if (world != 1) chk(fail(THY_ERR_UNSUPPORTED, "keepRetiredPoints is single-GPU only in this version"));
    chk(h->ret_x.reserve((size_t)h->ret_capacity * d));
  }
  cudaStream_t s = m->stream;
  if (st == THY_OK) {
    if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_retired, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_evidence, cudaEventDisableTiming) != cudaSuccess ||
        cudaMallocHost(&h->host_ao, (size_t)h->A * 2 * sizeof(double)) != cudaSuccess ||
        cudaMallocHost(&h->host_c, sizeof(int)) != cudaSuccess ||
        cudaMallocHost(&h->host_it, sizeof(long long)) != cudaSuccess ||
        cudaEventRecord(h->ev_evidence, h->side) != cudaSuccess)
      st = fail(THY_ERR_CUDA, "SLQ side stream setup failed");
    h->fused_walk = !cfg->disableFusedWalk && slq_walk_supported(m);
    chk(h->ctl.reserve(1));
    const char* ng = std::getenv("THY_SLQ_GRAPH");
    h->use_graph = !(ng && ng[0] == '0');
    const char* tr = std::getenv("THY_SLQ_TRACE");
    h->trace = tr && tr[0] == '1';
    h->htrace = tr && tr[0] == '2';
    if (h->trace)
      for (auto& e : h->tev)
        if (cudaEventCreate(&e) != cudaSuccess) st = fail(THY_ERR_CUDA, "SLQ trace events");
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
    k_iota<<<(unsigned)ceil_div(Pl, 256), 256, 0, s>>>((int)Pl, h->idx_in.ptr);
    std::vector<double> sc = {0.5, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (st == THY_OK && cudaMemsetAsync(h->accepted.ptr, 0, (size_t)K * sizeof(int), s) != cudaSuccess)
      st = fail(THY_ERR_CUDA, "SLQ initialisation failed");
    chk(h->scalars.upload(sc, s));
    std::vector<double> stt((size_t)h->A * 8, 0.0);
    for (int a = 0; a < h->A; ++a) stt[(size_t)a * 8 + 5] = -INFINITY;
    chk(h->abs_state.upload(stt, s));
    count_launch();
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
  delete h;
  return THY_OK;
}

// Live phase: rounds of "retire the K lowest, replace them above the threshold" until the reference's convergence test
// (SlqEngine.scala:249-262: iterations > minIterationCount && iterations >= 10 P H) or maxIterationCount.
static thy_status slq_live_phase(thy_slq_s* h, long long max_rounds, bool* converged) {
  cudaStream_t s = h->model->stream;
  std::vector<double> ao((size_t)h->A * 2);
  *converged = false;
  for (long long r = 0; r < max_rounds; ++r) {
    const long long remaining = (long long)h->cfg.maxIterationCount - h->iterations;
    if (remaining <= 0) { *converged = true; break; }
    const int Kr = (int)std::min<long long>(h->K, remaining);
    int c = 0;
    THY_TRY(slq_round(h, Kr, &c));
    // the side stream finishes the evidence update long before the walk ends: the host decides on convergence and
    // queues the next round's sort while the walk is still running
    const auto hw0 = std::chrono::steady_clock::now();
    THY_CUDA_CHECK(cudaEventSynchronize(h->ev_evidence));
    if (h->htrace) h->hsum[3] += std::chrono::duration<double>(std::chrono::steady_clock::now() - hw0).count();
    double hmax = -INFINITY;
    for (int a = 0; a < h->A; ++a) hmax = std::max(hmax, h->host_ao[(size_t)a * 2 + 1]);
    if (h->iterations > h->cfg.minIterationCount && std::isfinite(hmax) &&
        (double)h->iterations >= 10.0 * (double)h->P * hmax) {
      *converged = true;
      break;
    }
  }
  return THY_OK;
}

extern "C" thy_status thy_slq_run_rounds(thy_slq_t h, int n_rounds, int* out_converged) {
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (h->finished) return fail(THY_ERR_STATE, "the simulation has already been drained");
  bool conv = false;
  THY_TRY(slq_live_phase(h, n_rounds, &conv));
  if (out_converged) *out_converged = conv ? 1 : 0;
  return THY_OK;
}

extern "C" thy_status thy_slq_run(thy_slq_t h, thy_slq_stats* out_stats) {
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (h->finished) {
    if (out_stats) *out_stats = h->stats;
    return THY_OK;
  }
  cudaStream_t s = h->model->stream;
  std::vector<double> ao((size_t)h->A * 2);
  bool conv = false;
  THY_TRY(slq_live_phase(h, (long long)1 << 60, &conv));
  // drainSamplePool (SlqEngine.scala:171-180): the remaining live points in ascending order
  THY_TRY(slq_sort_local(h));
  const double* merged = h->keys_out.ptr;
  if (h->world > 1) {
    THY_TRY(comm_all_gather_f64(h->comm, h->keys_out.ptr, h->cand_all.ptr, (size_t)h->Pl, s));
    THY_TRY(slq_sort_keys(h, h->cand_all.ptr, h->cand_sorted.ptr, (int)h->P));
    merged = h->cand_sorted.ptr;
  }
  if (h->iterations + h->P > h->ret_capacity) return fail(THY_ERR_STATE, "SLQ: retired buffer exhausted");
  THY_CUDA_CHECK(cudaMemcpyAsync(h->ret_ll.ptr + h->iterations, merged, (size_t)h->P * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if (h->cfg.keepRetiredPoints && h->world == 1) {
    k_slq_store<<<(unsigned)ceil_div(h->Pl, 8), 256, 0, s>>>((int)h->Pl, h->d, h->idx_out.ptr, h->x.ptr,
                                                            h->ret_x.ptr + (size_t)h->iterations * h->d);
    count_launch();
  }
  THY_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_evidence, 0));   // the last round's evidence update (side stream)
  THY_TRY(slq_evidence(h, h->ret_ll.ptr + h->iterations, (int)h->P, h->P, 1, s));
  const long long live_iterations = h->iterations;
  h->iterations += h->P;
  std::vector<double> sc(8);
  THY_CUDA_CHECK(cudaMemcpyAsync(ao.data(), h->abs_out.ptr, ao.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
  THY_CUDA_CHECK(cudaMemcpyAsync(sc.data(), h->scalars.ptr, 8 * sizeof(double), cudaMemcpyDeviceToHost, s));
  double minl = 0.0;
  THY_CUDA_CHECK(cudaMemcpyAsync(&minl, h->ret_ll.ptr + live_iterations, sizeof(double), cudaMemcpyDeviceToHost, s));
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  double mean = 0.0, m2 = 0.0, hm = 0.0;
  for (int a = 0; a < h->A; ++a) {
    mean += ao[(size_t)a * 2];
    hm += ao[(size_t)a * 2 + 1];
  }
  mean /= h->A;
  hm /= h->A;
  for (int a = 0; a < h->A; ++a) m2 += (ao[(size_t)a * 2] - mean) * (ao[(size_t)a * 2] - mean);
  h->stats.iterations = live_iterations;
  h->stats.total_retired = h->iterations;
  h->stats.rounds = h->rounds;
  h->stats.log_evidence_mean = mean;
  h->stats.log_evidence_std = h->A > 1 ? std::sqrt(m2 / (h->A - 1)) : 0.0;
  h->stats.neg_entropy_mean = hm;
  h->stats.min_logpdf = minl;
  h->stats.walk_scale = sc[0];
  h->stats.walk_acceptance = sc[6] > 0 ? sc[5] / sc[6] : 0.0;
  h->stats.likelihood_evaluations = h->evals;
  h->finished = true;
  if (out_stats) *out_stats = h->stats;
  return THY_OK;
}

/* Retired log-likelihoods in retirement order (live phase followed by the drained pool). */
thy_status thy_slq_retired(thy_slq_t h, int64_t max_count, double* out_loglik, int64_t* out_count) {
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  const long long n = std::min<long long>(max_count, h->iterations);
  if (out_loglik && n > 0) {
    THY_CUDA_CHECK(cudaStreamSynchronize(h->model->stream));
    THY_CUDA_CHECK(cudaMemcpy(out_loglik, h->ret_ll.ptr, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (out_count) *out_count = h->iterations;
  return THY_OK;
}

/* The rank's live points and their log-likelihoods (the reference's `samples` map, SlqEngine.scala:120-143). */
thy_status thy_slq_live_points(thy_slq_t h, int64_t* out_count, double* out_x, double* out_loglik) {
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (out_count) *out_count = h->Pl;
  THY_CUDA_CHECK(cudaStreamSynchronize(h->model->stream));
  if (out_x) THY_CUDA_CHECK(cudaMemcpy(out_x, h->x.ptr, (size_t)h->Pl * h->d * sizeof(double), cudaMemcpyDeviceToHost));
  if (out_loglik) THY_CUDA_CHECK(cudaMemcpy(out_loglik, h->ll.ptr, (size_t)h->Pl * sizeof(double), cudaMemcpyDeviceToHost));
  return THY_OK;
}

/* Per-abscissa results of the quadrature so far: ln Z_a and H_a (QuadratureIntegrator.scala:29-62). */
thy_status thy_slq_abscissa_results(thy_slq_t h, double* out_log_evidence, double* out_neg_entropy) {
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

/* Posterior resampling of retired points (SamplingSimulation.scala:41-101) with weight w_i L_i under the expected
 * abscissa E[ln X_i] = -sum 1/n_k.  Requires keepRetiredPoints (single GPU). */
thy_status thy_slq_sample(thy_slq_t h, int n, uint64_t rng_seed, double* out_samples) {
  if (!h) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (!h->finished) return fail(THY_ERR_STATE, "run the simulation first (rebuildSampleSimulation)");
  if (!h->cfg.keepRetiredPoints || h->world != 1) return fail(THY_ERR_UNSUPPORTED, "retired points were not kept");
  if (n <= 0 || !out_samples) return fail(THY_ERR_INVALID_ARGUMENT, "bad arguments");
  const long long N = h->iterations;
  std::vector<double> l((size_t)N);
  THY_CUDA_CHECK(cudaMemcpy(l.data(), h->ret_ll.ptr, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost));
  // expected log X: live phase shrinks by 1/(P - j) within rounds of K, drain by 1/(P - j)
  std::vector<double> lx((size_t)N);
  double cur = 0.0;
  const long long live = h->stats.iterations;
  for (long long i = 0; i < N; ++i) {
    lx[(size_t)i] = cur;
    const long long j = (i < live) ? (i % h->K) : (i - live);
    cur -= 1.0 / (double)(h->P - j);
  }
  std::vector<double> lw((size_t)N);
  double mx = -INFINITY;
  for (long long i = 0; i < N; ++i) {
    const double hi = (i == 0) ? lx[0] : lx[(size_t)i - 1];
    const double lo = (i + 1 < N) ? lx[(size_t)i + 1] : lx[(size_t)i];
    const double gap = std::log(0.5) + hi + std::log(-std::expm1(lo - hi));
    lw[(size_t)i] = l[(size_t)i] + gap;
    mx = std::max(mx, lw[(size_t)i]);
  }
  std::vector<double> cdf((size_t)N);
  double run = 0.0;
  for (long long i = 0; i < N; ++i) {
    run += std::exp(lw[(size_t)i] - mx);
    cdf[(size_t)i] = run;
  }
  for (int k = 0; k < n; ++k) {
    const RngDraw u = rng_uniform2(rng_seed, (uint64_t)k, 0, RNG_SLQ_PICK, 0xFFFFFFFFull);
    const double target = u.u0 * run;
    const long long idx = std::min<long long>(N - 1, std::upper_bound(cdf.begin(), cdf.end(), target) - cdf.begin());
    THY_CUDA_CHECK(cudaMemcpy(out_samples + (size_t)k * h->d, h->ret_x.ptr + (size_t)idx * h->d, h->d * sizeof(double),
                              cudaMemcpyDeviceToHost));
  }
  return THY_OK;
}

}  // extern "C"
