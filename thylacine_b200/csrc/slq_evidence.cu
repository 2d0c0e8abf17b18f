// SLQ prior-mass simulation and quadrature on the device, spread over the whole GPU.
// Reference: /root/reference/src/main/scala/thylacine/model/integration/slq/
//   QuadratureAbscissa.scala:33-73    prior-mass simulation: the retired point leaves with the largest of n_live uniforms,
//                                     i.e. X shrinks by t = u^{1/n_live}; first abscissa = 1.0; trapezoid weights
//                                     1/2 (X_{i-1} - X_{i+1}), the two ends half the adjacent gap
//   QuadratureIntegrator.scala:29-62  Z_a = sum w_i exp(l_i - s),  H_a = sum w_i e^{l_i - s} (l_i - s) / Z_a - ln Z_a
// Round 1 ran ONE CTA per abscissa (250 us per round at 62 500 retirements: 8 SMs busy, 140 idle).  Here the round's
// retirements are cut into segments of 2048; k_ev_shrink draws ln t_j for every (abscissa, point) and leaves per-segment
// sums, k_ev_terms rebuilds ln X_j from the segment prefix, accumulates its segment's log-sum-exp terms, and the last
// CTA of an abscissa (atomic ticket) folds the segment partials, in segment order, into the carried state.
// Deterministic: every sum has a fixed order.  The trapezoid gap is evaluated from the shrink factors themselves,
//   ln 1/2 (X_{j-2} - X_j) = ln 1/2 + ln X_{j-2} + ln(1 - t_{j-2} t_{j-1}),
// not from the difference of two accumulated ln X (which cancels ~6 digits at 10^6 live points).
#include "slq_internal.h"
#include "philox.cuh"

namespace thy {

namespace {

constexpr int EVT = 256;          // threads
constexpr int EVI = 8;            // consecutive points per thread
constexpr int EVSEG = EVT * EVI;  // points per segment

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
__device__ __forceinline__ void lse_merge(double& m1, double& s1, double& sl1, double m2, double s2, double sl2) {
  if (!(m2 > -INFINITY)) return;
  if (m2 > m1) {
    const double f = exp(m1 - m2);
    s1 = s1 * f + s2;
    sl1 = sl1 * f + sl2;
    m1 = m2;
  } else {
    const double f = exp(m2 - m1);
    s1 += s2 * f;
    sl1 += sl2 * f;
  }
}
// ln( 1/2 X (1 - exp(gapsum)) ), gapsum = sum of the ln t between the two abscissas (< 0)
__device__ __forceinline__ double log_half_gap(double lx_hi, double gapsum) {
  return -0.6931471805599453 + lx_hi + log(-expm1(gapsum));
}

// Fixed-order block sum (per-thread value -> total in every thread).
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[wrp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < EVT / 32; ++w) t += sh[w];
  return t;
}
// Exclusive block scan of per-thread values.
__device__ __forceinline__ double block_excl(double v, double* sh) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  double incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  __syncthreads();
  if (lane == 31) sh[wrp] = incl;
  __syncthreads();
  double base = 0.0;
  for (int w = 0; w < wrp; ++w) base += sh[w];
  return base + (incl - v);
}

// ln t_j = ln(u) / n_live for every point of the batch, per abscissa; per-segment sums.
__global__ void __launch_bounds__(EVT) k_ev_shrink(int a0, int n, long long P, long long K, long long live, long long it0v,
                                                   const long long* __restrict__ it0_dev, uint64_t seed, double* __restrict__ lt, long long stride,
                                                   double* __restrict__ segsum, int nseg_max) {
  __shared__ double sh[EVT / 32];
  const int a = a0 + blockIdx.y, seg = blockIdx.x;
  const long long it0 = it0_dev ? *it0_dev : it0v;
  double* row = lt + (size_t)blockIdx.y * stride;
  const int j0 = seg * EVSEG + threadIdx.x * EVI;
  double local = 0.0;
#pragma unroll
  for (int q = 0; q < EVI; ++q) {
    const int j = j0 + q;
    if (j < n) {
      const long long i = it0 + j;
      const RngDraw u = rng_uniform2(seed, (uint64_t)a, 0, RNG_SLQ_SHRINK, (uint64_t)i);
      const double v = log(1.0 - u.u0) / (double)slq_n_live(i, P, K, live);   // u in (0, 1]
      row[j] = v;
      local += v;
    }
  }
  const double tot = block_sum(local, sh);
  if (threadIdx.x == 0) segsum[(size_t)blockIdx.y * nseg_max + seg] = tot;
}

__device__ __forceinline__ double segment_base(const double* seg_row, int seg, double start, double* sh) {
  // start + sum of the earlier segments' totals (fixed order: lane-strided partials, shuffle tree)
  double v = 0.0;
  if (threadIdx.x < 32)
    for (int q = threadIdx.x; q < seg; q += 32) v += seg_row[q];
  return start + block_sum(threadIdx.x < 32 ? v : 0.0, sh);
}

__global__ void __launch_bounds__(EVT) k_ev_terms(int Kr, const double* __restrict__ r, const double* __restrict__ lt,
                                                  long long stride, const double* __restrict__ segsum, int nseg_max,
                                                  double* __restrict__ part, unsigned int* __restrict__ tickets,
                                                  double* __restrict__ state, int finalize, double* __restrict__ out) {
  __shared__ double sh[EVT / 32];
  __shared__ double red_m[EVT], red_s[EVT], red_sl[EVT];
  __shared__ int s_last;
  const int a = blockIdx.y, seg = blockIdx.x, nseg = gridDim.x;
  double* st = state + (size_t)a * EV_STATE;
  const double* row = lt + (size_t)a * stride;
  const double* seg_row = segsum + (size_t)a * nseg_max;
  const double st0 = st[0], st1 = st[1], st2 = st[2], st3 = st[3], pend = st[4], st8 = st[8], st9 = st[9];
  const double base = segment_base(seg_row, seg, st0, sh);
  const int j0 = seg * EVSEG + threadIdx.x * EVI;
  double v[EVI], local = 0.0;
#pragma unroll
  for (int q = 0; q < EVI; ++q) {
    v[q] = (j0 + q < Kr) ? row[j0 + q] : 0.0;
    local += v[q];
  }
  double lx = base + block_excl(local, sh);   // ln X of point j0
  double m = -INFINITY, s = 0.0, sl = 0.0;
#pragma unroll
  for (int q = 0; q < EVI; ++q) {
    const int j = j0 + q;
    if (j < Kr) {
      // X_j is now known: finalise point j - 1 with w = 1/2 (X_{j-2} - X_j)   (first ever point: 1/2 (X_0 - X_1))
      bool have = true;
      double l_prev, lx_m2, gap;
      if (j >= 2) {
        const double t1 = (q >= 1) ? v[q - 1] : row[j - 1];
        const double t2 = (q >= 2) ? v[q - 2] : row[j - 2];
        l_prev = r[j - 1];
        gap = t2 + t1;
        lx_m2 = lx - t1 - t2;
      } else if (j == 1) {
        l_prev = r[0];
        if (pend < 0.5) {          // r[0] is the very first retired point
          gap = row[0];
          lx_m2 = lx - row[0];
        } else {
          gap = st9 + row[0];
          lx_m2 = st2;
        }
      } else {
        have = pend > 0.5;
        l_prev = st3;
        if (pend > 1.5) {
          gap = st8 + st9;
          lx_m2 = st1;
        } else {                   // the pending point is the very first retired point
          gap = st9;
          lx_m2 = st2;
        }
      }
      if (have) lse_add(m, s, sl, l_prev + log_half_gap(lx_m2, gap), l_prev);
      lx += v[q];
    }
  }
  red_m[threadIdx.x] = m;
  red_s[threadIdx.x] = s;
  red_sl[threadIdx.x] = sl;
  __syncthreads();
  for (int o = EVT / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      double m1 = red_m[threadIdx.x], s1 = red_s[threadIdx.x], sl1 = red_sl[threadIdx.x];
      lse_merge(m1, s1, sl1, red_m[threadIdx.x + o], red_s[threadIdx.x + o], red_sl[threadIdx.x + o]);
      red_m[threadIdx.x] = m1;
      red_s[threadIdx.x] = s1;
      red_sl[threadIdx.x] = sl1;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double* pp = part + ((size_t)a * nseg_max + seg) * 3;
    pp[0] = red_m[0];
    pp[1] = red_s[0];
    pp[2] = red_sl[0];
    __threadfence();
    s_last = (atomicAdd(&tickets[a], 1u) == (unsigned int)(nseg - 1)) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  // ---- last CTA of this abscissa: fold the segment partials (segment order) into the carried state ----
  double M = st[5], S = st[6], SL = st[7];
  double total = st0;
  for (int q = 0; q < nseg; ++q) {
    const double* pp = part + ((size_t)a * nseg_max + q) * 3;
    lse_merge(M, S, SL, __ldcg(pp), __ldcg(pp + 1), __ldcg(pp + 2));
    total += __ldcg(&seg_row[q]);
  }
  if (Kr > 0) {
    const double tl1 = row[Kr - 1];                       // ln t of the round's last point
    const double tl2 = (Kr >= 2) ? row[Kr - 2] : st9;
    const double lx_m1 = total - tl1;                     // ln X of the round's last point
    st[1] = (Kr >= 2) ? (lx_m1 - tl2) : st2;
    st[2] = lx_m1;
    st[3] = r[Kr - 1];
    const double np = (Kr >= 2) ? 2.0 : (pend + 1.0);
    st[4] = np > 2.0 ? 2.0 : np;
    st[8] = tl2;
    st[9] = tl1;
    st[0] = total;
  }
  if (finalize && st[4] > 1.5) {
    // last point of the whole simulation: 1/2 (X_{N-2} - X_{N-1})
    lse_add(M, S, SL, st[3] + log_half_gap(st[1], st[8]), st[3]);
    st[4] = 0.0;
  }
  st[5] = M;
  st[6] = S;
  st[7] = SL;
  const double logZ = (S > 0.0) ? M + log(S) : -INFINITY;
  out[a * 2 + 0] = logZ;
  out[a * 2 + 1] = (S > 0.0) ? SL / S - logZ : 0.0;
  tickets[a] = 0u;
}

// ln X_j for one abscissa over a batch (export for thy_slq_retired_log_x / integrate / sample).
__global__ void __launch_bounds__(EVT) k_ev_logx(int n, const double* __restrict__ lt, const double* __restrict__ segsum,
                                                 const double* __restrict__ carry, double* __restrict__ out) {
  __shared__ double sh[EVT / 32];
  const int seg = blockIdx.x;
  const double base = segment_base(segsum, seg, *carry, sh);
  const int j0 = seg * EVSEG + threadIdx.x * EVI;
  double v[EVI], local = 0.0;
#pragma unroll
  for (int q = 0; q < EVI; ++q) {
    v[q] = (j0 + q < n) ? lt[j0 + q] : 0.0;
    local += v[q];
  }
  double lx = base + block_excl(local, sh);
#pragma unroll
  for (int q = 0; q < EVI; ++q) {
    if (j0 + q < n) out[j0 + q] = lx;
    lx += v[q];
  }
}
__global__ void k_ev_carry(int nseg, const double* __restrict__ segsum, double* __restrict__ carry) {
  // single warp, fixed order
  double v = 0.0;
  for (int q = threadIdx.x; q < nseg; q += 32) v += segsum[q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (threadIdx.x == 0) *carry += v;
}

}  // namespace

thy_status slq_evidence_launch(const EvidenceBuffers& e, int A, int Kr, long long P, long long K, long long live,
                               long long it0, const long long* it0_dev, uint64_t seed, const double* r, int finalize,
                               cudaStream_t s) {
  const int nseg = Kr > 0 ? (Kr + EVSEG - 1) / EVSEG : 1;
  if (nseg > e.nseg_max || Kr > e.stride) return fail(THY_ERR_STATE, "SLQ evidence: batch exceeds the scratch buffers");
  const dim3 grid((unsigned)nseg, (unsigned)A);
  k_ev_shrink<<<grid, EVT, 0, s>>>(0, Kr, P, K, live, it0, it0_dev, seed, e.lt, e.stride, e.segsum, e.nseg_max);
  k_ev_terms<<<grid, EVT, 0, s>>>(Kr, r, e.lt, e.stride, e.segsum, e.nseg_max, e.part, e.tickets, e.state, finalize, e.out);
  count_launch(2);
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

thy_status slq_logx_launch(const EvidenceBuffers& e, int a, long long i0, int n, long long P, long long K, long long live,
                           uint64_t seed, double* carry_dev, double* out_dev, cudaStream_t s) {
  if (n <= 0) return THY_OK;
  const int nseg = (n + EVSEG - 1) / EVSEG;
  if (nseg > e.nseg_max || n > e.stride) return fail(THY_ERR_STATE, "SLQ log X: batch exceeds the scratch buffers");
  k_ev_shrink<<<dim3((unsigned)nseg, 1), EVT, 0, s>>>(a, n, P, K, live, i0, nullptr, seed, e.lt, e.stride, e.segsum, e.nseg_max);
  k_ev_logx<<<nseg, EVT, 0, s>>>(n, e.lt, e.segsum, carry_dev, out_dev);
  k_ev_carry<<<1, 32, 0, s>>>(nseg, e.segsum, carry_dev);
  count_launch(3);
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

}  // namespace thy
