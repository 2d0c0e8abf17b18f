// Small-batch (B <= 8) linear forward / weighted residual / adjoint: a TMA-staged, memory-bound streaming pass.
//
// With a handful of chains the contraction is a few GEMVs, not a dense GEMM: the tensor cores have nothing to chew on
// and the cost is reading A once -- 8 n d bytes per launch ("small B").  Same arithmetic as the fused DMMA
// kernel (reference: model/components/likelihood/Likelihood.scala:43-65, forwardmodel/LinearForwardModel.scala:103-114,
// distributions/GaussianDistribution.scala:52-68, CauchyDistribution.scala:51-79):
//     f = A x_b (+ offset);  r = y - f;  w = r / var;  q_b = sum_i r_i w_i;  G_b = A^T w
// Structure:
//   * persistent grid, one CTA per SM, each owning a contiguous range of 32-row blocks of A;
//   * a producer warp streams the blocks (A rows + (y, 1/var)) through an mbarrier ring with 1-D TMA bulk copies
//     (cp.async.bulk -> UBLKCP), deep enough to cover HBM latency with one CTA per SM;
//   * 8 consumer warps take the rows of a block round-robin; lanes stride the columns (coalesced, conflict-free
//     shared-memory reads), the chain vectors x_b and the gradient accumulators live in registers;
//   * the row dot products are completed with warp shuffles (recursive halving: 17 shuffles for 8 chains instead of 40);
//   * per-CTA partials are combined by the LAST CTA to finish (atomic ticket), in CTA order => deterministic, and the
//     whole likelihood term is ONE kernel launch.
#include "kernels.h"
#include "ptx.cuh"

namespace thy {

namespace {

constexpr int SNW = 8;   // consumer warps
constexpr int SMB = 32;  // rows per stage (matches the fused kernel's A layout: n_pad % 32 == 0)

struct StreamParams {
  const double* A;      // [n_pad][lds]
  const double2* yiv;   // [n_pad] (y - offset, 1/var)
  const int* col;       // [dl]
  const double* X;      // [B][d]
  double* logp;         // [B]
  double* grad;         // [B][d] or null
  double* partG;        // [G][BT][32 T]
  double* partL;        // [G][BT]
  unsigned int* ticket; // [1], zero between launches
  int B, d, dl, np, lds, n, nblk, stages;
  int distribution;
  double log_const, cauchy_sign;
};

__device__ __forceinline__ double obs_logp(int distribution, double q, int n, double log_const) {
  if (distribution == THY_DIST_GAUSSIAN) return -0.5 * q + log_const;
  return log_const - (1.0 + n) / 2.0 * log(1.0 + q);
}
__device__ __forceinline__ double obs_scale(int distribution, double q, int n, double cauchy_sign) {
  if (distribution == THY_DIST_GAUSSIAN) return 1.0;
  return -cauchy_sign * (1.0 + n) / (1.0 + q);
}

// All-reduce of BT per-lane values over the warp; every lane ends up with every total.
template <int BT>
__device__ __forceinline__ void warp_allreduce_multi(double (&v)[BT], int lane) {
  if constexpr (BT <= 2) {
#pragma unroll
    for (int b = 0; b < BT; ++b) v[b] = warp_sum(v[b]);
  } else {
    // recursive halving over lane bits 4, 3 (and 2 for BT = 8): each step halves the number of values a lane carries
    {
      const bool upper = (lane & 16) != 0;
#pragma unroll
      for (int i = 0; i < BT / 2; ++i) {
        const double send = upper ? v[i] : v[i + BT / 2];
        const double keep = upper ? v[i + BT / 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
    }
    {
      const bool upper = (lane & 8) != 0;
#pragma unroll
      for (int i = 0; i < BT / 4; ++i) {
        const double send = upper ? v[i] : v[i + BT / 4];
        const double keep = upper ? v[i + BT / 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
    }
    if constexpr (BT == 8) {
      const bool upper = (lane & 4) != 0;
      const double send = upper ? v[0] : v[1];
      const double keep = upper ? v[1] : v[0];
      v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    // v[0] now holds the partial sum of chain c(lane); finish over the remaining lane bits
    constexpr int LOWBITS = (BT == 8) ? 2 : 3;
    double s = v[0];
#pragma unroll
    for (int o = 1 << (LOWBITS - 1); o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    // chain b lives in the lanes whose top bits spell b (bit 4 is the most significant index bit)
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      int src;
      if constexpr (BT == 8) src = (((b >> 2) & 1) << 4) | (((b >> 1) & 1) << 3) | ((b & 1) << 2);
      else src = (((b >> 1) & 1) << 4) | ((b & 1) << 3);
      v[b] = __shfl_sync(0xffffffffu, s, src);
    }
  }
}

// Tail shared by both streaming kernels: CTA partial = sum over warps (warp order) -> global; the LAST CTA to arrive
// (atomic ticket) sums all CTA partials in CTA order and applies them to logp / grad, then re-arms the ticket.
// red: [SNW][BT][W] gradient partials, redq: [SNW][BT] quadratic-form partials (shared memory).
__device__ __forceinline__ void stream_cta_tail(const StreamParams& p, const double* red, const double* redq, int BT, int W,
                                                int* s_last, double* s_q) {
  const int G = gridDim.x, c = blockIdx.x;
  __syncthreads();
  for (int idx = threadIdx.x; idx < BT * W; idx += blockDim.x) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < SNW; ++w) s += red[w * BT * W + idx];
    p.partG[(size_t)c * BT * W + idx] = s;
  }
  if (threadIdx.x < BT) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < SNW; ++w) s += redq[w * BT + threadIdx.x];
    p.partL[(size_t)c * BT + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(p.ticket, 1u);
    *s_last = (t == (unsigned int)(G - 1)) ? 1 : 0;
  }
  __syncthreads();
  if (!*s_last) return;
  __threadfence();
  if (threadIdx.x < BT) {
    double s = 0.0;
    for (int cc = 0; cc < G; ++cc) s += __ldcg(&p.partL[(size_t)cc * BT + threadIdx.x]);
    s_q[threadIdx.x] = s;
    if (threadIdx.x < p.B) p.logp[threadIdx.x] += obs_logp(p.distribution, s, p.n, p.log_const);
  }
  __syncthreads();
  if (p.grad) {
    for (int idx = threadIdx.x; idx < p.B * p.dl; idx += blockDim.x) {
      const int b = idx / p.dl, j = idx % p.dl;
      double s = 0.0;
      for (int cc = 0; cc < G; ++cc) s += __ldcg(&p.partG[((size_t)cc * BT + b) * W + j]);
      p.grad[(size_t)b * p.d + p.col[j]] += obs_scale(p.distribution, s_q[b], p.n, p.cauchy_sign) * s;
    }
  }
  if (threadIdx.x == 0) *p.ticket = 0u;
}

template <int BT, int T>
__global__ void __launch_bounds__((SNW + 1) * 32, 1) k_stream(const StreamParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int stage_bytes = SMB * p.lds * 8 + SMB * 16;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  double* red = reinterpret_cast<double*>(empty_bar + p.stages);   // [SNW][BT][32 T] + [SNW][BT]
  double* redq = red + SNW * BT * 32 * T;
  __shared__ int s_last;
  __shared__ double s_q[BT];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, c = blockIdx.x;
  const int blk0 = (int)(((long long)c * p.nblk) / G), blk1 = (int)(((long long)(c + 1) * p.nblk) / G);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], SNW);
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  __syncthreads();

  double g[BT][T];
  double q[BT];
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    q[b] = 0.0;
#pragma unroll
    for (int t = 0; t < T; ++t) g[b][t] = 0.0;
  }

  if (warp == SNW) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int blk = blk0; blk < blk1; ++blk) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        unsigned char* dst = smem + (size_t)stage * stage_bytes;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
        ptx::bulk_g2s(dst, p.A + (size_t)blk * SMB * p.lds, SMB * p.lds * 8, &full_bar[stage]);
        ptx::bulk_g2s(dst + SMB * p.lds * 8, p.yiv + (size_t)blk * SMB, SMB * 16, &full_bar[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    double x[BT][T];
#pragma unroll
    for (int b = 0; b < BT; ++b)
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int k = lane + 32 * t;
        x[b][t] = (b < p.B && k < p.dl) ? p.X[(size_t)b * p.d + p.col[k]] : 0.0;
      }
    int stage = 0;
    uint32_t phase = 0;
    for (int blk = blk0; blk < blk1; ++blk) {
      ptx::mbar_wait(&full_bar[stage], phase);
      const double* As = reinterpret_cast<const double*>(smem + (size_t)stage * stage_bytes);
      const double2* Ys = reinterpret_cast<const double2*>(smem + (size_t)stage * stage_bytes + SMB * p.lds * 8);
#pragma unroll(BT * T <= 16 ? 2 : 1)
      for (int rr = warp; rr < SMB; rr += SNW) {
        double a[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int k = lane + 32 * t;
          a[t] = (k < p.np) ? As[rr * p.lds + k] : 0.0;
        }
        double z[BT];
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          double s = 0.0;
#pragma unroll
          for (int t = 0; t < T; ++t) s = fma(a[t], x[b][t], s);
          z[b] = s;
        }
        warp_allreduce_multi<BT>(z, lane);
        const double2 yv = Ys[rr];
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          const double r = yv.x - z[b];
          const double w = r * yv.y;
          q[b] = fma(r, w, q[b]);
#pragma unroll
          for (int t = 0; t < T; ++t) g[b][t] = fma(w, a[t], g[b][t]);
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    // per-warp partials -> shared memory
#pragma unroll
    for (int b = 0; b < BT; ++b) {
#pragma unroll
      for (int t = 0; t < T; ++t) red[(warp * BT + b) * 32 * T + lane + 32 * t] = g[b][t];
      if (lane == 0) redq[warp * BT + b] = q[b];
    }
  }
  stream_cta_tail(p, red, redq, BT, 32 * T, &s_last, s_q);
}

// ---- 2 <= B <= 8: the same pass on the fp64 tensor cores --------------------------------------------------------
// Eight chains are exactly one DMMA.8x8x4 row group, so the per-row work drops from ~90 DFMA + a shuffle tree to
// 6.5 DMMA and the pass stays HBM-bound (3.7 flop/B at 8 chains vs a ridge of 5.5).  Same fragment scheme as the fused
// kernel (GEMM1's C fragment is GEMM2's A fragment through the row permutation pi; pitch = 4 mod 8 => conflict-free),
// but the warps of a CTA split ROW blocks instead of chain groups: item i (8 rows) of a CTA belongs to warp i % SNW.
//
// Ring protocol (round 2: STATIC STAGE OWNERSHIP).  The ring depth is a multiple of SNW, so stage s = i % stages is
// always consumed by warp s % SNW -- the same warp on every reuse of the stage.  A warp takes its items in increasing
// order, hence it reaches the wait for use u+1 of a stage only after its own wait for use u has returned, i.e. after
// phase u of the stage's full barrier has completed; and the producer re-arms the stage for use u+1 only after that
// warp's arrival on the empty barrier for use u.  The full barrier is therefore never more than one phase away from the
// phase the waiting warp names, which is exactly the condition under which mbarrier.try_wait.parity is unambiguous.
// (Round 1 handed a stage to a DIFFERENT warp on each reuse; a warp one use ahead could then pass its wait on the
// parity of use u-1 and read stale data.  That guard, stream_dmma_ring_wraps, is gone: every model length runs here.)
constexpr int SR = 8;        // rows per stage = one DMMA row group

template <int NT>
struct StreamDmmaCfg {
  static constexpr int NP = 8 * NT, KS = 2 * NT, LDS = 8 * NT + 4;
  static constexpr int A_BYTES = SR * LDS * 8, STAGE_BYTES = A_BYTES + SR * 16;
  static constexpr size_t RED_BYTES = (size_t)(SNW * 8 * NP + SNW * 8) * sizeof(double);
  // ring depth: as many stages as fit beside the reduction scratch, rounded DOWN to a multiple of SNW (static ownership)
  static constexpr int STAGES_FIT = (int)((200 * 1024 - RED_BYTES) / STAGE_BYTES);
  static constexpr int STAGES = (STAGES_FIT / SNW) * SNW > 4 * SNW ? 4 * SNW : (STAGES_FIT / SNW) * SNW;
  static_assert(STAGES >= SNW && STAGES % SNW == 0, "the row-group ring needs at least one stage per consumer warp");
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 2 * STAGES * sizeof(uint64_t) + RED_BYTES;
};

template <int NT>
__global__ void __launch_bounds__((SNW + 1) * 32, 1) k_stream_dmma(const StreamParams p) {
  using C = StreamDmmaCfg<NT>;
  constexpr int NP = C::NP, KS = C::KS, LDS = C::LDS, A_BYTES = C::A_BYTES, STAGE_BYTES = C::STAGE_BYTES;
  constexpr int STAGES = C::STAGES;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  double* red = reinterpret_cast<double*>(empty_bar + STAGES);   // [SNW][8][NP]
  double* redq = red + SNW * 8 * NP;
  __shared__ int s_last;
  __shared__ double s_q[8];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, c = blockIdx.x;
  const int blk0 = (int)(((long long)c * p.nblk) / G), blk1 = (int)(((long long)(c + 1) * p.nblk) / G);
  const int nb = blk1 - blk0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);   // a stage is consumed by exactly one warp, always the same one
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  __syncthreads();

  if (warp == SNW) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nb; ++i) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        unsigned char* dst = smem + (size_t)stage * STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
        ptx::bulk_g2s(dst, p.A + (size_t)(blk0 + i) * SR * LDS, A_BYTES, &full_bar[stage]);
        ptx::bulk_g2s(dst + A_BYTES, p.yiv + (size_t)(blk0 + i) * SR, SR * 16, &full_bar[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    const int g8 = lane >> 2, q4 = lane & 3;
    const int pi_g = (g8 < 4) ? g8 : (g8 ^ 1);
    const int pi_s0 = (q4 < 2) ? (2 * q4) : (2 * q4 + 1);
    const int pi_s1 = (q4 < 2) ? (2 * q4 + 1) : (2 * q4);
    const int off1 = pi_g * LDS + q4;
    const int off2_s0 = pi_s0 * LDS + g8;
    const int off2_s1 = pi_s1 * LDS + g8;
    double xf[KS];
    double ga[NT][2];
    double ssq = 0.0;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      const int k = 4 * kk + q4;
      xf[kk] = (g8 < p.B && k < p.dl) ? p.X[(size_t)g8 * p.d + p.col[k]] : 0.0;
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) ga[t][0] = ga[t][1] = 0.0;
    // this warp's stages are warp, warp + SNW, ... (STAGES / SNW of them), visited round-robin
    int stage = warp;
    uint32_t phase = 0;
    for (int i = warp; i < nb; i += SNW) {
      ptx::mbar_wait(&full_bar[stage], phase);
      const double* As = reinterpret_cast<const double*>(smem + (size_t)stage * STAGE_BYTES);
      const double2* Ys = reinterpret_cast<const double2*>(smem + (size_t)stage * STAGE_BYTES + A_BYTES);
      // GEMM1 as two interleaved accumulator chains (even / odd k4 steps), summed afterwards
      double z0 = 0.0, z1 = 0.0, u0 = 0.0, u1 = 0.0;
#pragma unroll
      for (int kk = 0; kk + 1 < KS; kk += 2) {
        ptx::dmma884(z0, z1, xf[kk], As[off1 + 4 * kk]);
        ptx::dmma884(u0, u1, xf[kk + 1], As[off1 + 4 * (kk + 1)]);
      }
      z0 += u0;
      z1 += u1;
      const double2 y0 = Ys[pi_s0];
      const double2 y1 = Ys[pi_s1];
      const double r0 = y0.x - z0, r1 = y1.x - z1;
      const double w0 = r0 * y0.y, w1 = r1 * y1.y;
      ssq = fma(r0, w0, ssq);
      ssq = fma(r1, w1, ssq);
      if (p.grad) {
#pragma unroll
        for (int t = 0; t < NT; ++t) ptx::dmma884(ga[t][0], ga[t][1], w0, As[off2_s0 + 8 * t]);
#pragma unroll
        for (int t = 0; t < NT; ++t) ptx::dmma884(ga[t][0], ga[t][1], w1, As[off2_s1 + 8 * t]);
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);
      stage += SNW;
      if (stage >= STAGES) {
        stage = warp;
        phase ^= 1;
      }
    }
    const double qsum = quad_sum(ssq);
    if (q4 == 0) redq[warp * 8 + g8] = qsum;
#pragma unroll
    for (int t = 0; t < NT; ++t)
      *reinterpret_cast<double2*>(red + (warp * 8 + g8) * NP + 8 * t + 2 * q4) = make_double2(ga[t][0], ga[t][1]);
  }
  stream_cta_tail(p, red, redq, 8, NP, &s_last, s_q);
}

template <int NT>
thy_status launch_dmma_nt(thy_model_s* m, StreamParams& p, cudaStream_t s) {
  using C = StreamDmmaCfg<NT>;
  p.stages = C::STAGES;
  p.nblk = p.nblk * (SMB / SR);   // 8-row blocks
  int G = sm_count();
  if (G > p.nblk) G = p.nblk;
  THY_TRY(m->ws.partG.reserve((size_t)G * 8 * C::NP));
  THY_TRY(m->ws.partL.reserve((size_t)G * 8));
  if (!m->ws.ticket.ptr) {
    THY_TRY(m->ws.ticket.reserve(1));
    THY_CUDA_CHECK(cudaMemsetAsync(m->ws.ticket.ptr, 0, sizeof(unsigned int), s));
  }
  p.partG = m->ws.partG.ptr;
  p.partL = m->ws.partL.ptr;
  p.ticket = m->ws.ticket.ptr;
  THY_CUDA_CHECK(cudaFuncSetAttribute(k_stream_dmma<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  KernelTimer timer(m, s);
  k_stream_dmma<NT><<<G, (SNW + 1) * 32, C::SMEM, s>>>(p);
  timer.stop();
  count_launch();
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

template <int BT, int T>
thy_status launch_bt(thy_model_s* m, StreamParams& p, cudaStream_t s) {
  const int stage_bytes = SMB * p.lds * 8 + SMB * 16;
  const size_t red_bytes = (size_t)(SNW * BT * 32 * T + SNW * BT) * sizeof(double);
  int stages = (int)((200 * 1024 - red_bytes) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return fail(THY_ERR_UNSUPPORTED, "streaming kernel: tile does not fit shared memory");
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 2 * stages * sizeof(uint64_t) + red_bytes;
  int G = sm_count();
  if (G > p.nblk) G = p.nblk;
  THY_TRY(m->ws.partG.reserve((size_t)G * BT * 32 * T));
  THY_TRY(m->ws.partL.reserve((size_t)G * BT));
  if (!m->ws.ticket.ptr) {
    THY_TRY(m->ws.ticket.reserve(1));
    THY_CUDA_CHECK(cudaMemsetAsync(m->ws.ticket.ptr, 0, sizeof(unsigned int), s));
  }
  p.partG = m->ws.partG.ptr;
  p.partL = m->ws.partL.ptr;
  p.ticket = m->ws.ticket.ptr;
  THY_CUDA_CHECK(cudaFuncSetAttribute(k_stream<BT, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KernelTimer timer(m, s);
  k_stream<BT, T><<<G, (SNW + 1) * 32, smem, s>>>(p);
  timer.stop();
  count_launch();
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

template <int BT>
thy_status launch_t(thy_model_s* m, StreamParams& p, cudaStream_t s) {
  switch ((p.dl + 31) / 32) {
    case 1: return launch_bt<BT, 1>(m, p, s);
    case 2: return launch_bt<BT, 2>(m, p, s);
    case 3: return launch_bt<BT, 3>(m, p, s);
    case 4: return launch_bt<BT, 4>(m, p, s);
    default: return fail(THY_ERR_UNSUPPORTED, "streaming kernel: unsupported width");
  }
}

}  // namespace

bool stream_supported(const LikDev& lik, int64_t B) {
  return B >= 1 && B <= 32 && lik.dl <= 128 && fused_dmma_supported(lik, true);
}

// Auto-selection cost model (seconds): one pass over A at ~5.5 TB/s (or a launch latency) per group of 8 chains.
double stream_cost_model(const LikDev& lik, int64_t B) {
  const double bytes = (double)lik.n_pad * lik.lds * 8.0;
  const double per_pass = bytes / 5.5e12 > 6e-6 ? bytes / 5.5e12 : 6e-6;
  return (double)((B + 7) / 8) * per_pass;
}

static thy_status launch_stream_group(thy_model_s* m, const LikDevOwner& lik, int B, const double* X, double* logp,
                                      double* grad, double cauchy_sign, cudaStream_t s) {
  const LikDev& lk = lik.view;
  StreamParams p{};
  p.A = lk.A;
  p.yiv = lk.yiv_lin;
  p.col = lk.col;
  p.X = X;
  p.logp = logp;
  p.grad = grad;
  p.B = B;
  p.d = m->d;
  p.dl = lk.dl;
  p.np = lk.lds - 4;
  p.lds = lk.lds;
  p.n = lk.n;
  p.nblk = lk.n_pad / SMB;
  p.distribution = lk.distribution;
  p.log_const = lk.log_const;
  p.cauchy_sign = cauchy_sign;
  if (B <= 1) return launch_t<1>(m, p, s);   // one chain: lanes stride the columns, warp-shuffle dot products
  m->last_path = "stream_dmma";              // 2..8 chains: one DMMA row group, any model length
  switch (fused_nt_for(lk.dl)) {             // 2..8 chains: one DMMA row group
    case 1: return launch_dmma_nt<1>(m, p, s);
    case 2: return launch_dmma_nt<2>(m, p, s);
    case 3: return launch_dmma_nt<3>(m, p, s);
    case 4: return launch_dmma_nt<4>(m, p, s);
    case 5: return launch_dmma_nt<5>(m, p, s);
    case 7: return launch_dmma_nt<7>(m, p, s);
    case 10: return launch_dmma_nt<10>(m, p, s);
    case 13: return launch_dmma_nt<13>(m, p, s);
    case 16: return launch_dmma_nt<16>(m, p, s);
    default: return fail(THY_ERR_UNSUPPORTED, "streaming kernel: unsupported width");
  }
}

// Up to 32 chains: one pass over A per group of 8 chains.
thy_status launch_stream(thy_model_s* m, const LikDevOwner& lik, int64_t B, const double* X, double* logp, double* grad,
                         double cauchy_sign, cudaStream_t s) {
  for (int64_t b0 = 0; b0 < B; b0 += 8) {
    const int nb = (int)((B - b0 < 8) ? (B - b0) : 8);
    THY_TRY(launch_stream_group(m, lik, nb, X + b0 * m->d, logp + b0, grad ? grad + b0 * m->d : nullptr, cauchy_sign, s));
  }
  return THY_OK;
}

}  // namespace thy
