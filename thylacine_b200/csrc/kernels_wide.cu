// Wide path: any parameter width (d > 128, e.g. the 1000-d Leapfrog-MCMC config) and any link, on the fp64 tensor cores.
//
// For wide models the chain vectors no longer fit in registers, so the fused single-pass kernel does not apply; but the
// arithmetic intensity is so high (2 d flop per byte of W) that materialising the weighted residual W between the two
// GEMMs costs < 2 % of the compute time.  Two DMMA.8x8x4 GEMM kernels with fused epilogues:
//   forward   Z = X A^T (+offset), f = link(Z), r = y - f, w = r / var, per-chain partial sums of r.w;
//             W (times link'(Z) for the analytic Jacobian) is written in the tiled layout the adjoint consumes
//                                                                 Likelihood.scala:43-49, LinearForwardModel.scala:103-114
//   finish    k_finish_generic: observation log-pdf, Cauchy multiplier    Gaussian/CauchyDistribution.scala
//   adjoint   grad += scale_b * W A                                        Likelihood.scala:51-65
// Data movement: every operand lives in HBM as CONTIGUOUS 128 x 32 (or 32 x 128) tiles with a padded row pitch
// (36 / 132 doubles = 4 mod 16), so one pipeline stage is two 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) and every
// DMMA fragment load from shared memory is bank-conflict free.  A is re-tiled once per model (lazily, on the device);
// X is packed per call (B d doubles: negligible); W is produced directly in tiled form by the forward epilogue.
// CTA tile 128 chains x 128 (observations | columns), 8 consumer warps as 4 x 2, warp tile 32 x 64 (64 accumulators per
// lane), 1 producer warp, 3-stage mbarrier ring, K chunk 32.  Grid: chain tiles fastest, so CTAs that run together share
// the same A tiles through L2.
#include "kernels.h"
#include "ptx.cuh"

namespace thy {

namespace {

constexpr int WT = 128;          // CTA tile edge
constexpr int WK = 32;           // K chunk
constexpr int WP = WK + 4;       // pitch of K-contiguous tiles (36 = 4 mod 16)
constexpr int WQ = WT + 4;       // pitch of N-contiguous tiles (132 = 4 mod 16)
constexpr int WNW = 8;           // consumer warps (4 x 2)
constexpr int WSTAGES = 3;
constexpr int TILE_KC = WT * WP; // doubles in a 128 x 32 (padded) tile
constexpr int TILE_NC = WK * WQ; // doubles in a 32 x 128 (padded) tile

struct WideParams {
  const double* XT;     // [bt][kc][128][36]   packed chain vectors
  const double* AF;     // [it][kc][128][36]   A tiled for the forward GEMM
  const double* AG;     // [jt][ic][32][132]   A tiled for the adjoint GEMM
  double* WTl;          // [bt][ic][128][36]   weighted residuals (tiled), may be null (value only)
  double* partS;        // [B][n_tiles]
  const double* scale;  // [B]
  double* grad;         // [B][d]
  long long B;
  int d, ktiles, n_tiles, ictiles, jtiles, write_deriv;
};

// ---- layout kernels ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wide_retile_af(LikDev lk, int n_tiles, int ktiles, double* __restrict__ AF) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_tiles * ktiles * TILE_KC;
  if (idx >= total) return;
  const int c = (int)(idx % WP), r = (int)((idx / WP) % WT);
  const long long tile = idx / TILE_KC;
  const int kc = (int)(tile % ktiles), it = (int)(tile / ktiles);
  const int i = it * WT + r, k = kc * WK + c;
  AF[idx] = (c < WK && i < lk.n && k < lk.dl) ? lk.A[(size_t)i * lk.lds + k] : 0.0;
}

__global__ void __launch_bounds__(256) k_wide_retile_ag(LikDev lk, int jtiles, int ictiles, double* __restrict__ AG) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)jtiles * ictiles * TILE_NC;
  if (idx >= total) return;
  const int c = (int)(idx % WQ), r = (int)((idx / WQ) % WK);
  const long long tile = idx / TILE_NC;
  const int ic = (int)(tile % ictiles), jt = (int)(tile / ictiles);
  const int i = ic * WK + r, j = jt * WT + c;
  AG[idx] = (c < WT && i < lk.n && j < lk.dl) ? lk.A[(size_t)i * lk.lds + j] : 0.0;
}

__global__ void __launch_bounds__(256) k_wide_pack_x(LikDev lk, int d, long long B, const double* __restrict__ X, int ktiles,
                                                     double* __restrict__ XT) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = ((B + WT - 1) / WT) * (long long)ktiles * TILE_KC;
  if (idx >= total) return;
  const int c = (int)(idx % WP), r = (int)((idx / WP) % WT);
  const long long tile = idx / TILE_KC;
  const int kc = (int)(tile % ktiles);
  const long long b = (tile / ktiles) * WT + r;
  const int k = kc * WK + c;
  XT[idx] = (c < WK && b < B && k < lk.dl) ? X[b * d + lk.col[k]] : 0.0;
}

// ---- shared main loop ----------------------------------------------------------------------------------------------
// acc[mi][ni] (8x8 C fragments) += P[128 x 32] . Q, with P K-contiguous (pitch 36) and Q either K-contiguous rows of
// observations (forward: Q[n][k], pitch 36) or K-major rows of columns (adjoint: Q[k][n], pitch 132).
template <bool Q_KMAJOR>
__device__ __forceinline__ void wide_stage_mma(const double* __restrict__ Ps, const double* __restrict__ Qs,
                                               double (&acc)[4][8][2], int wm, int wn, int g8, int q4) {
#pragma unroll
  for (int kk = 0; kk < WK / 4; ++kk) {
    double a[4], b[8];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) a[mi] = Ps[(wm * 32 + mi * 8 + g8) * WP + kk * 4 + q4];
#pragma unroll
    for (int ni = 0; ni < 8; ++ni)
      b[ni] = Q_KMAJOR ? Qs[(kk * 4 + q4) * WQ + wn * 64 + ni * 8 + g8] : Qs[(wn * 64 + ni * 8 + g8) * WP + kk * 4 + q4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) ptx::dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
  }
}

template <bool Q_KMAJOR>
__device__ __forceinline__ void wide_pipeline(const double* __restrict__ Pg, const double* __restrict__ Qg, int nchunks,
                                              unsigned char* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                              double (&acc)[4][8][2]) {
  constexpr int P_BYTES = TILE_KC * 8;
  constexpr int Q_BYTES = (Q_KMAJOR ? TILE_NC : TILE_KC) * 8;
  constexpr int STAGE_BYTES = P_BYTES + Q_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == WNW) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < nchunks; ++c) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        unsigned char* dst = smem + (size_t)stage * STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
        ptx::bulk_g2s(dst, Pg + (size_t)c * TILE_KC, P_BYTES, &full_bar[stage]);
        ptx::bulk_g2s(dst + P_BYTES, Qg + (size_t)c * (Q_KMAJOR ? TILE_NC : TILE_KC), Q_BYTES, &full_bar[stage]);
        if (++stage == WSTAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    return;
  }
  const int wm = warp >> 1, wn = warp & 1, g8 = lane >> 2, q4 = lane & 3;
  int stage = 0;
  uint32_t phase = 0;
  for (int c = 0; c < nchunks; ++c) {
    ptx::mbar_wait(&full_bar[stage], phase);
    const double* Ps = reinterpret_cast<const double*>(smem + (size_t)stage * STAGE_BYTES);
    const double* Qs = reinterpret_cast<const double*>(smem + (size_t)stage * STAGE_BYTES + P_BYTES);
    wide_stage_mma<Q_KMAJOR>(Ps, Qs, acc, wm, wn, g8, q4);
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);
    if (++stage == WSTAGES) {
      stage = 0;
      phase ^= 1;
    }
  }
}

__device__ __forceinline__ void wide_init_barriers(uint64_t* full_bar, uint64_t* empty_bar) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < WSTAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], WNW);
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  __syncthreads();
}

// ---- forward -------------------------------------------------------------------------------------------------------
// grid (chain tiles, observation tiles).  FUSED_EPI: the residual epilogue runs on the accumulators (identity link: a
// handful of flops per element).  Otherwise the raw pre-activations are stored in the tiled W buffer and k_wide_link
// applies link / residual / weights at full occupancy: fp64 transcendentals are long dependent chains, and with one
// 8-warp CTA per SM they would otherwise dominate the kernel.
template <bool FUSED_EPI>
__global__ void __launch_bounds__((WNW + 1) * 32, 1) k_wide_fwd(const LikDev lk, const WideParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int STAGE_BYTES = 2 * TILE_KC * 8;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)WSTAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + WSTAGES;
  __shared__ double sred[2][WT];
  wide_init_barriers(full_bar, empty_bar);
  const int bt = blockIdx.x, it = blockIdx.y;
  double acc[4][8][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  wide_pipeline<false>(p.XT + (size_t)bt * p.ktiles * TILE_KC, p.AF + (size_t)it * p.ktiles * TILE_KC, p.ktiles, smem, full_bar,
                       empty_bar, acc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < WNW) {
    const int wm = warp >> 1, wn = warp & 1, g8 = lane >> 2, q4 = lane & 3;
    if constexpr (!FUSED_EPI) {
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) {
        const int nloc = wn * 64 + ni * 8 + 2 * q4;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          const int m = wm * 32 + mi * 8 + g8;
          double* dst = p.WTl + ((size_t)bt * p.ictiles + (it * (WT / WK) + nloc / WK)) * TILE_KC + m * WP + (nloc % WK);
          *reinterpret_cast<double2*>(dst) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
      }
    } else {
      double srow[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) {
        const int nloc = wn * 64 + ni * 8 + 2 * q4;       // observation inside the CTA tile (and nloc + 1)
        const int i0 = it * WT + nloc;
        double offv[2], yv[2], ivv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = i0 + e;
          const bool ok = i < lk.n;
          const double2 t = ok ? lk.yiv[i] : make_double2(0.0, 0.0);
          offv[e] = ok ? lk.off[i] : 0.0;
          yv[e] = t.x;
          ivv[e] = t.y;
        }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          double wv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double res = yv[e] - (acc[mi][ni][e] + offv[e]);
            const double w = res * ivv[e];
            srow[mi] = fma(res, w, srow[mi]);
            wv[e] = w;
          }
          if (p.WTl) {
            const int m = wm * 32 + mi * 8 + g8;
            double* dst = p.WTl + ((size_t)bt * p.ictiles + (it * (WT / WK) + nloc / WK)) * TILE_KC + m * WP + (nloc % WK);
            *reinterpret_cast<double2*>(dst) = make_double2(wv[0], wv[1]);
          }
        }
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const double s = quad_sum(srow[mi]);
        if (q4 == 0) sred[wn][wm * 32 + mi * 8 + g8] = s;
      }
    }
  }
  if constexpr (FUSED_EPI) {
    __syncthreads();
    if (threadIdx.x < WT) {
      const long long b = (long long)bt * WT + threadIdx.x;
      if (b < p.B) p.partS[b * p.n_tiles + it] = sred[0][threadIdx.x] + sred[1][threadIdx.x];
    }
  }
}

// Link / residual / weights on the tiled pre-activations, in place: one warp per (chain, 128-observation tile).
__global__ void __launch_bounds__(256) k_wide_link(const LikDev lk, const WideParams p) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= p.B * p.n_tiles) return;
  const long long b = wid / p.n_tiles;
  const int it = (int)(wid % p.n_tiles);
  const long long bt = b / WT;
  const int m = (int)(b % WT);
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < WT / WK; ++q) {
    const int ic = it * (WT / WK) + q;
    const int i = ic * WK + lane;
    double* ptr = p.WTl + ((size_t)bt * p.ictiles + ic) * TILE_KC + m * WP + lane;
    double w = 0.0;
    if (i < lk.n) {
      const double z = *ptr + lk.off[i];
      const double f = link_eval(lk.link, z);
      const double2 yv = lk.yiv[i];
      const double res = yv.x - f;
      w = res * yv.y;
      s = fma(res, w, s);
      if (p.write_deriv) w *= link_deriv_from(lk.link, z, f);
    }
    *ptr = w;
  }
  s = warp_sum(s);
  if (lane == 0) p.partS[b * p.n_tiles + it] = s;
}

// ---- adjoint -------------------------------------------------------------------------------------------------------
// grid (chain tiles, column tiles)
__global__ void __launch_bounds__((WNW + 1) * 32, 1) k_wide_adj(const LikDev lk, const WideParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int STAGE_BYTES = (TILE_KC + TILE_NC) * 8;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)WSTAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + WSTAGES;
  wide_init_barriers(full_bar, empty_bar);
  const int bt = blockIdx.x, jt = blockIdx.y;
  double acc[4][8][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  wide_pipeline<true>(p.WTl + (size_t)bt * p.ictiles * TILE_KC, p.AG + (size_t)jt * p.ictiles * TILE_NC, p.ictiles, smem,
                      full_bar, empty_bar, acc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= WNW) return;
  const int wm = warp >> 1, wn = warp & 1, g8 = lane >> 2, q4 = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
    const long long b = (long long)bt * WT + wm * 32 + mi * 8 + g8;
    if (b >= p.B) continue;
    const double sc = p.scale[b];
    double* grow = p.grad + b * p.d;
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      const int j = jt * WT + wn * 64 + ni * 8 + 2 * q4;
      if (j < lk.dl) grow[lk.col[j]] += sc * acc[mi][ni][0];
      if (j + 1 < lk.dl) grow[lk.col[j + 1]] += sc * acc[mi][ni][1];
    }
  }
}

}  // namespace

// user links: THY_JAC_FINITE_DIFFERENCE means the link's own forward difference inside the compiled step
static bool wide_user_link(const LikDev& lk) { return lk.link >= THY_LINK_USER_BASE; }

bool wide_supported(const LikDev& lk) {
  if (wide_user_link(lk)) return lk.distribution != THY_DIST_UNIFORM && lk.n >= 1 && lk.dl >= 1;
  return lk.distribution != THY_DIST_UNIFORM && lk.jacobian != THY_JAC_FINITE_DIFFERENCE && lk.n >= 1 && lk.dl >= 1;
}

// k_finish_generic lives in kernels_generic.cu
thy_status launch_finish_generic(const LikDev& lk, long long B, const double* partS, int n_tiles, double* logp, double* scale,
                                 double cauchy_sign, cudaStream_t s);

static thy_status ensure_tiled(thy_model_s* m, LikDevOwner& own, bool want_adj, cudaStream_t s) {
  const LikDev& lk = own.view;
  const int n_tiles = ceil_div(lk.n, WT), ktiles = ceil_div(lk.dl, WK);
  const int ictiles = n_tiles * (WT / WK), jtiles = ceil_div(lk.dl, WT);
  (void)m;
  if (!own.AF.ptr) {
    const long long total = (long long)n_tiles * ktiles * TILE_KC;
    THY_TRY(own.AF.reserve((size_t)total));
    k_wide_retile_af<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(lk, n_tiles, ktiles, own.AF.ptr);
    count_launch();
  }
  if (want_adj && !own.AG.ptr) {
    const long long total = (long long)jtiles * ictiles * TILE_NC;
    THY_TRY(own.AG.reserve((size_t)total));
    k_wide_retile_ag<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(lk, jtiles, ictiles, own.AG.ptr);
    count_launch();
  }
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

thy_status launch_wide_likelihood(thy_model_s* m, LikDevOwner& own, int64_t B, const double* X, double* logp, double* grad,
                                  double cauchy_sign, cudaStream_t s) {
  const LikDev& lk = own.view;
  const bool want_adj = grad != nullptr;
  const bool fused_epi = lk.link == THY_LINK_IDENTITY;
  const bool want_buf = want_adj || !fused_epi;
  THY_TRY(ensure_tiled(m, own, want_adj, s));
  const int n_tiles = ceil_div(lk.n, WT), ktiles = ceil_div(lk.dl, WK);
  const int ictiles = n_tiles * (WT / WK), jtiles = ceil_div(lk.dl, WT);
  // chunk the batch so that the tiled W stays below ~8 GiB
  const size_t w_per_chain = (size_t)ictiles * WP * sizeof(double);
  long long chunk = (long long)((8ull << 30) / w_per_chain) / WT * WT;
  if (chunk < WT) chunk = WT;
  if (chunk > B) chunk = (long long)round_up(B, WT);
  const long long btiles_max = chunk / WT;
  THY_TRY(m->ws.XT.reserve((size_t)btiles_max * ktiles * TILE_KC));
  if (want_buf) THY_TRY(m->ws.W.reserve((size_t)btiles_max * ictiles * TILE_KC));
  THY_TRY(m->ws.partL.reserve((size_t)chunk * n_tiles + (size_t)chunk));
  double* partS = m->ws.partL.ptr;
  double* scale = partS + (size_t)chunk * n_tiles;
  const size_t smem_f = (size_t)WSTAGES * 2 * TILE_KC * 8 + 2 * WSTAGES * sizeof(uint64_t);
  const size_t smem_g = (size_t)WSTAGES * (TILE_KC + TILE_NC) * 8 + 2 * WSTAGES * sizeof(uint64_t);
  THY_CUDA_CHECK(cudaFuncSetAttribute(k_wide_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
  THY_CUDA_CHECK(cudaFuncSetAttribute(k_wide_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
  THY_CUDA_CHECK(cudaFuncSetAttribute(k_wide_adj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
  for (long long c0 = 0; c0 < B; c0 += chunk) {
    const long long nb = (B - c0 < chunk) ? (B - c0) : chunk;
    const int btiles = ceil_div(nb, WT);
    WideParams p{};
    p.XT = m->ws.XT.ptr;
    p.AF = own.AF.ptr;
    p.AG = own.AG.ptr;
    p.WTl = want_buf ? m->ws.W.ptr : nullptr;
    p.partS = partS;
    p.scale = scale;
    p.grad = grad ? grad + c0 * m->d : nullptr;
    p.B = nb;
    p.d = m->d;
    p.ktiles = ktiles;
    p.n_tiles = n_tiles;
    p.ictiles = ictiles;
    p.jtiles = jtiles;
    p.write_deriv = (lk.link != THY_LINK_IDENTITY) ? 1 : 0;
    const long long xt_total = (long long)btiles * ktiles * TILE_KC;
    k_wide_pack_x<<<(unsigned)ceil_div(xt_total, 256), 256, 0, s>>>(lk, m->d, nb, X + c0 * m->d, ktiles, m->ws.XT.ptr);
    {
      KernelTimer timer(m, s);
      if (fused_epi)
        k_wide_fwd<true><<<dim3(btiles, n_tiles), (WNW + 1) * 32, smem_f, s>>>(lk, p);
      else
        k_wide_fwd<false><<<dim3(btiles, n_tiles), (WNW + 1) * 32, smem_f, s>>>(lk, p);
      timer.stop();
    }
    if (wide_user_link(lk)) {
      UserLink& ul = *m->user_links[lk.link - THY_LINK_USER_BASE];
      THY_TRY(user_link_launch(ul, p.WTl, lk.off, lk.yiv, p.partS, nb, lk.n, n_tiles, ictiles, p.write_deriv, lk.differential,
                               WT, WK, WP, TILE_KC, s));
    } else if (!fused_epi) {
      k_wide_link<<<(unsigned)ceil_div(nb * n_tiles, 8), 256, 0, s>>>(lk, p);
      count_launch();
    }
    THY_TRY(launch_finish_generic(lk, nb, partS, n_tiles, logp + c0, scale, cauchy_sign, s));
    count_launch(2);
    if (want_adj) {
      k_wide_adj<<<dim3(btiles, jtiles), (WNW + 1) * 32, smem_g, s>>>(lk, p);
      count_launch();
    }
    THY_CUDA_CHECK(cudaGetLastError());
  }
  return THY_OK;
}

}  // namespace thy
