// Generic likelihood path: any link / distribution / Jacobian mode and any shape.
// Two tiled fp64 GEMM kernels with the weighted residual W materialised between them:
//   forward   Z = X A^T (+offset), f = link(Z), r = y - f, w = r / var          Likelihood.scala:43-49
//   finish    per-chain reduction of r.w  -> log-pdf of the observation distribution, Cauchy multiplier
//   adjoint   grad += scale_b * (W .* link'(Z)) A                                Likelihood.scala:51-65
//   FD        J[:,j] = (f(theta + h e_j) - f(theta)) * (1/h), all d+1 perturbed models in one launch
//                                                    core/computation/FiniteDifferenceJacobian.scala:30-55
// This is the correctness backstop; the fused DMMA and streaming kernels are the fast paths.
#include "kernels.h"

namespace thy {

constexpr int GT = 64;   // tile edge (chains x obs, chains x columns)
constexpr int GK = 16;   // k chunk

// ---- forward: W[b][i] (and optionally Z[b][i]) + per-tile partial row sums --------------------------------
// grid: (n_pad/GT tiles over observations, ceil(B/GT) tiles over chains); block 256 = 16 x 16, 4x4 per thread.
__global__ void __launch_bounds__(256) k_fwd_generic(LikDev lk, int d, long long B, const double* __restrict__ X,
                                                     double* __restrict__ W, double* __restrict__ Z,
                                                     double* __restrict__ partS, int n_tiles, int write_deriv) {
  __shared__ double Xs[GT][GK + 1];
  __shared__ double As[GT][GK + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long b0 = (long long)blockIdx.y * GT;
  const int i0 = blockIdx.x * GT;
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;

  for (int k0 = 0; k0 < lk.dl; k0 += GK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = threadIdx.x + 256 * e;
      const int row = idx >> 4, kk = idx & 15;
      const int k = k0 + kk;
      double xv = 0.0, av = 0.0;
      if (k < lk.dl) {
        const long long b = b0 + row;
        if (b < B) xv = X[b * d + lk.col[k]];
        const int i = i0 + row;
        if (i < lk.n_pad) av = lk.A[(size_t)i * lk.lds + k];
      }
      Xs[row][kk] = xv;
      As[row][kk] = av;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      double xr[4], ar[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) xr[r] = Xs[ty * 4 + r][kk];
#pragma unroll
      for (int c = 0; c < 4; ++c) ar[c] = As[tx + 16 * c][kk];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fma(xr[r], ar[c], acc[r][c]);
    }
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const long long b = b0 + ty * 4 + r;
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + tx + 16 * c;
      if (b < B && i < lk.n_pad) {
        double w = 0.0, z = 0.0;
        if (i < lk.n) {
          z = acc[r][c] + lk.off[i];
          const double f = link_eval(lk.link, z);
          const double2 yv = lk.yiv[i];
          if (lk.distribution == THY_DIST_UNIFORM) {
            // count violations of the half-open box [lower, upper)
            s += (f >= yv.y && f < yv.x) ? 0.0 : 1.0;
          } else {
            const double res = yv.x - f;
            w = res * yv.y;
            s += res * w;
            if (write_deriv) w *= link_deriv_from(lk.link, z, f);
          }
        }
        W[b * lk.n_pad + i] = w;
        if (Z) Z[b * lk.n_pad + i] = z;
      }
    }
    // reduce over the 16 tx lanes of this (warp-half) row group
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (tx == 0 && b < B) partS[b * n_tiles + blockIdx.x] = s;
  }
}

// ---- finish: logp += observation log-pdf; scale[b] = multiplier applied to the adjoint ----------------------
__global__ void __launch_bounds__(256) k_finish_generic(LikDev lk, long long B, const double* __restrict__ partS,
                                                        int n_tiles, double* __restrict__ logp,
                                                        double* __restrict__ scale, double cauchy_sign) {
  const int lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  double s = 0.0;
  for (int t = lane; t < n_tiles; t += 32) s += partS[b * n_tiles + t];
  s = warp_sum(s);
  if (lane == 0) {
    double lp, sc = 1.0;
    if (lk.distribution == THY_DIST_GAUSSIAN) {
      lp = -0.5 * s + lk.log_const;
    } else if (lk.distribution == THY_DIST_CAUCHY) {
      lp = lk.log_const - (1.0 + lk.n) / 2.0 * log(1.0 + s);
      // reference: +(1+n)/(1+q) S^-1 (f - y) = -(1+n)/(1+q) w ; correct sign: +(1+n)/(1+q) w
      sc = -cauchy_sign * (1.0 + lk.n) / (1.0 + s);
    } else {
      lp = (s == 0.0) ? lk.log_const : -INFINITY;
      sc = 0.0;
    }
    logp[b] += lp;
    if (scale) scale[b] = sc;
  }
}

// ---- adjoint: grad[b][col[j]] += scale[b] * sum_i W[b][i] A[i][j] --------------------------------------------
__global__ void __launch_bounds__(256) k_adj_generic(LikDev lk, int d, long long B, const double* __restrict__ W,
                                                     const double* __restrict__ scale, double* __restrict__ grad) {
  __shared__ double Ws[GT][GK + 1];
  __shared__ double As[GK][GT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long b0 = (long long)blockIdx.y * GT;
  const int j0 = blockIdx.x * GT;
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
  for (int i0 = 0; i0 < lk.n_pad; i0 += GK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = threadIdx.x + 256 * e;
      {
        const int row = idx >> 4, kk = idx & 15;
        const long long b = b0 + row;
        Ws[row][kk] = (b < B) ? W[b * lk.n_pad + i0 + kk] : 0.0;
      }
      {
        const int kk = idx >> 6, jj = idx & 63;
        const int j = j0 + jj;
        As[kk][jj] = (j < lk.dl) ? lk.A[(size_t)(i0 + kk) * lk.lds + j] : 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      double wr[4], ar[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) wr[r] = Ws[ty * 4 + r][kk];
#pragma unroll
      for (int c = 0; c < 4; ++c) ar[c] = As[kk][tx + 16 * c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fma(wr[r], ar[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const long long b = b0 + ty * 4 + r;
    if (b >= B) continue;
    const double sc = scale[b];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx + 16 * c;
      if (j < lk.dl) grad[b * d + lk.col[j]] += sc * acc[r][c];
    }
  }
}

// ---- finite-difference adjoint --------------------------------------------------------------------------------
// One launch covers every (chain, perturbed coordinate) pair, i.e. all dl+1 forward-model evaluations per chain
// (the unperturbed one is Z from the forward kernel).  grad_j = sum_i w_i (f(theta + h e_j)_i - f(theta)_i) (1/h).
// incremental = 0: the perturbed pre-activation is recomputed from the nudged parameter vector (literal);
// incremental = 1: z + A[:,j] * ((theta_j + h) - theta_j)  (exact-arithmetic equal, one rounding instead of dl).
constexpr int FD_TI = 1024;  // observation rows staged per shared-memory tile

__global__ void __launch_bounds__(128) k_fd_adjoint(LikDev lk, int d, long long B, const double* __restrict__ X,
                                                    const double* __restrict__ W, const double* __restrict__ Z,
                                                    const double* __restrict__ scale, double* __restrict__ grad,
                                                    int incremental) {
  extern __shared__ double sh[];  // theta[dl] | z0[FD_TI] | f0[FD_TI] | w[FD_TI]
  double* zs = sh + lk.dl;
  double* fs = zs + FD_TI;
  double* ws = fs + FD_TI;
  const long long b = blockIdx.x;
  if (b >= B) return;
  for (int k = threadIdx.x; k < lk.dl; k += blockDim.x) sh[k] = X[b * d + lk.col[k]];
  __syncthreads();
  const int j = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = j < lk.dl;
  const double h = lk.differential;
  const double inv_h = 1 / h;
  const double nudged = active ? sh[j] + h : 0.0;
  const double dj = active ? nudged - sh[j] : 0.0;
  double acc = 0.0;
  for (int i0 = 0; i0 < lk.n; i0 += FD_TI) {
    const int ni = (lk.n - i0 < FD_TI) ? (lk.n - i0) : FD_TI;
    if (incremental) {
      // the unperturbed model f(theta) is evaluated once per observation and shared by all perturbed coordinates
      for (int i = threadIdx.x; i < ni; i += blockDim.x) {
        const double z0 = Z[b * lk.n_pad + i0 + i];
        zs[i] = z0;
        fs[i] = link_eval(lk.link, z0);
        ws[i] = W[b * lk.n_pad + i0 + i];
      }
      __syncthreads();
      if (active) {
        for (int i = 0; i < ni; ++i) {
          const double z1 = fma(lk.A[(size_t)(i0 + i) * lk.lds + j], dj, zs[i]);
          const double jac = (link_eval(lk.link, z1) - fs[i]) * inv_h;
          acc = fma(ws[i], jac, acc);
        }
      }
      __syncthreads();
    } else if (active) {
      for (int i = i0; i < i0 + ni; ++i) {
        // both evaluations through the same summation order, as the reference's closure would
        const double* arow = lk.A + (size_t)i * lk.lds;
        double s0 = 0.0, s1 = 0.0;
        for (int k = 0; k < lk.dl; ++k) {
          const double a = arow[k];
          s0 = fma(a, sh[k], s0);
          s1 = fma(a, (k == j) ? nudged : sh[k], s1);
        }
        const double z0 = s0 + lk.off[i];
        const double z1 = s1 + lk.off[i];
        const double jac = (link_eval(lk.link, z1) - link_eval(lk.link, z0)) * inv_h;
        acc = fma(W[b * lk.n_pad + i], jac, acc);
      }
    }
  }
  if (active) grad[b * d + lk.col[j]] += scale[b] * acc;
}

// Literal finite differences without the redundant work.  The reference evaluates every perturbed model as a full
// forward map; for f = link(A theta + offset) that is, per observation row, the dot product sum_k a_k theta^(j)_k with
// theta^(j) = theta + h e_j.  Accumulated in k order (one fma chain), every term before k = j is the SAME rounding
// sequence as the unperturbed dot product, so the perturbed chain can start from the unperturbed PREFIX P_j and only
// run k = j .. dl-1: bit-for-bit the value of the full chain (and of k_fd_adjoint's literal branch), for dl^2 / 2
// instead of 2 dl^2 fused multiply-adds per row.
// Mapping: a tile of 32 observation rows sits in shared memory with its prefix arrays; lane = row, and a warp walks
// one block of FDP_J perturbed columns at a time, so every a_k / theta_k read feeds FDP_J chains (register blocking: the
// plain mapping was shared-memory bound) and all lanes of a warp have the same trip count.  Column blocks are taken in
// (q, last - q) pairs so every warp does the same number of chain steps.  The Jacobian entries overwrite the prefix
// slots they started from; a last pass adds w_i J_ij over the rows in ascending order (the plain kernel's order).
constexpr int FDP_J = 8;

__device__ __forceinline__ void fdp_block(const LikDev& lk, int dl, int pitch, int j0, int r, bool row_live,
                                          const double* __restrict__ theta, const double* __restrict__ At,
                                          double* __restrict__ P, double h, double inv_h, double f0r, double offr) {
  const double* a = At + (size_t)r * pitch;
  double* pr = P + (size_t)r * pitch;
  double s[FDP_J];
#pragma unroll
  for (int u = 0; u < FDP_J; ++u) s[u] = 0.0;
  // triangular start: chain u begins at k = j0 + u with the nudged coordinate on top of the unperturbed prefix
#pragma unroll
  for (int kk = 0; kk < FDP_J; ++kk) {
    const int k = j0 + kk;
    if (k < dl) {
      const double ak = a[k], th = theta[k];
#pragma unroll
      for (int u = 0; u < FDP_J; ++u) {
        if (u == kk) s[u] = fma(ak, th + h, pr[k]);
        else if (u < kk) s[u] = fma(ak, th, s[u]);
      }
    }
  }
#pragma unroll 4
  for (int k = j0 + FDP_J; k < dl; ++k) {
    const double ak = a[k], th = theta[k];
#pragma unroll
    for (int u = 0; u < FDP_J; ++u) s[u] = fma(ak, th, s[u]);
  }
  if (row_live) {
#pragma unroll
    for (int u = 0; u < FDP_J; ++u)
      if (j0 + u < dl) pr[j0 + u] = (link_eval(lk.link, s[u] + offr) - f0r) * inv_h;
  }
}

__global__ void __launch_bounds__(128, 2) k_fd_adjoint_prefix(LikDev lk, int d, long long B, const double* __restrict__ X,
                                                           const double* __restrict__ W, const double* __restrict__ scale,
                                                           double* __restrict__ grad) {
  extern __shared__ double sh[];
  constexpr int ROWS = 32;
  const int dl = lk.dl, pitch = dl + 1 + ((dl & 1) ? 0 : 1);   // odd pitch: lanes walking one column hit distinct banks
  double* theta = sh;
  double* At = theta + dl;                 // [ROWS][pitch]
  double* P = At + (size_t)ROWS * pitch;   // [ROWS][pitch]: prefix sums, overwritten by the Jacobian entries
  double* f0 = P + (size_t)ROWS * pitch;   // [ROWS]
  double* ws = f0 + ROWS;                  // [ROWS]
  double* os = ws + ROWS;                  // [ROWS]
  const long long b = blockIdx.x;
  if (b >= B) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int k = threadIdx.x; k < dl; k += blockDim.x) theta[k] = X[b * d + lk.col[k]];
  const double h = lk.differential, inv_h = 1 / h;
  const int nblk = (dl + FDP_J - 1) / FDP_J;
  const int npair = (nblk + 1) / 2;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};    // columns threadIdx.x + 128 m (dl <= 512)
  for (int i0 = 0; i0 < lk.n; i0 += ROWS) {
    const int ni = (lk.n - i0 < ROWS) ? (lk.n - i0) : ROWS;
    __syncthreads();
    for (int r = warp; r < ROWS; r += nwarps) {   // a warp copies whole rows: coalesced, no index division
      const double* src = lk.A + (size_t)(i0 + r) * lk.lds;
      double* dst = At + (size_t)r * pitch;
      for (int k = lane; k < dl; k += 32) dst[k] = (r < ni) ? src[k] : 0.0;
    }
    __syncthreads();
    if (warp == 0) {   // unperturbed chain and its prefixes, one row per lane
      const double* a = At + (size_t)lane * pitch;
      double* pr = P + (size_t)lane * pitch;
      double s0 = 0.0;
      for (int k = 0; k < dl; ++k) {
        pr[k] = s0;
        s0 = fma(a[k], theta[k], s0);
      }
      const bool live = lane < ni;
      const double off = live ? lk.off[i0 + lane] : 0.0;
      os[lane] = off;
      f0[lane] = link_eval(lk.link, s0 + off);
      ws[lane] = live ? W[b * lk.n_pad + i0 + lane] : 0.0;
    }
    __syncthreads();
    const double f0r = f0[lane], offr = os[lane];
    const bool row_live = lane < ni;
    for (int q = warp; q < npair; q += nwarps) {
      fdp_block(lk, dl, pitch, q * FDP_J, lane, row_live, theta, At, P, h, inv_h, f0r, offr);
      const int q2 = nblk - 1 - q;
      if (q2 != q) fdp_block(lk, dl, pitch, q2 * FDP_J, lane, row_live, theta, At, P, h, inv_h, f0r, offr);
    }
    __syncthreads();
#pragma unroll
    for (int mth = 0; mth < 4; ++mth) {
      const int j = threadIdx.x + 128 * mth;
      if (j < dl) {
        double a2 = acc[mth];
        for (int r = 0; r < ni; ++r) a2 = fma(ws[r], P[(size_t)r * pitch + j], a2);   // ascending rows
        acc[mth] = a2;
      }
    }
  }
#pragma unroll
  for (int mth = 0; mth < 4; ++mth) {
    const int j = threadIdx.x + 128 * mth;
    if (j < dl) grad[b * d + lk.col[j]] += scale[b] * acc[mth];
  }
}

// shared memory of k_fd_adjoint_prefix (0: the width is outside its scope, use the plain kernel)
static size_t fd_prefix_smem(int dl) {
  if (dl > 512) return 0;
  const size_t pitch = dl + 1 + ((dl & 1) ? 0 : 1);
  const size_t bytes = ((size_t)dl + 32 * (2 * pitch + 3)) * sizeof(double);
  return bytes <= (size_t)110 * 1024 ? bytes : 0;
}

thy_status launch_finish_generic(const LikDev& lk, long long B, const double* partS, int n_tiles, double* logp, double* scale,
                                 double cauchy_sign, cudaStream_t s) {
  k_finish_generic<<<(unsigned)ceil_div(B, 8), 256, 0, s>>>(lk, B, partS, n_tiles, logp, scale, cauchy_sign);
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

thy_status launch_generic_likelihood(thy_model_s* m, const LikDevOwner& lik, int64_t B, const double* X, double* logp,
                                     double* grad, double cauchy_sign, cudaStream_t s) {
  const LikDev& lk = lik.view;
  const int n_tiles = lk.n_pad / GT + ((lk.n_pad % GT) ? 1 : 0);
  // chunk the batch so that W (and Z) stay below ~1 GiB each
  long long chunk = (long long)((1ull << 30) / ((size_t)lk.n_pad * sizeof(double)));
  chunk = chunk / GT * GT;
  if (chunk < GT) chunk = GT;
  if (chunk > (1 << 20)) chunk = 1 << 20;
  if (chunk > B) chunk = (long long)round_up(B, GT);
  const bool fd = (lk.jacobian == THY_JAC_FINITE_DIFFERENCE);
  const bool want_adj = grad && lk.distribution != THY_DIST_UNIFORM;
  THY_TRY(m->ws.W.reserve((size_t)chunk * lk.n_pad));
  if (fd && want_adj) THY_TRY(m->ws.Z.reserve((size_t)chunk * lk.n_pad));
  THY_TRY(m->ws.partL.reserve((size_t)chunk * n_tiles + (size_t)chunk));
  double* partS = m->ws.partL.ptr;
  double* scale = partS + (size_t)chunk * n_tiles;
  for (long long c0 = 0; c0 < B; c0 += chunk) {
    const long long nb = (B - c0 < chunk) ? (B - c0) : chunk;
    const double* Xc = X + c0 * m->d;
    dim3 gf(n_tiles, (unsigned)ceil_div(nb, GT));
    k_fwd_generic<<<gf, 256, 0, s>>>(lk, m->d, nb, Xc, m->ws.W.ptr, (fd && want_adj) ? m->ws.Z.ptr : nullptr, partS,
                                     n_tiles, (!fd && lk.link != THY_LINK_IDENTITY) ? 1 : 0);
    k_finish_generic<<<(unsigned)ceil_div(nb, 8), 256, 0, s>>>(lk, nb, partS, n_tiles, logp + c0, scale, cauchy_sign);
    count_launch(2);
    if (want_adj) {
      if (fd) {
        const size_t psmem = (m->opt_fd_incremental != 0) ? 0 : fd_prefix_smem(lk.dl);
        if (psmem > 0) {
          THY_CUDA_CHECK(cudaFuncSetAttribute(k_fd_adjoint_prefix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
          k_fd_adjoint_prefix<<<(unsigned)nb, 128, psmem, s>>>(lk, m->d, nb, Xc, m->ws.W.ptr, scale, grad + c0 * m->d);
          count_launch();
          THY_CUDA_CHECK(cudaGetLastError());
          continue;
        }
        dim3 ga((unsigned)nb, (unsigned)ceil_div(lk.dl, 128));
        const size_t fd_smem = (size_t)(lk.dl + 3 * FD_TI) * sizeof(double);
        if (fd_smem > 48 * 1024)
          THY_CUDA_CHECK(cudaFuncSetAttribute(k_fd_adjoint, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fd_smem));
        k_fd_adjoint<<<ga, 128, fd_smem, s>>>(lk, m->d, nb, Xc, m->ws.W.ptr, m->ws.Z.ptr, scale,
                                                            grad + c0 * m->d, m->opt_fd_incremental == 1 ? 1 : 0);
      } else {
        dim3 ga((unsigned)ceil_div(lk.dl, GT), (unsigned)ceil_div(nb, GT));
        k_adj_generic<<<ga, 256, 0, s>>>(lk, m->d, nb, m->ws.W.ptr, scale, grad + c0 * m->d);
      }
      count_launch();
    }
    THY_CUDA_CHECK(cudaGetLastError());
  }
  return THY_OK;
}

}  // namespace thy
