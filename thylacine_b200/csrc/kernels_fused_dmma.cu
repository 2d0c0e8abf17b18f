// Fused linear forward / weighted residual / adjoint kernel on the fp64 tensor cores (DMMA.8x8x4), sm_100a.
//
// What it computes, per chain b (reference: model/components/likelihood/Likelihood.scala:43-65 with
// forwardmodel/LinearForwardModel.scala:103-114 and distributions/GaussianDistribution.scala:52-68 or
// CauchyDistribution.scala:51-79; paths relative to /root/reference/src/main/scala/thylacine/):
//     f = A x_b (+ offset);  r = y - f;  w = r / var;  q_b = sum_i r_i w_i;  G_b = A^T w
//     Gaussian: logp += -q/2 + const,          grad += G
//     Cauchy  : logp += c - (1+n)/2 log(1+q),  grad += -sign (1+n)/(1+q) G   (scalar per chain, so it factors out)
//
// Structure (FlashAttention-like chained GEMMs, A plays both K and V):
//   * one warp owns 8 chains; X^T fragments (8 chains x d) and the G^T accumulators (8 chains x d) live in registers;
//   * A is streamed through shared memory in blocks of MB = 32 observation rows by a dedicated producer warp using
//     1-D TMA bulk copies (cp.async.bulk -> UBLKCP) into a STAGES-deep mbarrier ring;
//   * GEMM1  Z^T[8 x 8 obs] = X^T[8 x d] . A_blk^T    (K = d):   B fragment read straight from the A tile;
//   * epilogue in registers: r, w, q -- the GEMM1 C fragment IS the GEMM2 A fragment (no shuffle, no smem round trip)
//     because the contraction index of GEMM2 may be permuted freely: the 8 observation columns of a C tile are
//     assigned to rows pi = [0,1,2,3,5,4,7,6] of the A tile, which together with a row pitch = 4 (mod 8) doubles makes
//     every fragment load of both GEMMs bank-conflict free;
//   * GEMM2  G^T[8 x d] += W^T[8 x 8 obs] . A_blk     (K = 8 obs, two k4 steps);
//   * persistent grid (one CTA per SM), stream-K split of the (chain tile, row block) space so every SM gets the same
//     number of row blocks; tiles split across CTAs go through partial buffers and are combined, in CTA order, by
//     the last warp to deliver its segment (atomic ticket per (tile, warp)): deterministic, and ONE launch per term.
// Z = A X (n x B, 328 MB at the 100-d / 10k-obs / 4096-chain config) is never materialised.
#include "kernels.h"
#include "philox.cuh"
#include "ptx.cuh"

namespace thy {

namespace {

constexpr int NW = 8;        // consumer warps per CTA (8 chains each)
constexpr int TILE_B = 64;   // chains per tile
constexpr int MB = 32;       // observation rows per pipeline stage
constexpr int JG = MB / 8;   // 8-row groups per stage

__host__ __device__ inline long long unit_begin(long long c, long long U, long long G) { return (c * U) / G; }

struct FusedParams {
  const double* A;        // [n_pad][LDS]
  const double2* yiv;     // [n_pad] (y - offset, 1/var)
  const int* col;         // [dl]
  const double* X;        // [B][d]
  double* logp;           // [B]
  double* grad;           // [B][d] or null
  double* partG;          // [2G][TILE_B][NP]
  double* partL;          // [2G][TILE_B]
  unsigned int* tickets;  // [T][NW] segments of (tile, warp) finished so far; zero between launches
  long long B;
  int d, dl, n, nblk;     // nblk = n_pad / MB
  long long T, U;         // chain tiles, units
  int stages;
  int distribution;
  double log_const, cauchy_sign;
  // fuse_prior: the per-coordinate priors of the same coordinates are evaluated in the output step and the kernel
  // WRITES logp / grad (no prior kernel, no read-modify-write: the outputs may then live in pinned host memory)
  int fuse_prior;
  const int* pr_kind;
  const double* pr_p0;
  const double* pr_p1;
  double pr_log_const;
  // resident variant: 8-row groups actually holding observations (rows beyond n are zero padding and are skipped)
  int nrg;
  int stage_x;   // resident variant: chain rows reach shared memory by cp.async one group ahead (launcher decides)
  int split;     // resident variant: 1, or 2 = two warps share a group's row blocks (small batches; launcher decides)
  // HMC step folded into the output step (needs fuse_prior): see HmcFuse in kernels.h
  HmcFuse hmc;
};

template <int NT>
struct Cfg {
  static constexpr int NP = 8 * NT;       // padded columns
  static constexpr int KS = 2 * NT;       // k4 steps of GEMM1
  static constexpr int LDS = 8 * NT + 4;  // row pitch in doubles: 4 (mod 8) -> conflict-free fragments
  static constexpr int A_BYTES = MB * LDS * 8;
  static constexpr int Y_BYTES = MB * 16;
  static constexpr int STAGE_BYTES = A_BYTES + Y_BYTES;
};

__device__ __forceinline__ double apply_logp(int distribution, double q, int n, double log_const) {
  if (distribution == THY_DIST_GAUSSIAN) return -0.5 * q + log_const;
  return log_const - (1.0 + n) / 2.0 * log(1.0 + q);
}
__device__ __forceinline__ double apply_scale(int distribution, double q, int n, double cauchy_sign) {
  if (distribution == THY_DIST_GAUSSIAN) return 1.0;
  return -cauchy_sign * (1.0 + n) / (1.0 + q);
}

// Output of one chain row with the per-coordinate priors folded in (lane (g, q) of the owning warp holds columns
// 8t + 2q, 8t + 2q + 1 of the gradient).  Every lane of the warp must call this (quad shuffles inside); rows beyond the
// batch are masked by `live`.  The kernel WRITES logp / grad here instead of adding to what the prior kernel left.
template <int NT, bool WANT_GRAD>
__device__ __forceinline__ void emit_row_with_priors(const FusedParams& p, long long chain, bool live, int q4, double qsum,
                                                     const double (&ga)[WANT_GRAD ? NT : 1][2]) {
  const double lik = apply_logp(p.distribution, qsum, p.n, p.log_const);
  const double sc = WANT_GRAD ? apply_scale(p.distribution, qsum, p.n, p.cauchy_sign) : 0.0;
  // fused priors (the term reads the whole flat vector in order: column j IS coordinate j)
  const double* xrow = p.X + chain * p.d;
  const int mode = WANT_GRAD ? p.hmc.mode : 0;
  double* grow = WANT_GRAD ? p.grad + chain * p.d : nullptr;
  double* prow = (WANT_GRAD && mode != HMC_FUSE_NONE) ? p.hmc.p + chain * p.d : nullptr;
  double acc = 0.0, kin = 0.0;
  bool inside = true;
  const double he = -p.hmc.eps / 2;
  // Coordinates are taken in chunks of CH column pairs: all loads of a chunk are issued before anything depends on them
  // (the warps of a CTA reach this step together, so its latency is not hidden by other warps' DMMAs).
  constexpr int CH = 4;
#pragma unroll
  for (int t0 = 0; t0 < NT; t0 += CH) {
    double xv[CH][2], q0[CH][2], q1[CH][2], pm[CH][2];
    int kd[CH][2];
    bool ok[CH][2];
#pragma unroll
    for (int tt = 0; tt < CH; ++tt) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = 8 * (t0 + tt) + 2 * q4 + h;
        ok[tt][h] = live && (t0 + tt < NT) && j < p.dl;
        xv[tt][h] = ok[tt][h] ? xrow[j] : 0.0;
        kd[tt][h] = ok[tt][h] ? p.pr_kind[j] : -1;
        q0[tt][h] = ok[tt][h] ? p.pr_p0[j] : 0.0;
        q1[tt][h] = ok[tt][h] ? p.pr_p1[j] : 0.0;
        pm[tt][h] = (ok[tt][h] && prow) ? prow[j] : 0.0;
      }
    }
#pragma unroll
    for (int tt = 0; tt < CH; ++tt) {
      if (t0 + tt >= NT) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (!ok[tt][h]) continue;
        const int j = 8 * (t0 + tt) + 2 * q4 + h;
        double gj = 0.0;
        if (kd[tt][h] == PK_GAUSS_DIAG) {
          const double dl = q0[tt][h] - xv[tt][h];
          const double w = dl * q1[tt][h];
          acc = fma(dl, w, acc);   // explicit: emit_row_fast must round the same way
          gj = w;
        } else if (kd[tt][h] == PK_UNIFORM) {
          inside = inside && (xv[tt][h] >= q0[tt][h]) && (xv[tt][h] < q1[tt][h]);
        }
        if (WANT_GRAD) {
          gj = fma(sc, ga[WANT_GRAD ? (t0 + tt) : 0][h], gj);
          if (mode == HMC_FUSE_NONE) {
            grow[j] = gj;
          } else if (mode == HMC_FUSE_MID) {
            // second half kick of this leapfrog step, first half kick and drift of the next (HmcmcEngine.scala:55-80;
            // the reference works with -grad logp: p += (-g)(-eps/2), twice, then x += eps p)
            const double kick = __dmul_rn(-gj, he);
            double pv = __dadd_rn(pm[tt][h], kick);
            pv = __dadd_rn(pv, kick);
            prow[j] = pv;
            p.hmc.x_next[chain * p.d + j] = __dadd_rn(xv[tt][h], __dmul_rn(pv, p.hmc.eps));
          } else {
            const double pv = __dadd_rn(pm[tt][h], __dmul_rn(-gj, he));   // last half kick
            kin += pv * pv;
          }
        }
      }
    }
  }
  acc = quad_sum(acc);
  const unsigned ballot = __ballot_sync(0xffffffffu, inside);
  const int lane = threadIdx.x & 31;
  const bool inside_all = ((ballot >> (lane & ~3)) & 0xFu) == 0xFu;
  const double lp_new = (inside_all ? (-0.5 * acc + p.pr_log_const) : -INFINITY) + lik;
  if (mode != HMC_FUSE_LAST) {
    if (live && q4 == 0) p.logp[chain] = lp_new;
    return;
  }
  if (WANT_GRAD) {
    // trajectory end (HmcmcEngine.scala:82-172): H1 = p.p/2 - logp, accept iff dH < 0 || U < exp(-dH); state update
    kin = quad_sum(kin);
    bool accept = false;
    double dH = 0.0;
    if (live) {
      const double H1 = kin / 2.0 + (-lp_new);
      dH = H1 - p.hmc.H0[chain];
      const RngDraw u = rng_uniform2(p.hmc.seed, (uint64_t)chain, 0, RNG_HMC_ACCEPT, *p.hmc.traj);
      accept = (dH < 0) || (u.u0 < exp(-dH));   // NaN dH -> reject, as in the reference
    }
    if (accept) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = 8 * t + 2 * q4 + h;
          if (j < p.dl) {
            // the gradient is recomputed (same operations, same value) rather than kept in 2 NT more registers
            const double xj = xrow[j];
            const double gpr = (p.pr_kind[j] == PK_GAUSS_DIAG) ? (p.pr_p0[j] - xj) * p.pr_p1[j] : 0.0;
            p.hmc.x[chain * p.d + j] = xj;
            p.hmc.g[chain * p.d + j] = fma(sc, ga[WANT_GRAD ? t : 0][h], gpr);
          }
        }
      }
    }
    if (live && q4 == 0) {
      if (accept) {
        p.hmc.lp[chain] = lp_new;
        p.hmc.acceptances[chain] += 1;
        p.hmc.block_acc[chain] += 1;
      }
      p.hmc.attempts[chain] += 1;
      p.hmc.dH[chain] = dH;
    }
  }
}

// Output step of the RESIDENT kernel when nothing but value and gradient leave it (no fused HMC step) and chain rows are
// 16-byte aligned pairs (d even, aligned base pointers).  emit_row_with_priors above costs ~1000 instructions per
// 8-chain group -- prior records fetched from global memory coordinate by coordinate, a reconvergence region per prior
// kind, scalar loads and stores -- and in the resident kernel a group is only 676 DMMAs: the warps spent 42 % of their
// time outside their DMMA loops, so two warps per scheduler were both feeding the pipe a third of the time
// (profiles/r02_resident_output_step.md).  Here the prior records come from shared memory (staged once per CTA as
// (p0, p1, kind)), the X row is re-read as 13 double2 loads issued together, the prior terms are branch-free, and the
// gradient leaves as 16-byte stores.  Same operations in the same order as the general step: bit-identical results.
// MID: the leapfrog step folded in (HMC_FUSE_MID: second half kick of this step, first half kick and drift of the next,
// HmcmcEngine.scala:55-80) -- momenta and next positions leave instead of the gradient, with the very operations of
// emit_row_with_priors (__dmul_rn / __dadd_rn), so replay parity holds for both.
template <int NT, bool MID>
__device__ __forceinline__ void emit_row_fast(const FusedParams& p, const double4* __restrict__ pri, const double2* xrow,
                                              long long chain, bool live, int q4, double qsum, const double (&ga)[NT][2]) {
  const double lik = apply_logp(p.distribution, qsum, p.n, p.log_const);
  const double sc = apply_scale(p.distribution, qsum, p.n, p.cauchy_sign);
  // xrow: the chain's row + q4 (global or staged)
  double2* grow = MID ? nullptr : reinterpret_cast<double2*>(p.grad + chain * p.d) + q4;
  double2* prow = MID ? reinterpret_cast<double2*>(p.hmc.p + chain * p.d) + q4 : nullptr;
  double2* nrow = MID ? reinterpret_cast<double2*>(p.hmc.x_next + chain * p.d) + q4 : nullptr;
  const double he = -p.hmc.eps / 2;
  double2 xv[NT], pm[MID ? NT : 1];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const bool ok = live && (8 * t + 2 * q4) < p.dl;
    xv[t] = ok ? xrow[4 * t] : make_double2(0.0, 0.0);
    if (MID) pm[t] = ok ? prow[4 * t] : make_double2(0.0, 0.0);
  }
  double acc = 0.0;
  int outside = 0;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int j = 8 * t + 2 * q4;
    const bool ok = live && j < p.dl;
    const double4 k0 = pri[j], k1 = pri[j + 1];   // beyond dl: kind = -1, no term
    const double d0 = k0.x - xv[t].x, d1 = k1.x - xv[t].y;
    const double w0 = d0 * k0.y, w1 = d1 * k1.y;
    const bool g0 = k0.z == (double)PK_GAUSS_DIAG, g1 = k1.z == (double)PK_GAUSS_DIAG;
    acc = g0 ? fma(d0, w0, acc) : acc;   // the general step's fma, coordinate by coordinate
    acc = g1 ? fma(d1, w1, acc) : acc;
    outside |= (int)(k0.z == (double)PK_UNIFORM) & ((int)!(xv[t].x >= k0.x) | (int)!(xv[t].x < k0.y));
    outside |= (int)(k1.z == (double)PK_UNIFORM) & ((int)!(xv[t].y >= k1.x) | (int)!(xv[t].y < k1.y));
    const double gj0 = fma(sc, ga[t][0], g0 ? w0 : 0.0), gj1 = fma(sc, ga[t][1], g1 ? w1 : 0.0);
    if (!MID) {
      if (ok) grow[4 * t] = make_double2(gj0, gj1);
    } else if (ok) {
      const double k0k = __dmul_rn(-gj0, he), k1k = __dmul_rn(-gj1, he);
      const double pv0 = __dadd_rn(__dadd_rn(pm[MID ? t : 0].x, k0k), k0k), pv1 = __dadd_rn(__dadd_rn(pm[MID ? t : 0].y, k1k), k1k);
      prow[4 * t] = make_double2(pv0, pv1);
      nrow[4 * t] = make_double2(__dadd_rn(xv[t].x, __dmul_rn(pv0, p.hmc.eps)), __dadd_rn(xv[t].y, __dmul_rn(pv1, p.hmc.eps)));
    }
  }
  acc = quad_sum(acc);
  const unsigned ballot = __ballot_sync(0xffffffffu, outside == 0);
  const int lane = threadIdx.x & 31;
  const bool inside_all = ((ballot >> (lane & ~3)) & 0xFu) == 0xFu;
  if (live && q4 == 0) p.logp[chain] = (inside_all ? (-0.5 * acc + p.pr_log_const) : -INFINITY) + lik;
}

// Stream-K combine.  The last warp to deliver its segment of a (tile, warp) row group sums all segments, always in CTA
// order, so the result does not depend on which CTA happens to finish last (run-to-run bit-identical).  Kept out of
// line: it runs at most twice per CTA and must not cost the main loop any registers.
template <int NT, bool WANT_GRAD, bool FUSE_PRIOR>
__device__ __noinline__ void combine_segments(const FusedParams& p, long long tile, int warp, int lane, long long G) {
  using C = Cfg<NT>;
  const int g8 = lane >> 2, q4 = lane & 3;
  const int row = warp * 8 + g8;
  const long long chain = tile * TILE_B + row;
  const long long tb = tile * p.nblk, te = tb + p.nblk;
  long long c_lo = (tb * G) / p.U;
  while (c_lo > 0 && unit_begin(c_lo, p.U, G) > tb) --c_lo;
  while (unit_begin(c_lo + 1, p.U, G) <= tb) ++c_lo;
  long long c_hi = ((te - 1) * G) / p.U;
  while (c_hi > 0 && unit_begin(c_hi, p.U, G) > te - 1) --c_hi;
  while (unit_begin(c_hi + 1, p.U, G) <= te - 1) ++c_hi;
  unsigned int nseg = 0;
  for (long long cc = c_lo; cc <= c_hi; ++cc) nseg += (unit_begin(cc, p.U, G) != unit_begin(cc + 1, p.U, G)) ? 1u : 0u;
  __threadfence();
  __syncwarp();
  unsigned int prev = 0;
  if (lane == 0) prev = atomicAdd(&p.tickets[tile * NW + warp], 1u);
  prev = __shfl_sync(0xffffffffu, prev, 0);
  if (prev != nseg - 1) return;
  __threadfence();
  double qtot = 0.0;
  double gs[WANT_GRAD ? NT : 1][2];
  if (WANT_GRAD) {
#pragma unroll
    for (int t = 0; t < NT; ++t) gs[t][0] = gs[t][1] = 0.0;
  }
  for (long long cc = c_lo; cc <= c_hi; ++cc) {
    const long long cb = unit_begin(cc, p.U, G);
    if (cb == unit_begin(cc + 1, p.U, G)) continue;   // CTA without work wrote nothing
    const long long sl = 2 * cc + ((cb >= tb) ? 0 : 1);
    qtot += __ldcg(&p.partL[sl * TILE_B + row]);
    if (WANT_GRAD) {
      const double* prow = p.partG + ((size_t)sl * TILE_B + row) * C::NP;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const double2 v = __ldcg(reinterpret_cast<const double2*>(prow + 8 * t + 2 * q4));
        gs[WANT_GRAD ? t : 0][0] += v.x;
        gs[WANT_GRAD ? t : 0][1] += v.y;
      }
    }
  }
  if (FUSE_PRIOR) {
    emit_row_with_priors<NT, WANT_GRAD>(p, chain, chain < p.B, q4, qtot, gs);
  } else if (chain < p.B) {
    if (q4 == 0) p.logp[chain] += apply_logp(p.distribution, qtot, p.n, p.log_const);
    if (WANT_GRAD) {
      const double sc = apply_scale(p.distribution, qtot, p.n, p.cauchy_sign);
      double* grow = p.grad + chain * p.d;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int j = 8 * t + 2 * q4;
        if (j < p.dl) grow[p.col[j]] += sc * gs[WANT_GRAD ? t : 0][0];
        if (j + 1 < p.dl) grow[p.col[j + 1]] += sc * gs[WANT_GRAD ? t : 0][1];
      }
    }
  }
  if (lane == 0) p.tickets[tile * NW + warp] = 0u;   // re-arm for the next launch
}

template <int NT, bool WANT_GRAD, bool FUSE_PRIOR>
__global__ void __launch_bounds__((NW + 1) * 32, 1) k_fused_dmma(const FusedParams p) {
  using C = Cfg<NT>;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + p.stages;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long G = gridDim.x;
  const long long c = blockIdx.x;
  const long long u0 = unit_begin(c, p.U, G), u1 = unit_begin(c + 1, p.U, G);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], NW);
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  __syncthreads();

  if (warp == NW) {
    // ===== producer warp: one lane drives the TMA engine =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long u = u0; u < u1; ++u) {
        const int blk = (int)(u % p.nblk);
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        unsigned char* dst = smem + (size_t)stage * C::STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
        ptx::bulk_g2s(dst, p.A + (size_t)blk * MB * C::LDS, C::A_BYTES, &full_bar[stage]);
        ptx::bulk_g2s(dst + C::A_BYTES, p.yiv + (size_t)blk * MB, C::Y_BYTES, &full_bar[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    return;
  }

  // ===== consumer warps =====
  const int g8 = lane >> 2;  // row within the 8-chain group / column within an 8-wide B fragment
  const int q4 = lane & 3;   // k index inside a k4 step
  // observation-row permutation pi = [0,1,2,3,5,4,7,6]
  const int pi_g = (g8 < 4) ? g8 : (g8 ^ 1);
  const int pi_s0 = (q4 < 2) ? (2 * q4) : (2 * q4 + 1);      // pi(2q)   = 0,2,5,7
  const int pi_s1 = (q4 < 2) ? (2 * q4 + 1) : (2 * q4);      // pi(2q+1) = 1,3,4,6
  const int off1 = pi_g * C::LDS + q4;                        // GEMM1 B fragment: A_s[pi(g)][4kk + q]
  const int off2_s0 = pi_s0 * C::LDS + g8;                    // GEMM2 B fragment: A_s[pi(2q+s)][8g' + g]
  const int off2_s1 = pi_s1 * C::LDS + g8;

  double xf[C::KS];
  double ga[WANT_GRAD ? NT : 1][2];
  double ssq = 0.0;
  long long cur_tile = -1;
  long long seg_begin = 0;
  int stage = 0;
  uint32_t phase = 0;

  auto flush = [&](long long tile, long long ub, long long ue) {
    // segment [ub, ue) of `tile` is complete
    const double qsum = quad_sum(ssq);
    const long long chain = tile * TILE_B + warp * 8 + g8;
    const bool full_tile = (ue - ub) == p.nblk;
    if (full_tile) {
      if (FUSE_PRIOR) {
        emit_row_with_priors<NT, WANT_GRAD>(p, chain, chain < p.B, q4, qsum, ga);
      } else if (chain < p.B) {
        if (q4 == 0) p.logp[chain] += apply_logp(p.distribution, qsum, p.n, p.log_const);
        if (WANT_GRAD) {
          const double sc = apply_scale(p.distribution, qsum, p.n, p.cauchy_sign);
          double* grow = p.grad + chain * p.d;
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            const int j = 8 * t + 2 * q4;
            if (j < p.dl) grow[p.col[j]] += sc * ga[WANT_GRAD ? t : 0][0];
            if (j + 1 < p.dl) grow[p.col[j + 1]] += sc * ga[WANT_GRAD ? t : 0][1];
          }
        }
      }
    } else {
      // partial segment: slot 2c for the CTA's first segment, 2c+1 otherwise
      const long long slot = 2 * c + ((ub == u0) ? 0 : 1);
      const int row = warp * 8 + g8;
      if (q4 == 0) p.partL[slot * TILE_B + row] = qsum;
      if (WANT_GRAD) {
        double* prow = p.partG + ((size_t)slot * TILE_B + row) * C::NP;
#pragma unroll
        for (int t = 0; t < NT; ++t)
          *reinterpret_cast<double2*>(prow + 8 * t + 2 * q4) = make_double2(ga[WANT_GRAD ? t : 0][0], ga[WANT_GRAD ? t : 0][1]);
      }
      combine_segments<NT, WANT_GRAD, FUSE_PRIOR>(p, tile, warp, lane, G);
    }
  };

  for (long long u = u0; u < u1; ++u) {
    const long long tile = u / p.nblk;
    if (tile != cur_tile) {
      if (cur_tile >= 0) flush(cur_tile, seg_begin, u);
      cur_tile = tile;
      seg_begin = u;
      ssq = 0.0;
      const long long chain = tile * TILE_B + warp * 8 + g8;
#pragma unroll
      for (int kk = 0; kk < C::KS; ++kk) {
        const int k = 4 * kk + q4;
        xf[kk] = (chain < p.B && k < p.dl) ? p.X[chain * p.d + (FUSE_PRIOR ? k : p.col[k])] : 0.0;
      }
      if (WANT_GRAD) {
#pragma unroll
        for (int t = 0; t < NT; ++t) ga[t][0] = ga[t][1] = 0.0;
      }
    }

    ptx::mbar_wait(&full_bar[stage], phase);
    const double* As = reinterpret_cast<const double*>(smem + (size_t)stage * C::STAGE_BYTES);
    const double2* Ys = reinterpret_cast<const double2*>(smem + (size_t)stage * C::STAGE_BYTES + C::A_BYTES);

    // ---- GEMM1: Z^T = X^T A_blk^T, JG independent accumulator chains ----
    double z[JG][2];
#pragma unroll
    for (int j = 0; j < JG; ++j) z[j][0] = z[j][1] = 0.0;
#pragma unroll
    for (int kk = 0; kk < C::KS; ++kk) {
#pragma unroll
      for (int j = 0; j < JG; ++j) {
        const double bfrag = As[off1 + j * 8 * C::LDS + 4 * kk];
        ptx::dmma884(z[j][0], z[j][1], xf[kk], bfrag);
      }
    }
    // ---- epilogue: residual, weight, quadratic form; C fragment -> A fragment of GEMM2 ----
#pragma unroll
    for (int j = 0; j < JG; ++j) {
      const double2 y0 = Ys[j * 8 + pi_s0];
      const double2 y1 = Ys[j * 8 + pi_s1];
      const double r0 = y0.x - z[j][0];
      const double r1 = y1.x - z[j][1];
      const double w0 = r0 * y0.y;
      const double w1 = r1 * y1.y;
      ssq = fma(r0, w0, ssq);
      ssq = fma(r1, w1, ssq);
      z[j][0] = w0;
      z[j][1] = w1;
    }
    // ---- GEMM2: G^T += W^T A_blk ----
    if (WANT_GRAD) {
#pragma unroll
      for (int j = 0; j < JG; ++j) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const double b0 = As[off2_s0 + j * 8 * C::LDS + 8 * t];
          ptx::dmma884(ga[t][0], ga[t][1], z[j][0], b0);
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const double b1 = As[off2_s1 + j * 8 * C::LDS + 8 * t];
          ptx::dmma884(ga[t][0], ga[t][1], z[j][1], b1);
        }
      }
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);
    if (++stage == p.stages) {
      stage = 0;
      phase ^= 1;
    }
  }
  if (cur_tile >= 0) flush(cur_tile, seg_begin, u1);
}


// ---- resident variant: the whole coefficient matrix stays in shared memory for the CTA's lifetime -------------------------
// For models whose A fits one SM's shared memory (a collapsed posterior: d x d; BASELINE config 1: 1000 x 10) streaming A
// through a ring per (tile, row block) unit is the wrong shape: every unit pays a barrier round trip, the X-fragment
// prologue and the output step for a few row blocks of work, and 4096 chains are only 64 tiles for 148 SMs.  Here each
// CTA loads A once (TMA bulk copies on one mbarrier, overlapped with the first X-fragment loads), and its warps then work
// through 8-chain groups independently -- no CTA-wide synchronisation after the load, so one warp's prologue / output step
// overlaps its neighbours' DMMAs.  Work is distributed by WARP (group = 8 chains), and the CTA width is chosen by the
// launcher so that small batches still cover all SMs (4096 chains = 512 groups -> 128 CTAs of 4 warps).  Row groups
// beyond the last observation are skipped (n' = 100 pads to 128 rows: 13 groups instead of 16).  Same fragment scheme,
// same arithmetic per row as the streaming kernel above.
template <int NT, int JGN, bool WANT_GRAD>
__device__ __forceinline__ void resident_rows(const double* __restrict__ As, const double2* __restrict__ Ys, int off1,
                                              int off2_s0, int off2_s1, int pi_s0, int pi_s1, const double (&xf)[2 * NT],
                                              double (&ga)[WANT_GRAD ? NT : 1][2], double& ssq) {
  using C = Cfg<NT>;
  double z[JGN][2];
#pragma unroll
  for (int j = 0; j < JGN; ++j) z[j][0] = z[j][1] = 0.0;
#pragma unroll
  for (int kk = 0; kk < C::KS; ++kk) {
#pragma unroll
    for (int j = 0; j < JGN; ++j) ptx::dmma884(z[j][0], z[j][1], xf[kk], As[off1 + j * 8 * C::LDS + 4 * kk]);
  }
#pragma unroll
  for (int j = 0; j < JGN; ++j) {
    const double2 y0 = Ys[j * 8 + pi_s0];
    const double2 y1 = Ys[j * 8 + pi_s1];
    const double r0 = y0.x - z[j][0];
    const double r1 = y1.x - z[j][1];
    const double w0 = r0 * y0.y;
    const double w1 = r1 * y1.y;
    ssq = fma(r0, w0, ssq);
    ssq = fma(r1, w1, ssq);
    z[j][0] = w0;
    z[j][1] = w1;
  }
  if (WANT_GRAD) {
#pragma unroll
    for (int j = 0; j < JGN; ++j) {
#pragma unroll
      for (int t = 0; t < NT; ++t) ptx::dmma884(ga[t][0], ga[t][1], z[j][0], As[off2_s0 + j * 8 * C::LDS + 8 * t]);
#pragma unroll
      for (int t = 0; t < NT; ++t) ptx::dmma884(ga[t][0], ga[t][1], z[j][1], As[off2_s1 + j * 8 * C::LDS + 8 * t]);
    }
  }
}

// (12 warps per CTA at 170 registers each were measured too: no faster at 65 536 chains, slower at 16 384.)
template <int NT, bool WANT_GRAD, bool FUSE_PRIOR>
__global__ void __launch_bounds__(NW * 32, 1) k_resident_dmma(const FusedParams p) {
  using C = Cfg<NT>;
  extern __shared__ __align__(128) unsigned char smem[];
  const int rows = p.nrg * 8;
  double* As = reinterpret_cast<double*>(smem);
  double2* Ys = reinterpret_cast<double2*>(smem + (size_t)rows * C::LDS * 8);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)rows * (C::LDS * 8 + 16));
  double4* pri = reinterpret_cast<double4*>(smem + (size_t)rows * (C::LDS * 8 + 16) + 16);   // [8 NT + 1] (p0, p1, kind, -)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  // fast output step (emit_row_fast): value + gradient only, rows of X / grad made of 16-byte aligned pairs
  const bool fast_mid = p.hmc.mode == HMC_FUSE_MID;
  const bool fast_out = FUSE_PRIOR && WANT_GRAD && (p.d & 1) == 0 && (p.dl & 1) == 0 &&
                        (reinterpret_cast<uintptr_t>(p.X) & 15) == 0 &&
                        (fast_mid ? ((reinterpret_cast<uintptr_t>(p.hmc.p) | reinterpret_cast<uintptr_t>(p.hmc.x_next)) & 15) == 0
                                  : (p.hmc.mode == HMC_FUSE_NONE && (reinterpret_cast<uintptr_t>(p.grad) & 15) == 0));
  if (FUSE_PRIOR) {
    for (int j = threadIdx.x; j <= 8 * NT; j += blockDim.x) {
      const bool in = j < p.dl;
      pri[j] = make_double4(in ? p.pr_p0[j] : 0.0, in ? p.pr_p1[j] : 0.0, in ? (double)p.pr_kind[j] : -1.0, 0.0);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(bar, (uint32_t)rows * (C::LDS * 8 + 16));
    for (int r0 = 0; r0 < rows; r0 += 64) {   // <= 64 rows (67 KB) per bulk copy
      const int nr = (rows - r0 < 64) ? (rows - r0) : 64;
      ptx::bulk_g2s(As + (size_t)r0 * C::LDS, p.A + (size_t)r0 * C::LDS, (uint32_t)nr * C::LDS * 8, bar);
    }
    ptx::bulk_g2s(Ys, p.yiv, (uint32_t)rows * 16, bar);
  }

  const int g8 = lane >> 2, q4 = lane & 3;
  const int pi_g = (g8 < 4) ? g8 : (g8 ^ 1);
  const int pi_s0 = (q4 < 2) ? (2 * q4) : (2 * q4 + 1);
  const int pi_s1 = (q4 < 2) ? (2 * q4 + 1) : (2 * q4);
  const int off1 = pi_g * C::LDS + q4;
  const int off2_s0 = pi_s0 * C::LDS + g8;
  const int off2_s1 = pi_s1 * C::LDS + g8;
  const long long ngroups = (p.B + 7) / 8;
  bool loaded = false;
  // Chain rows one group ahead (stage_x): the 8 rows of a group are one contiguous block of X; the warp copies the block
  // of its NEXT group into its own double buffer with 16-byte cp.async while it works on the current one, so neither the
  // X fragments nor the output step's re-read of the row wait for global memory.
  const bool stage = FUSE_PRIOR && p.stage_x != 0;
  double* xw = reinterpret_cast<double*>(pri + 8 * NT + 1) + (size_t)warp * 2 * 8 * p.d;
  auto issue_rows = [&](long long g, int b) {
    if (g < ngroups) {
      const long long c0 = g * 8;
      const int nrows = (int)((p.B - c0 < 8) ? (p.B - c0) : 8);
      const int chunks = nrows * p.d / 2;
      const char* src = reinterpret_cast<const char*>(p.X + c0 * p.d);
      const uint32_t dst = ptx::smem_u32(xw + (size_t)b * 8 * p.d);
      for (int c = lane; c < chunks; c += 32) ptx::cp_async16(dst + 16u * c, src + 16 * (size_t)c);
    }
    ptx::cp_async_commit();   // an empty group keeps the wait count uniform
  };
  // Small batches (split == 2): a group is the work of a PAIR of warps.  With one group per warp and at most four warps on
  // an SM every warp is alone on its scheduler and its 676 DMMAs issue 26 clocks apart (one dependent chain: 62 % of the
  // pipe); two warps that take alternate row blocks of the same group -- both GEMMs of a row block stay in one warp --
  // finish in 338 x 32 clocks, and the odd warp hands its partial gradient and sum of squares to the even one through
  // shared memory (one named barrier per group, buffers alternate with the iteration).
  const int split = p.split == 2 ? 2 : 1;
  const int slot = split == 2 ? (warp >> 1) : warp, sw = split == 2 ? (warp & 1) : 0, slots = nwarps / split;
  double2* exch = reinterpret_cast<double2*>(pri + 8 * NT + 1) + (size_t)slot * 2 * (NT + 1) * 32 + lane;
  const long long gfirst = (long long)blockIdx.x * slots + slot, gstep = (long long)gridDim.x * slots;
  if (stage) issue_rows(gfirst, 0);
  int buf = 0;
  for (long long grp = gfirst; grp < ngroups; grp += gstep, buf ^= 1) {
    const long long chain = grp * 8 + g8;
    double xf[C::KS];
    const double* xsrow = xw + (size_t)buf * 8 * p.d + (size_t)g8 * p.d;
    if (stage) {
      issue_rows(grp + gstep, buf ^ 1);
      ptx::cp_async_wait_group<1>();
      __syncwarp();
#pragma unroll
      for (int kk = 0; kk < C::KS; ++kk) {
        const int k = 4 * kk + q4;
        xf[kk] = (chain < p.B && k < p.dl) ? xsrow[k] : 0.0;
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < C::KS; ++kk) {
        const int k = 4 * kk + q4;
        xf[kk] = (chain < p.B && k < p.dl) ? p.X[chain * p.d + (FUSE_PRIOR ? k : p.col[k])] : 0.0;
      }
    }
    if (!loaded) {
      ptx::mbar_wait(bar, 0);
      loaded = true;
    }
    double ga[WANT_GRAD ? NT : 1][2];
    if (WANT_GRAD) {
#pragma unroll
      for (int t = 0; t < NT; ++t) ga[t][0] = ga[t][1] = 0.0;
    }
    double ssq = 0.0;
    // row groups in balanced blocks of 4 / 3 (13 groups = 4 + 3 + 3 + 3, never 4 + 4 + 4 + 1: a block's groups are the
    // independent accumulator chains of GEMM1, and a single chain leaves the DMMA pipe waiting on its own latency)
    const int nblocks = (p.nrg + JG - 1) / JG, base = p.nrg / nblocks, extra = p.nrg % nblocks;
    int rg = 0;
    for (int blk = 0; blk < nblocks; ++blk) {
      const int cnt = base + (blk < extra ? 1 : 0);
      if (split == 2 && (blk & 1) != sw) {   // the other warp of the pair takes this block
        rg += cnt;
        continue;
      }
      const double* Ab = As + (size_t)rg * 8 * C::LDS;
      const double2* Yb = Ys + rg * 8;
      if (cnt == 4) resident_rows<NT, 4, WANT_GRAD>(Ab, Yb, off1, off2_s0, off2_s1, pi_s0, pi_s1, xf, ga, ssq);
      else if (cnt == 3) resident_rows<NT, 3, WANT_GRAD>(Ab, Yb, off1, off2_s0, off2_s1, pi_s0, pi_s1, xf, ga, ssq);
      else if (cnt == 2) resident_rows<NT, 2, WANT_GRAD>(Ab, Yb, off1, off2_s0, off2_s1, pi_s0, pi_s1, xf, ga, ssq);
      else resident_rows<NT, 1, WANT_GRAD>(Ab, Yb, off1, off2_s0, off2_s1, pi_s0, pi_s1, xf, ga, ssq);
      rg += cnt;
    }
    if (split == 2) {
      double2* ex = exch + (size_t)buf * (NT + 1) * 32;
      if (sw == 1) {
        if (WANT_GRAD) {
#pragma unroll
          for (int t = 0; t < NT; ++t) ex[t * 32] = make_double2(ga[WANT_GRAD ? t : 0][0], ga[WANT_GRAD ? t : 0][1]);
        }
        ex[NT * 32] = make_double2(ssq, 0.0);
      }
      asm volatile("bar.sync %0, 64;" ::"r"(1 + slot) : "memory");
      if (sw == 1) continue;   // the even warp finishes the group
      if (WANT_GRAD) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const double2 v = ex[t * 32];
          ga[WANT_GRAD ? t : 0][0] += v.x;
          ga[WANT_GRAD ? t : 0][1] += v.y;
        }
      }
      ssq += ex[NT * 32].x;
    }
    const double qsum = quad_sum(ssq);
    if (FUSE_PRIOR) {
      if (WANT_GRAD && fast_out) {
        const double2* xrow = reinterpret_cast<const double2*>(stage ? xsrow : p.X + chain * p.d) + q4;
        if (fast_mid) emit_row_fast<WANT_GRAD ? NT : 1, true>(p, pri, xrow, chain, chain < p.B, q4, qsum, ga);
        else emit_row_fast<WANT_GRAD ? NT : 1, false>(p, pri, xrow, chain, chain < p.B, q4, qsum, ga);
      } else {
        emit_row_with_priors<NT, WANT_GRAD>(p, chain, chain < p.B, q4, qsum, ga);
      }
      if (stage) __syncwarp();   // every lane is done with this buffer before the next iteration refills it
    } else if (chain < p.B) {
      if (q4 == 0) p.logp[chain] += apply_logp(p.distribution, qsum, p.n, p.log_const);
      if (WANT_GRAD) {
        const double sc = apply_scale(p.distribution, qsum, p.n, p.cauchy_sign);
        double* grow = p.grad + chain * p.d;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int j = 8 * t + 2 * q4;
          if (j < p.dl) grow[p.col[j]] += sc * ga[WANT_GRAD ? t : 0][0];
          if (j + 1 < p.dl) grow[p.col[j + 1]] += sc * ga[WANT_GRAD ? t : 0][1];
        }
      }
    }
  }
  // a CTA must not retire while its bulk copies are in flight: warp 0 always sees them land
  if (warp == 0 && !loaded) ptx::mbar_wait(bar, 0);
}

static size_t resident_smem_bytes(int nrg, int lds) {   // A rows + (y, 1/var) + mbarrier + prior records (lds = 8 NT + 4)
  return (size_t)nrg * 8 * ((size_t)lds * 8 + 16) + 16 + (size_t)(lds - 3) * 32;
}

template <int NT>
thy_status launch_resident_nt(thy_model_s* m, FusedParams& p, bool want_grad, cudaStream_t s) {
  size_t smem = resident_smem_bytes(p.nrg, Cfg<NT>::LDS);
  const long long ngroups = (p.B + 7) / 8;
  const int sms = sm_count();
  // CTA width: enough warps to hold the work, few enough that small batches still reach every SM.  When that leaves at
  // most four warps on an SM (one per scheduler), a group becomes the work of a pair of warps (p.split = 2).
  int warps = (int)((ngroups + sms - 1) / sms);
  warps = warps < 1 ? 1 : (warps > NW ? NW : warps);
  if (warps == 3) warps = 4;
  if (warps > 4) warps = NW;
  const size_t exch_bytes = (size_t)(NW / 2) * 2 * (Cfg<NT>::KS / 2 + 1) * 32 * 16;
  p.split = (2 * warps <= NW && p.nrg > JG && smem + exch_bytes <= (size_t)224 * 1024) ? 2 : 1;
  long long grid = (ngroups + warps - 1) / warps;
  if (grid > sms) grid = sms;
  if (p.split == 2) {
    warps *= 2;   // the same groups per CTA, two warps each
    smem += exch_bytes;
  }
  // chain rows staged one group ahead when there is room for two 8-row buffers per warp and rows are 16-byte pairs
  const size_t xs_bytes = (size_t)warps * 2 * 8 * p.d * sizeof(double);
  p.stage_x = (p.split == 1 && p.fuse_prior && (p.d & 1) == 0 && (reinterpret_cast<uintptr_t>(p.X) & 15) == 0 &&
               ngroups > 2 * grid * warps && smem + xs_bytes <= (size_t)224 * 1024) ? 1 : 0;   // pays from the third group on
  if (p.stage_x) smem += xs_bytes;
  KernelTimer timer(m, s);
  auto launch = [&](auto kernel) -> thy_status {
    THY_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<(unsigned)grid, warps * 32, smem, s>>>(p);
    return THY_OK;
  };
  if (p.fuse_prior) {
    THY_TRY(want_grad ? launch(k_resident_dmma<NT, true, true>) : launch(k_resident_dmma<NT, false, true>));
  } else {
    THY_TRY(want_grad ? launch(k_resident_dmma<NT, true, false>) : launch(k_resident_dmma<NT, false, false>));
  }
  timer.stop();
  count_launch();
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

template <int NT>
thy_status launch_nt(thy_model_s* m, FusedParams& p, bool want_grad, cudaStream_t s) {
  using C = Cfg<NT>;
  const int G = sm_count();
  int stages = (200 * 1024) / C::STAGE_BYTES;
  if (stages > 6) stages = 6;
  if (stages < 2) return fail(THY_ERR_UNSUPPORTED, "fused DMMA: tile does not fit shared memory");
  p.stages = stages;
  const size_t smem = (size_t)stages * C::STAGE_BYTES + 2 * stages * sizeof(uint64_t);
  THY_TRY(m->ws.partG.reserve((size_t)2 * G * TILE_B * C::NP));
  THY_TRY(m->ws.partL.reserve((size_t)2 * G * TILE_B));
  p.partG = m->ws.partG.ptr;
  p.partL = m->ws.partL.ptr;
  const size_t n_tickets = (size_t)p.T * NW;
  if (m->ws.fused_tickets.count < n_tickets) {
    THY_TRY(m->ws.fused_tickets.reserve(n_tickets));
    THY_CUDA_CHECK(cudaMemsetAsync(m->ws.fused_tickets.ptr, 0, m->ws.fused_tickets.count * sizeof(unsigned int), s));
  }
  p.tickets = m->ws.fused_tickets.ptr;
  KernelTimer timer(m, s);
  auto launch = [&](auto kernel) -> thy_status {
    THY_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<G, (NW + 1) * 32, smem, s>>>(p);
    return THY_OK;
  };
  if (p.fuse_prior) {
    THY_TRY(want_grad ? launch(k_fused_dmma<NT, true, true>) : launch(k_fused_dmma<NT, false, true>));
  } else {
    THY_TRY(want_grad ? launch(k_fused_dmma<NT, true, false>) : launch(k_fused_dmma<NT, false, false>));
  }
  timer.stop();
  count_launch();
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

}  // namespace

int fused_nt_for(int dl) {
  static const int list[] = {1, 2, 3, 4, 5, 7, 10, 13, 16};
  for (int nt : list)
    if (8 * nt >= dl) return nt;
  return 0;
}

bool fused_dmma_supported(const LikDev& lik, bool want_grad) {
  (void)want_grad;
  const int nt = fused_nt_for(lik.dl);
  return nt > 0 && lik.lds == 8 * nt + 4 && lik.n_pad % MB == 0;
}

// Auto-selection cost model (seconds): whole 64-chain tiles at ~30 TFLOP/s, plus launch / combine overhead.
double fused_cost_model(const LikDev& lik, int64_t B) {
  const double flops = 4.0 * lik.n_pad * (lik.lds - 4) * 64.0 * (double)((B + 63) / 64);
  const double t = flops / 30e12;
  return 5e-6 + (t > 6e-6 ? t : 6e-6);
}

bool fused_prior_supported(const thy_model_s* m, const LikDevOwner& lik) {
  return m->lik_dev.size() == 1 && m->prior_view.n_special == 0 && lik.view.dl == m->d && lik.contiguous &&
         !lik.col_host.empty() && lik.col_host[0] == 0;
}

// Resident variant: A (rows that hold observations) + (y, 1/var) fit one CTA's shared memory.  Preferred whenever the
// batch gives every SM at least a few groups, or the model is so short that streaming it makes no sense.
bool resident_dmma_supported(const LikDev& lik, int64_t B) {
  if (!fused_dmma_supported(lik, true)) return false;
  const int nrg = (lik.n + 7) / 8;
  if (resident_smem_bytes(nrg, lik.lds) > (size_t)200 * 1024) return false;
  return lik.n_pad <= 256 || (B + 7) / 8 >= sm_count() / 2;
}

thy_status launch_fused_dmma(thy_model_s* m, const LikDevOwner& lik, int64_t B, const double* X, double* logp, double* grad,
                             double cauchy_sign, cudaStream_t s, bool fuse_prior, const HmcFuse* hmc) {
  const LikDev& lk = lik.view;
  FusedParams p{};
  if (hmc) {
    if (!fuse_prior || !grad) return fail(THY_ERR_STATE, "the fused HMC step needs the fused-prior output step and a gradient");
    p.hmc = *hmc;
  }
  if (fuse_prior) {
    if (!fused_prior_supported(m, lik)) return fail(THY_ERR_STATE, "fused priors requested for an unsupported model");
    p.fuse_prior = 1;
    p.pr_kind = m->prior_view.kind;
    p.pr_p0 = m->prior_view.p0;
    p.pr_p1 = m->prior_view.p1;
    p.pr_log_const = m->prior_view.log_const;
  }
  p.A = lk.A;
  p.yiv = lk.yiv_lin;
  p.col = lk.col;
  p.X = X;
  p.logp = logp;
  p.grad = grad;
  p.B = B;
  p.d = m->d;
  p.dl = lk.dl;
  p.n = lk.n;
  p.nblk = lk.n_pad / MB;
  p.T = (B + TILE_B - 1) / TILE_B;
  p.U = p.T * p.nblk;
  p.distribution = lk.distribution;
  p.log_const = lk.log_const;
  p.cauchy_sign = cauchy_sign;
  p.nrg = (lk.n + 7) / 8;
  const bool wg = grad != nullptr;
  if (m->opt_kernel_path != 5 && resident_dmma_supported(lk, B)) {   // 5: the streaming variant is asked for explicitly
    m->last_path = "resident_dmma";
    switch (fused_nt_for(lk.dl)) {
      case 1: return launch_resident_nt<1>(m, p, wg, s);
      case 2: return launch_resident_nt<2>(m, p, wg, s);
      case 3: return launch_resident_nt<3>(m, p, wg, s);
      case 4: return launch_resident_nt<4>(m, p, wg, s);
      case 5: return launch_resident_nt<5>(m, p, wg, s);
      case 7: return launch_resident_nt<7>(m, p, wg, s);
      case 10: return launch_resident_nt<10>(m, p, wg, s);
      case 13: return launch_resident_nt<13>(m, p, wg, s);
      case 16: return launch_resident_nt<16>(m, p, wg, s);
      default: return fail(THY_ERR_UNSUPPORTED, "resident DMMA: unsupported width");
    }
  }
  switch (fused_nt_for(lk.dl)) {
    case 1: return launch_nt<1>(m, p, wg, s);
    case 2: return launch_nt<2>(m, p, wg, s);
    case 3: return launch_nt<3>(m, p, wg, s);
    case 4: return launch_nt<4>(m, p, wg, s);
    case 5: return launch_nt<5>(m, p, wg, s);
    case 7: return launch_nt<7>(m, p, wg, s);
    case 10: return launch_nt<10>(m, p, wg, s);
    case 13: return launch_nt<13>(m, p, wg, s);
    case 16: return launch_nt<16>(m, p, wg, s);
    default: return fail(THY_ERR_UNSUPPORTED, "fused DMMA: unsupported width");
  }
}

}  // namespace thy
