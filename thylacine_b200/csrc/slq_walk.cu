// SLQ replacement step as ONE kernel: spawn from a survivor, M constrained random-walk steps, commit into the retired slot.
// Reference behaviour being replaced: SlqEngine.scala:120-197 (a new sample is accepted only above the current
// likelihood threshold); the random walk itself is this build's substitute for the cubical cover (DESIGN.md "SLQ").
//
// Scope: models whose single likelihood term is linear (Gaussian or Cauchy observations), reads the whole flat vector
// in order, and fits shared memory together with the per-coordinate priors; everything else takes the multi-kernel
// path in slq.cu (propose / prior / likelihood / accept per step), which draws the SAME random numbers.
//
// Layout: one warp owns 8 walkers for all M steps.  The walker state lives in registers as the A-operand fragments of
// DMMA.8x8x4 (lane (g, q) holds coordinates 4 kk + q of walker g), the coefficient matrix is staged in shared memory
// once per CTA with the row pitch = 4 (mod 8) the fused evaluation kernel uses, so the forward pass Z = X A^T runs on
// the fp64 tensor cores straight from the walker registers; priors, residuals and the Metropolis test are quad
// reductions.  No walker state touches HBM between the spawn and the commit.
#include "slq_internal.h"
#include "ptx.cuh"

namespace thy {

namespace {

constexpr int WNW = 8;  // warps per CTA

struct WalkParams {
  const double* A;       // [n_pad][LDS]
  const double2* yiv;    // [n_pad] (y - offset, 1/var)
  PriorDev pr;
  int n, n_pad, d;
  int distribution;
  double log_const;
  const int* c_dev;      // walkers this round = local retirees (device word written by the selection step)
  long long Pl;
  int M;
  uint64_t seed;
  const long long* round_dev;   // round number (device word: the launch replays from a CUDA graph)
  int rank;
  const int* sidx;       // [c] retired slots (ascending)
  const unsigned char* flags;   // [Pl] 1 = retired this round
  double* x;             // [Pl][d] live pool
  double* ll;            // [Pl]
  double* lpr;           // [Pl]
  const double* sigma;   // [d]
  const double* scalars; // [0] walk scale, [1] threshold
  int* accepted;         // [c] per-walker acceptance tally (+=)
  // pre-drawn randomness of this round for walker groups [0, pre_groups) (k_slq_walk_rng), or null: drawn in place
  const float* zpre;     // [group][M][KS][32] proposal normals in the walk kernel's lane order
  const double* lacc;    // [group][M][8] ln(1 - u) of the Metropolis test
  int pre_groups;
};

// The walk's random numbers do not depend on its state.  With live points sharded over several GPUs a rank has few walkers
// (7 812 of 62 500 at 8 ranks), a warp runs alone on its scheduler, and the 40 sequential steps of a walker were mostly
// Philox rounds and Box-Muller evaluated one dependent instruction at a time (profiles/r02_ncu_walk_low_occupancy.json:
// ~1600 instructions per step, `wait` 41 %).  k_slq_walk_rng draws all of a round's normals and acceptance thresholds for
// all (group, step) pairs at once -- one warp per pair, on the pool-spread stream, while the all-gather and the selection
// occupy the main stream -- in exactly the order and precision the in-place path uses (same Philox calls, same fp32
// Box-Muller), so both paths produce the same walk.
template <int NT>
__global__ void __launch_bounds__(256) k_slq_walk_rng(int groups, int M, uint64_t seed, const long long* __restrict__ round_dev,
                                                      int rank, float* __restrict__ zpre, double* __restrict__ lacc) {
  constexpr int KS = 2 * NT;
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)groups * M) return;
  const int grp = (int)(wid / M), step = (int)(wid % M);
  const int g8 = lane >> 2, q4 = lane & 3;
  const uint64_t key = ((uint64_t)rank << 40) | (uint64_t)(grp * 8 + g8);
  const uint64_t gstep = (uint64_t)*round_dev * (uint64_t)M + (uint64_t)step;
  float* zrow = zpre + ((size_t)wid * KS) * 32 + lane;
#pragma unroll
  for (int kk = 0; kk < KS; kk += 2) {
    double z0, z1;
    rng_normal_pair_fast(seed, key, slq_move_item(4 * kk + q4), RNG_SLQ_MOVE, gstep, z0, z1);
    zrow[(size_t)kk * 32] = (float)z0;         // z0, z1 are fp32 values widened to double: the round trip is exact
    zrow[(size_t)(kk + 1) * 32] = (float)z1;
  }
  if (q4 == 0) {
    const RngDraw u = rng_uniform2(seed, key, 0, RNG_SLQ_ACCEPT, gstep);
    lacc[(size_t)wid * 8 + g8] = log(1.0 - u.u0);
  }
}

template <int NT>
__global__ void __launch_bounds__(WNW * 32) k_slq_walk(const WalkParams p) {
  constexpr int KS = 2 * NT;
  constexpr int LDS = 8 * NT + 4;
  extern __shared__ __align__(16) unsigned char smem[];
  double* As = reinterpret_cast<double*>(smem);
  double2* Ys = reinterpret_cast<double2*>(As + (size_t)p.n_pad * LDS);
  double* sg = reinterpret_cast<double*>(Ys + p.n_pad);   // [8 NT] proposal scale per coordinate (0 beyond d)
  double* p0 = sg + 8 * NT;
  double* p1 = p0 + 8 * NT;
  int* kind = reinterpret_cast<int*>(p1 + 8 * NT);

  // ---- stage the model once per CTA ----
  {
    const double2* src = reinterpret_cast<const double2*>(p.A);
    double2* dst = reinterpret_cast<double2*>(As);
    const int n2 = p.n_pad * LDS / 2;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) dst[i] = src[i];
    for (int i = threadIdx.x; i < p.n_pad; i += blockDim.x) Ys[i] = p.yiv[i];
    const double scale = p.scalars[0];
    for (int j = threadIdx.x; j < 8 * NT; j += blockDim.x) {
      const bool in = j < p.d;
      sg[j] = in ? scale * p.sigma[j] : 0.0;
      p0[j] = in ? p.pr.p0[j] : 0.0;
      p1[j] = in ? p.pr.p1[j] : 0.0;
      kind[j] = in ? p.pr.kind[j] : -1;
    }
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g8 = lane >> 2, q4 = lane & 3;
  const int pi_g = (g8 < 4) ? g8 : (g8 ^ 1);                  // observation-row permutation of the fused kernel
  const int pi_s0 = (q4 < 2) ? (2 * q4) : (2 * q4 + 1);
  const int pi_s1 = (q4 < 2) ? (2 * q4 + 1) : (2 * q4);
  const int off1 = pi_g * LDS + q4;
  const double thr = p.scalars[1];
  const int ngrp = (p.n + 7) / 8;   // row groups that hold observations (rows beyond n are zero padding)
  const int d = p.d;
  const int c = *p.c_dev;
  const uint64_t round = (uint64_t)*p.round_dev;

  for (long long grp = (long long)blockIdx.x * WNW + warp; grp * 8 < c; grp += (long long)gridDim.x * WNW) {
    const long long w = grp * 8 + g8;
    const bool live = w < c;
    const uint64_t key = ((uint64_t)p.rank << 40) | (uint64_t)(live ? w : 0);
    // ---- spawn: a uniformly chosen survivor (same draw as k_slq_spawn) ----
    long long src = 0;
    int slot = 0;
    if (live) {
      if (q4 == 0) src = slq_pick_survivor(p.seed, key, round, p.Pl, p.flags);
      slot = p.sidx[w];
    }
    src = __shfl_sync(0xffffffffu, src, lane & ~3);   // one lane of the quad draws, the quad shares the pick
    double xf[KS], xp[KS];
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      const int k = 4 * kk + q4;
      xf[kk] = (live && k < d) ? p.x[(size_t)src * d + k] : 0.0;
    }
    double lcur = live ? p.ll[src] : 0.0;
    double pcur = live ? p.lpr[src] : 0.0;
    int acc = 0;

    const bool pre = p.zpre != nullptr && grp < p.pre_groups;
    for (int step = 0; step < p.M; ++step) {
      const uint64_t gstep = round * (uint64_t)p.M + (uint64_t)step;
      const float* zrow = pre ? p.zpre + (((size_t)grp * p.M + step) * KS) * 32 + lane : nullptr;
      float zf[KS];
      if (pre) {
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) zf[kk] = __ldg(zrow + (size_t)kk * 32);
      }
      const double lthr = pre ? p.lacc[((size_t)grp * p.M + step) * 8 + g8] : 0.0;
      // ---- propose + prior of the proposal ----
      double pacc = 0.0;
      bool inside = true;
#pragma unroll
      for (int kk = 0; kk < KS; kk += 2) {
        double z0, z1;
        if (pre) {
          z0 = (double)zf[kk];
          z1 = (double)zf[kk + 1];
        } else {
          rng_normal_pair_fast(p.seed, key, slq_move_item(4 * kk + q4), RNG_SLQ_MOVE, gstep, z0, z1);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k = 4 * (kk + h) + q4;
          const double v = xf[kk + h] + sg[k] * (h == 0 ? z0 : z1);
          xp[kk + h] = v;
          const int kd = kind[k];
          if (kd == PK_GAUSS_DIAG) {
            const double dl = p0[k] - v;
            pacc += dl * (dl * p1[k]);
          } else if (kd == PK_UNIFORM) {
            inside = inside && (v >= p0[k]) && (v < p1[k]);
          }
        }
      }
      pacc = quad_sum(pacc);
      const unsigned ballot = __ballot_sync(0xffffffffu, inside);
      const bool inside_all = ((ballot >> (lane & ~3)) & 0xFu) == 0xFu;
      const double lpn = inside_all ? (-0.5 * pacc + p.pr.log_const) : -INFINITY;
      // ---- forward pass on the tensor cores, residual quadratic form ----
      // (Eight independent accumulator chains per block instead of two were measured and dropped: 1276 vs 1155 us per
      // round at 62 500 walkers, 299 vs 284 us at 7 812 -- the step is bound by its ~1600 dependent scalar instructions,
      // Philox and the prior test, not by the DMMA chain; profiles/r02_ncu_walk_low_occupancy.json.)
      double ssq = 0.0;
      for (int j = 0; j < ngrp; j += 2) {
        double za0 = 0.0, za1 = 0.0, zb0 = 0.0, zb1 = 0.0;
        const double* Aj = As + (size_t)j * 8 * LDS + off1;
        const bool two = (j + 1) < ngrp;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
          ptx::dmma884(za0, za1, xp[kk], Aj[4 * kk]);
          if (two) ptx::dmma884(zb0, zb1, xp[kk], Aj[8 * LDS + 4 * kk]);
        }
        {
          const double2 y0 = Ys[j * 8 + pi_s0], y1 = Ys[j * 8 + pi_s1];
          const double r0 = y0.x - za0, r1 = y1.x - za1;
          ssq = fma(r0, r0 * y0.y, ssq);
          ssq = fma(r1, r1 * y1.y, ssq);
        }
        if (two) {
          const double2 y0 = Ys[(j + 1) * 8 + pi_s0], y1 = Ys[(j + 1) * 8 + pi_s1];
          const double r0 = y0.x - zb0, r1 = y1.x - zb1;
          ssq = fma(r0, r0 * y0.y, ssq);
          ssq = fma(r1, r1 * y1.y, ssq);
        }
      }
      const double q = quad_sum(ssq);
      const double lnew = (p.distribution == THY_DIST_GAUSSIAN) ? (-0.5 * q + p.log_const)
                                                                : (p.log_const - (1.0 + p.n) / 2.0 * log(1.0 + q));
      // ---- Metropolis test w.r.t. the prior, hard likelihood constraint (same draw as k_slq_accept) ----
      double lu = lthr;
      if (!pre) {
        const RngDraw u = rng_uniform2(p.seed, key, 0, RNG_SLQ_ACCEPT, gstep);
        lu = log(1.0 - u.u0);
      }
      const bool ok = live && isfinite(lpn) && isfinite(lnew) && (lnew > thr) && (lpn - pcur >= lu);
      if (ok) {
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) xf[kk] = xp[kk];
        lcur = lnew;
        pcur = lpn;
        ++acc;
      }
    }
    // ---- commit into the retired slot ----
    if (live) {
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        const int k = 4 * kk + q4;
        if (k < d) p.x[(size_t)slot * d + k] = xf[kk];
      }
      if (q4 == 0) {
        p.ll[slot] = lcur;
        p.lpr[slot] = pcur;
        p.accepted[w] += acc;
      }
    }
  }
}

template <int NT>
thy_status launch_walk_nt(const WalkParams& p, int c_max, size_t smem, cudaStream_t s) {
  THY_CUDA_CHECK(cudaFuncSetAttribute(k_slq_walk<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  THY_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_slq_walk<NT>, WNW * 32, smem));
  if (per_sm < 1) per_sm = 1;
  // sized for the largest possible retire count; the kernel reads the actual one from device memory (grid-stride loop)
  const long long groups = (c_max + 7) / 8;
  long long blocks = (groups + WNW - 1) / WNW;
  const long long cap = (long long)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  k_slq_walk<NT><<<(unsigned)blocks, WNW * 32, smem, s>>>(p);
  count_launch();
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

size_t walk_smem_bytes(const LikDev& lk, int nt) {
  const int LDS = 8 * nt + 4;
  return (size_t)lk.n_pad * LDS * 8 + (size_t)lk.n_pad * 16 + (size_t)8 * nt * (3 * 8 + 4) + 16;
}

}  // namespace

bool slq_walk_supported(const thy_model_s* m) {
  if (m->lik_dev.size() != 1 || m->prior_view.n_special != 0) return false;
  const LikDevOwner& own = *m->lik_dev[0];
  const LikDev& lk = own.view;
  if (lk.link != THY_LINK_IDENTITY || lk.jacobian != THY_JAC_CONSTANT) return false;
  if (lk.distribution != THY_DIST_GAUSSIAN && lk.distribution != THY_DIST_CAUCHY) return false;
  if (lk.dl != m->d || !own.contiguous || own.col_host.empty() || own.col_host[0] != 0) return false;
  const int nt = fused_nt_for(lk.dl);
  if (nt <= 0 || lk.lds != 8 * nt + 4) return false;
  return walk_smem_bytes(lk, nt) <= (size_t)200 * 1024;
}

thy_status launch_slq_walk(thy_model_s* m, int c_max, const int* c_dev, long long Pl, int M, uint64_t seed,
                           const long long* round_dev, int rank, const int* sidx, const unsigned char* flags, double* x,
                           double* ll, double* lpr, const double* sigma, const double* scalars, int* accepted, cudaStream_t s,
                           const float* zpre, const double* lacc, int pre_groups) {
  if (c_max <= 0) return THY_OK;
  const LikDev& lk = m->lik_dev[0]->view;
  WalkParams p{};
  p.A = lk.A;
  p.yiv = lk.yiv_lin;
  p.pr = m->prior_view;
  p.n = lk.n;
  p.n_pad = lk.n_pad;
  p.d = m->d;
  p.distribution = lk.distribution;
  p.log_const = lk.log_const;
  p.c_dev = c_dev;
  p.Pl = Pl;
  p.M = M;
  p.seed = seed;
  p.round_dev = round_dev;
  p.rank = rank;
  p.sidx = sidx;
  p.flags = flags;
  p.zpre = zpre;
  p.lacc = lacc;
  p.pre_groups = zpre ? pre_groups : 0;
  p.x = x;
  p.ll = ll;
  p.lpr = lpr;
  p.sigma = sigma;
  p.scalars = scalars;
  p.accepted = accepted;
  const int nt = fused_nt_for(lk.dl);
  const size_t smem = walk_smem_bytes(lk, nt);
  switch (nt) {
#define THY_WALK_CASE(N) case N: return launch_walk_nt<N>(p, c_max, smem, s);
    THY_WALK_CASE(1) THY_WALK_CASE(2) THY_WALK_CASE(3) THY_WALK_CASE(4) THY_WALK_CASE(5) THY_WALK_CASE(7)
    THY_WALK_CASE(10) THY_WALK_CASE(13) THY_WALK_CASE(16)
#undef THY_WALK_CASE
    default: return fail(THY_ERR_UNSUPPORTED, "SLQ fused walk: unsupported width");
  }
}

// doubles / floats of pre-drawn randomness per walker group (8 walkers) for M steps
size_t slq_walk_rng_floats_per_group(const thy_model_s* m, int M) {
  const int nt = fused_nt_for(m->lik_dev[0]->view.dl);
  return (size_t)M * 2 * nt * 32;
}

thy_status launch_slq_walk_rng(thy_model_s* m, int groups, int M, uint64_t seed, const long long* round_dev, int rank,
                               float* zpre, double* lacc, cudaStream_t s) {
  if (groups <= 0) return THY_OK;
  const int nt = fused_nt_for(m->lik_dev[0]->view.dl);
  const unsigned grid = (unsigned)(((long long)groups * M + 7) / 8);
  switch (nt) {
#define THY_RNG_CASE(N) case N: k_slq_walk_rng<N><<<grid, 256, 0, s>>>(groups, M, seed, round_dev, rank, zpre, lacc); break;
    THY_RNG_CASE(1) THY_RNG_CASE(2) THY_RNG_CASE(3) THY_RNG_CASE(4) THY_RNG_CASE(5) THY_RNG_CASE(7)
    THY_RNG_CASE(10) THY_RNG_CASE(13) THY_RNG_CASE(16)
#undef THY_RNG_CASE
    default: return fail(THY_ERR_UNSUPPORTED, "SLQ fused walk: unsupported width");
  }
  count_launch();
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

}  // namespace thy
