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

// ONE CTA per SM, its width chosen per launch (walk_cta_warps): a rank that replaces few points spreads its walker groups
// evenly over the SMs -- two CTAs landing on one SM put three warps on a scheduler there, and the kernel ends with its
// slowest warp (DMMA.8x8x4 issues every 26 / 33 / 50 clocks per warp with 1 / 2 / 3 warps on a scheduler,
// profiles/r02_dmma_issue_model.jsonl).  Widest CTA: 12 warps at <= 170 registers with the randomness drawn ahead,
// 16 at <= 128 when it is drawn in place; 8 for wide models (NT > 7), whose state needs up to 255 registers.
template <int NT, bool PRE> struct WalkCta { static constexpr int MAXW = (NT > 7) ? 8 : (PRE ? 12 : 16); };
constexpr int WNW_MIN = 8;   // the width slq_walk_supported() budgets shared memory for
constexpr int WCH = 8;  // 8-row blocks whose DMMA chains are interleaved; shared-memory rows are padded to 8 WCH

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
  int* accepted;         // [0] the round's acceptances over all walkers (one atomic per warp; k_slq_adapt reads and clears it)
  // pre-drawn randomness of this round for walker groups [0, pre_groups) (k_slq_walk_rng), or null: drawn in place
  const float* zpre;     // [group][M][KS][32] proposal normals in the walk kernel's lane order
  const double* lacc;    // [group][M][8] ln(1 - u) of the Metropolis test
  int pre_groups;
  long long grp_begin, grp_end;   // walker groups this launch handles (pre-drawn launch: [0, pre_groups); in place: the rest)
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
  for (int kk = 0; kk < KS; kk += 4) {
    double z[4];
    rng_normal_quad_fast(seed, key, slq_move_item(4 * kk + q4), RNG_SLQ_MOVE, gstep, z);
#pragma unroll
    for (int h = 0; h < 4; ++h)
      if (kk + h < KS) zrow[(size_t)(kk + h) * 32] = (float)z[h];   // fp32 values widened to double: the round trip is exact
  }
  if (q4 == 0) {
    const RngDraw u = rng_uniform2(seed, key, 0, RNG_SLQ_ACCEPT, gstep);
    lacc[(size_t)wid * 8 + g8] = log(1.0 - u.u0);
  }
}

// Second version (round 2, profiles/r02_walk_rewrite.md).  The first issued ~880 instructions per step of a group -- one
// LDS and a WARPSYNC/NOP pair per predicated DMMA, 14 reconvergence regions for the prior kinds, shared-memory addresses
// rebuilt from %tid inside the loop -- plus ~950 for seven Philox calls when the randomness is drawn in place.  Now: row
// blocks come in unpredicated groups of WCH (shared-memory rows are padded with zero coefficients and zero weights, so a
// padding block adds exactly 0), the prior test is branch-free, four proposal normals come from one Philox call, PRE
// (randomness drawn ahead) is a template parameter so neither instantiation carries the other's registers or code, the
// proposal lives in shared memory only, and the round's acceptances leave through one integer atomic per warp.  The
// arithmetic of a step and its order are unchanged; the proposal normals are a different (equally valid) stream.
template <int NT, bool PRE>
__global__ void __launch_bounds__(WalkCta<NT, PRE>::MAXW * 32, 1) k_slq_walk(const WalkParams p) {
  const int WNW = blockDim.x >> 5;
  constexpr int KS = 2 * NT;
  constexpr int LDS = 8 * NT + 4;
  extern __shared__ __align__(16) unsigned char smem[];
  const int rows = (p.n_pad + 8 * WCH - 1) / (8 * WCH) * (8 * WCH);   // zero rows / zero weights beyond n_pad
  double* As = reinterpret_cast<double*>(smem);
  double2* Ys = reinterpret_cast<double2*>(As + (size_t)rows * LDS);
  double4* pri = reinterpret_cast<double4*>(Ys + rows);   // [8 NT] (proposal scale, p0, p1, kind as a double)
  double* xps_all = reinterpret_cast<double*>(pri + 8 * NT);   // [WNW][KS][32] the warps' proposals as DMMA A fragments

  const int c = *p.c_dev;
  const long long grp_end = min(p.grp_end, ((long long)c + 7) / 8);
  if (p.grp_begin + (long long)blockIdx.x * WNW >= grp_end) return;   // nothing for this CTA (whole CTA: uniform)
  // ---- stage the model once per CTA ----
  {
    const double2* src = reinterpret_cast<const double2*>(p.A);
    double2* dst = reinterpret_cast<double2*>(As);
    const int n2 = p.n_pad * LDS / 2, r2 = rows * LDS / 2;
    for (int i = threadIdx.x; i < r2; i += blockDim.x) dst[i] = i < n2 ? src[i] : make_double2(0.0, 0.0);
    for (int i = threadIdx.x; i < rows; i += blockDim.x) Ys[i] = i < p.n_pad ? p.yiv[i] : make_double2(0.0, 0.0);
    const double scale = p.scalars[0];
    for (int j = threadIdx.x; j < 8 * NT; j += blockDim.x) {
      const bool in = j < p.d;
      pri[j] = make_double4(in ? scale * p.sigma[j] : 0.0, in ? p.pr.p0[j] : 0.0, in ? p.pr.p1[j] : 0.0,
                            in ? (double)p.pr.kind[j] : -1.0);
    }
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g8 = lane >> 2, q4 = lane & 3;
  const int pi_g = (g8 < 4) ? g8 : (g8 ^ 1);                  // observation-row permutation of the fused kernel
  const int pi_s0 = (q4 < 2) ? (2 * q4) : (2 * q4 + 1);
  const int pi_s1 = (q4 < 2) ? (2 * q4 + 1) : (2 * q4);
  const double* Alane = As + pi_g * LDS + q4;
  const double2* Y0 = Ys + pi_s0;
  const double2* Y1 = Ys + pi_s1;
  const double4* prl = pri + q4;
  double* xps = xps_all + (size_t)warp * KS * 32 + lane;
  const double thr = p.scalars[1];
  const int nchunk = ((p.n + 7) / 8 + WCH - 1) / WCH;   // chunks of WCH 8-row blocks that hold observations
  const int d = p.d, M = p.M;
  const uint64_t round = (uint64_t)*p.round_dev;
  const bool gaussian = p.distribution == THY_DIST_GAUSSIAN;
  const double log_const = p.log_const, prior_const = p.pr.log_const;
  int acc_warp = 0;

  for (long long grp = p.grp_begin + (long long)blockIdx.x * WNW + warp; grp < grp_end; grp += (long long)gridDim.x * WNW) {
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
    double xf[KS];   // the proposal lives in shared memory only (xps): 2 KS fewer registers per thread
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      const int k = 4 * kk + q4;
      xf[kk] = (live && k < d) ? p.x[(size_t)src * d + k] : 0.0;
    }
    double lcur = live ? p.ll[src] : 0.0;
    double pcur = live ? p.lpr[src] : 0.0;
    int acc = 0;

    const float* zrow = nullptr;
    const double* lrow = nullptr;
    float zn[KS];
    double ln = 0.0;
    if (PRE) {
      zrow = p.zpre + ((size_t)grp * M * KS) * 32 + lane;
      lrow = p.lacc + (size_t)grp * M * 8 + g8;
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) zn[kk] = __ldg(zrow + kk * 32);
      ln = __ldg(lrow);
    }
    for (int step = 0; step < M; ++step) {
      const uint64_t gstep = round * (uint64_t)M + (uint64_t)step;
      float zf[KS];
      double lu = ln;
      if (PRE) {
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) zf[kk] = zn[kk];
        if (step + 1 < M) {   // the next step's randomness travels while this step computes
          zrow += KS * 32;
          lrow += 8;
#pragma unroll
          for (int kk = 0; kk < KS; ++kk) zn[kk] = __ldg(zrow + kk * 32);
          ln = __ldg(lrow);
        }
      }
      // ---- propose + prior of the proposal (branch-free over the prior kinds) ----
      double pacc = 0.0;
      int outside = 0;   // integer logic: `&&` would compile to one reconvergence region per coordinate
#pragma unroll
      for (int kk = 0; kk < KS; kk += 4) {
        double z[4];
        if (PRE) {
#pragma unroll
          for (int h = 0; h < 4; ++h) z[h] = (kk + h < KS) ? (double)zf[(kk + h < KS) ? kk + h : 0] : 0.0;
        } else {
          rng_normal_quad_fast(p.seed, key, slq_move_item(4 * kk + q4), RNG_SLQ_MOVE, gstep, z);
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          if (kk + h < KS) {
            const double4 pk = prl[4 * (kk + h)];
            const double v = xf[kk + h] + pk.x * z[h];
            xps[(kk + h) * 32] = v;
            const double dl = pk.y - v;
            const double gq = dl * (dl * pk.z);
            pacc += (pk.w == (double)PK_GAUSS_DIAG) ? gq : 0.0;
            outside |= (int)(pk.w == (double)PK_UNIFORM) & ((int)!(v >= pk.y) | (int)!(v < pk.z));
          }
        }
      }
      pacc = quad_sum(pacc);
      const unsigned ballot = __ballot_sync(0xffffffffu, outside == 0);
      const bool inside_all = ((ballot >> (lane & ~3)) & 0xFu) == 0xFu;
      const double lpn = inside_all ? (-0.5 * pacc + prior_const) : -INFINITY;
      // ---- forward pass on the tensor cores, residual quadratic form: WCH independent 8-row blocks at a time ----
      // The K loop is kept ROLLED: one iteration = one K-slice of all WCH row blocks, i.e. WCH independent DMMAs (one warp
      // issues a DMMA.8x8x4 every 26 clocks down a dependent chain and every 18.5 across independent ones,
      // profiles/r02_dmma_issue_model.jsonl; ptxas regroups an unrolled sequence into one serial chain per accumulator
      // whatever order the PTX has).  It is also a third of the code of the unrolled form.  The proposal comes back from
      // shared memory (a rolled loop cannot index registers); each lane reads only what it wrote itself.
      double ssq = 0.0;
      const double* Aj = Alane;
      const double2 *y0p = Y0, *y1p = Y1;
      for (int jc = 0; jc < nchunk; ++jc, Aj += WCH * 8 * LDS, y0p += WCH * 8, y1p += WCH * 8) {
        double z0[WCH], z1[WCH];
#pragma unroll
        for (int cb = 0; cb < WCH; ++cb) z0[cb] = z1[cb] = 0.0;
        const double* bp = Aj;
        const double* xq = xps;
#pragma unroll 1
        for (int kk = 0; kk < KS; ++kk, bp += 4, xq += 32) {
          const double a = *xq;
          double b[WCH];
#pragma unroll
          for (int cb = 0; cb < WCH; ++cb) b[cb] = bp[cb * 8 * LDS];
#pragma unroll
          for (int cb = 0; cb < WCH; ++cb) ptx::dmma884(z0[cb], z1[cb], a, b[cb]);
        }
#pragma unroll
        for (int cb = 0; cb < WCH; ++cb) {
          const double2 y0 = y0p[cb * 8], y1 = y1p[cb * 8];
          const double r0 = y0.x - z0[cb], r1 = y1.x - z1[cb];
          ssq = fma(r0, r0 * y0.y, ssq);
          ssq = fma(r1, r1 * y1.y, ssq);
        }
      }
      const double q = quad_sum(ssq);
      const double lnew = gaussian ? (-0.5 * q + log_const) : (log_const - (1.0 + p.n) / 2.0 * log(1.0 + q));
      // ---- Metropolis test w.r.t. the prior, hard likelihood constraint (same draw as k_slq_accept) ----
      if (!PRE) {
        const RngDraw u = rng_uniform2(p.seed, key, 0, RNG_SLQ_ACCEPT, gstep);
        lu = log(1.0 - u.u0);
      }
      const bool ok = live && isfinite(lpn) && isfinite(lnew) && (lnew > thr) && (lpn - pcur >= lu);
      if (ok) {
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) xf[kk] = xps[kk * 32];
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
      }
    }
    acc_warp += __reduce_add_sync(0xffffffffu, (live && q4 == 0) ? acc : 0);
  }
  if (lane == 0 && acc_warp != 0) atomicAdd(p.accepted, acc_warp);
}

size_t walk_smem_bytes(const LikDev& lk, int nt, int warps) {
  const int LDS = 8 * nt + 4;
  const size_t rows = (size_t)round_up(lk.n_pad, 8 * WCH);
  return rows * LDS * 8 + rows * 16 + (size_t)8 * nt * 32 + (size_t)warps * 2 * nt * 32 * 8 + 16;
}
constexpr size_t WALK_SMEM_BUDGET = (size_t)200 * 1024;

// CTA shape for `groups` walker groups expected on this rank.  One pass over the SMs suffices: ONE CTA per SM, as narrow
// as holds them, so the groups spread evenly.  Otherwise: two CTAs of maxw / 2 warps per SM and a grid-stride loop (measured
// at 7 813 groups: 1 044-1 089 us per round against 1 213 for one 14-warp CTA per SM that wastes less of its last pass).
struct WalkShape { int warps, ctas_per_sm; };
WalkShape walk_cta_shape(long long groups, int sms, int maxw) {
  if (groups <= (long long)sms * maxw) return {(int)std::max<long long>(1, (groups + sms - 1) / sms), 1};
  return {maxw / 2, 2};
}

template <int NT, bool PRE>
thy_status launch_walk_pre(WalkParams p, const LikDev& lk, long long grp_begin, long long grp_end, long long expected_groups,
                           cudaStream_t s) {
  if (grp_end <= grp_begin) return THY_OK;
  p.grp_begin = grp_begin;
  p.grp_end = grp_end;
  const int sms = sm_count();
  const WalkShape shape = walk_cta_shape(std::min(expected_groups, grp_end - grp_begin), sms, WalkCta<NT, PRE>::MAXW);
  int warps = shape.warps;
  while (warps > 1 && walk_smem_bytes(lk, NT, warps) * shape.ctas_per_sm > WALK_SMEM_BUDGET) --warps;
  size_t smem = walk_smem_bytes(lk, NT, warps);
  // One CTA per SM is only an even spread if the block scheduler cannot put two on one SM: kernels of the quadrature
  // stream occupy some SMs for a few microseconds at the moment this grid is placed, and the CTAs they displace would
  // double up elsewhere.  Asking for more than half of an SM's shared memory rules that out.
  if (shape.ctas_per_sm == 1) smem = std::max(smem, (size_t)116 * 1024);
  THY_CUDA_CHECK(cudaFuncSetAttribute(k_slq_walk<NT, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // sized for the largest possible retire count; the kernel reads the actual one from device memory (grid-stride loop)
  long long blocks = (grp_end - grp_begin + warps - 1) / warps;
  if (blocks > (long long)sms * shape.ctas_per_sm) blocks = (long long)sms * shape.ctas_per_sm;
  k_slq_walk<NT, PRE><<<(unsigned)blocks, warps * 32, smem, s>>>(p);
  count_launch();
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

// Groups [0, pre_groups) read the randomness drawn ahead; the rest (none, unless a rank replaces more than 1.25x its share
// of a round) draw in place.  Both launches read the actual walker count from device memory; a CTA with nothing to do
// returns before it stages the model.
template <int NT>
thy_status launch_walk_nt(const WalkParams& p, const LikDev& lk, int c_max, int expected, cudaStream_t s) {
  const long long groups = (c_max + 7) / 8, exp_groups = (std::min(expected, c_max) + 7) / 8;
  const long long pre = p.zpre ? std::min<long long>(p.pre_groups, groups) : 0;
  THY_TRY((launch_walk_pre<NT, true>(p, lk, 0, pre, exp_groups, s)));
  return launch_walk_pre<NT, false>(p, lk, pre, groups, pre > 0 ? 1 : exp_groups, s);
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
  return walk_smem_bytes(lk, nt, WNW_MIN) <= WALK_SMEM_BUDGET;
}

thy_status launch_slq_walk(thy_model_s* m, int c_max, const int* c_dev, long long Pl, int M, uint64_t seed,
                           const long long* round_dev, int rank, const int* sidx, const unsigned char* flags, double* x,
                           double* ll, double* lpr, const double* sigma, const double* scalars, int* accepted, cudaStream_t s,
                           const float* zpre, const double* lacc, int pre_groups, int expected) {
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
  switch (nt) {
#define THY_WALK_CASE(N) case N: return launch_walk_nt<N>(p, lk, c_max, expected > 0 ? expected : c_max, s);
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
