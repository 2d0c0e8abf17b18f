// Counter-based RNG: Philox4x32-10 (Salmon et al., SC'11), keyed (seed) and indexed (chain, item, purpose, step).
// The reference draws from unseeded global generators (scala.util.Random, Math.random(), Breeze RandBasis seeded from
// hashCode: util/MathOps.scala:24,113-114, HmcmcEngine.scala:128), so sample-level parity is statistical only; a
// counter-based generator makes every device draw reproducible and replayable on the host (tests do exactly that).
#pragma once
#include <cstdint>

namespace thy {

struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;

  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    const uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
  }

  // ctr = (c0,c1,c2,c3), key = (k0,k1) -> 4 x 32 random bits
  __host__ __device__ static inline void gen(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                             uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(M0, c0, hi0, lo0);
      mulhilo(M1, c2, hi1, lo1);
      const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

// 53-bit uniform in [0, 1)  (Math.random() semantics)
__host__ __device__ inline double u01(uint32_t hi, uint32_t lo) {
  const uint64_t bits = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)bits * (1.0 / 9007199254740992.0);
}

// purposes (c3) -- keep distinct streams for every use
enum RngPurpose : uint32_t {
  RNG_PRIOR = 1, RNG_PRIOR_CHI = 2, RNG_HMC_MOMENTUM = 3, RNG_HMC_ACCEPT = 4, RNG_LFM_PICK = 5, RNG_LFM_PARTNER = 6,
  RNG_LFM_ACCEPT = 7, RNG_SLQ_MOVE = 8, RNG_SLQ_ACCEPT = 9, RNG_SLQ_SHRINK = 10, RNG_SLQ_PICK = 11
};

struct RngDraw {
  double u0, u1;  // two uniforms in [0,1)
};

// item: (index within the stream), step: time-like counter; 64-bit chain id split across c0/c1 low bits.
__host__ __device__ inline RngDraw rng_uniform2(uint64_t seed, uint64_t chain, uint32_t item, uint32_t purpose, uint64_t step) {
  uint32_t r[4];
  // counter words: chain (low 32), item, step (low 32), purpose | step-high/chain-high folded into the key
  const uint32_t k0 = (uint32_t)seed ^ (uint32_t)(chain >> 32) * 0x85EBCA6Bu;
  const uint32_t k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32) * 0xC2B2AE35u;
  Philox::gen((uint32_t)chain, item, (uint32_t)step, purpose, k0, k1, r);
  RngDraw d;
  d.u0 = u01(r[0], r[1]);
  d.u1 = u01(r[2], r[3]);
  return d;
}

// one standard normal (Box-Muller, cosine branch); u0 -> (0,1] to keep the log finite
__host__ __device__ inline double rng_normal(uint64_t seed, uint64_t chain, uint32_t item, uint32_t purpose, uint64_t step) {
  const RngDraw d = rng_uniform2(seed, chain, item, purpose, step);
  const double r = sqrt(-2.0 * log(1.0 - d.u0));
  return r * cospi(2.0 * d.u1);
}

// SLQ random-walk proposals: FOUR standard normals per Philox call (both Box-Muller branches of two (radius, angle)
// pairs; every word of the 4x32 output is used), evaluated with fp32 special functions (MUFU log / sincospi).  A proposal
// only has to be symmetric and cheap -- its 24-bit resolution and the +-5.8 sigma truncation do not enter the acceptance
// test, which stays in fp64.  ~10x cheaper per normal than rng_normal.
#ifdef __CUDACC__
__device__ __forceinline__ void rng_normal_quad_fast(uint64_t seed, uint64_t chain, uint32_t item, uint32_t purpose,
                                                     uint64_t step, double (&z)[4]) {
  uint32_t r[4];
  const uint32_t k0 = (uint32_t)seed ^ (uint32_t)(chain >> 32) * 0x85EBCA6Bu;
  const uint32_t k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32) * 0xC2B2AE35u;
  Philox::gen((uint32_t)chain, item, (uint32_t)step, purpose, k0, k1, r);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u = (float)((r[2 * h] >> 8) + 1u) * (1.0f / 16777216.0f);   // (0, 1]
    const float v = (float)(r[2 * h + 1] >> 8) * (1.0f / 8388608.0f);       // [0, 2): angle / pi
    const float rad = sqrtf(-2.0f * __logf(u));
    float sn, cs;
    sincospif(v, &sn, &cs);
    z[2 * h] = (double)(rad * cs);
    z[2 * h + 1] = (double)(rad * sn);
  }
}
#endif
// coordinate j of a walker -> (item, branch) of its proposal quadruple: lanes of the fused walk kernel own coordinates
// 4 kk + q, so grouping kk, kk + 1, kk + 2, kk + 3 keeps all four normals of a call in one thread
__host__ __device__ inline uint32_t slq_move_item(int j) { return (uint32_t)(((j >> 4) << 2) | (j & 3)); }
__host__ __device__ inline int slq_move_branch(int j) { return (j >> 2) & 3; }

}  // namespace thy
