// SLQ threshold step on the device: exact selection of the K lowest live points WITHOUT sorting the pool, as ONE
// persistent kernel.
// Reference step being replaced: SlqEngine.scala:139-143 (`keys.min` of the live pool, one point per iteration) inside
// the sampling loop :120-169; here the K = sampleParallelism lowest are retired per round (DESIGN.md "SLQ").
//
// Round 1 sorted every rank's whole local pool with cub::DeviceRadixSort (8 passes over keys AND indices, ~11 kernels,
// 116-150 us at 125 000 keys) although only the K-th smallest VALUE and the SET of points below it are needed on the
// critical path.  k_slq_select is a hand-written MSD radix select, one CTA per SM, grid barriers between its phases
// (cooperative launch: all CTAs are co-resident):
//   passes 0-5   11, 11, 11, 11, 11, 9 bits of the order-preserving 64-bit image of the fp64 log-likelihood.  Every CTA
//                histograms the keys that still match the decided prefix in shared memory (whole warps hit one bin in
//                the early passes, so equal digits are aggregated with match.any first), merges into a global
//                histogram, and after the barrier every CTA scans it and extends the prefix by one digit.  After the
//                last pass the prefix IS the K-th smallest key, bit for bit, and the remainder says how many of the keys
//                equal to it still have to go.  EARLY EXIT: as soon as the decided bin holds <= 64 keys they are
//                collected and ranked directly (typically after the third pass: three passes and two barriers saved).
//   count        per 2048-key slice of the GLOBAL key array: keys below / equal to the threshold; the retired VALUES
//                (all keys below the threshold + the ties that go) are appended, unordered, to the quadrature's input
//                (they are sorted on the quadrature stream, off the critical path).
//   flag         local slices only: retire flags and an index-ordered (deterministic) compaction of the retired slots.
//                Ties at the threshold are taken in global index order (rank-major), so exactly K points retire
//                world-wide even when walkers that never moved left exact duplicates in the pool (the reference nudges
//                duplicates by one ulp instead, SlqEngine.scala:107-118).
// Across GPUs the keys of all ranks are all-gathered first (8 MB at 10^6 live points); every rank then runs the same
// selection on the same array and derives its own retirees -- no second collective, identical thresholds everywhere.
// The keys are read from L2 after the first pass (8 MB << 126 MB); all comparisons are made on the integer images, so
// the count / flag phases agree bit for bit with the radix passes (-0.0 vs +0.0, NaNs).
#include "slq_internal.h"

namespace thy {

namespace {

constexpr int ST = 512;            // threads per CTA
constexpr int SW = ST / 32;        // warps
constexpr int SEL_BINS = 2048;
constexpr int NPASS = 6;
constexpr int CAND_MAX = 64;       // early exit: this few keys left in the decided bin are ranked directly
__constant__ int c_shift[NPASS] = {53, 42, 31, 20, 9, 0};
__constant__ int c_bits[NPASS] = {11, 11, 11, 11, 11, 9};

// order-preserving image of an fp64 value in the unsigned integers (NaNs sort by bit pattern, like a radix sort)
__device__ __forceinline__ unsigned long long key_image(double v) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// Grid-wide barrier on a monotonically increasing counter (never reset: every launch of an engine uses the same grid).
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int ticket = atomicAdd(bar, 1u);
    const unsigned int target = (ticket / nblocks + 1u) * nblocks;
    while ((int)(*(volatile unsigned int*)bar - target) < 0) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
}

// exclusive block scan of one int per thread (ST threads); sh: [SW] scratch
__device__ __forceinline__ int block_excl_int(int x, int* sh, int* total) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  int incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  __syncthreads();
  if (lane == 31) sh[wrp] = incl;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SW; ++w) {
    const int v = sh[w];
    base += (w < wrp) ? v : 0;
    tot += v;
  }
  if (total) *total = tot;
  return base + incl - x;
}

struct SelectArgs {
  const double* keys;   // [world * Pl] rank-major
  long long Pl;
  int world, rank, K, spr;
  SelectBuffers b;
  double* vals;         // [K] retired values, unordered
};

__global__ void __launch_bounds__(ST, 1) k_slq_select(const SelectArgs a) {
  __shared__ unsigned int sh[SEL_BINS];
  __shared__ int ssc[SW];
  __shared__ unsigned long long s_prefix;
  __shared__ unsigned int s_remaining, s_bin;
  __shared__ long long s_tie_base;
  __shared__ int s_ret_base;
  __shared__ unsigned long long s_cand[CAND_MAX];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const unsigned int G = gridDim.x;
  const long long n = a.Pl * a.world;
  const long long n_up = (n + 31) / 32 * 32;
  unsigned int* bar = a.b.ticket;

  // ================= MSD radix passes =================
  // Up to RK keys per thread stay in registers for all six passes (10^6 keys on 148 x 512 threads: 14 per thread), so a
  // pass costs its histogram updates and one grid barrier, not another trip to L2; larger pools stream from memory with
  // four independent loads in flight per thread.
  constexpr int RK = 16;
  const bool in_regs = n <= (long long)G * ST * RK;
  unsigned long long kr[RK];
  if (in_regs) {
#pragma unroll
    for (int q = 0; q < RK; ++q) {
      const long long i = ((long long)q * G + blockIdx.x) * ST + threadIdx.x;
      kr[q] = (i < n) ? key_image(a.keys[i]) : 0ull;
    }
  }
  unsigned long long prefix = 0ull;
  unsigned int remaining = (unsigned int)a.K, bin_count = 0u;
  for (int pass = 0; pass < NPASS; ++pass) {
    const int shift = c_shift[pass], bits = c_bits[pass], nb = 1 << bits, hi = shift + bits;
    unsigned int* hist = a.b.hist + pass * SEL_BINS;
    for (int i = threadIdx.x; i < SEL_BINS; i += ST) sh[i] = 0u;
    __syncthreads();
    auto tally = [&](bool valid, unsigned long long k) {   // whole warp calls this together
      unsigned int digit = 0xffffffffu;
      if (valid && ((hi >= 64) || ((k >> hi) == (prefix >> hi)))) digit = (unsigned int)(k >> shift) & (nb - 1);
      const unsigned int grp = __match_any_sync(0xffffffffu, digit);
      if (digit != 0xffffffffu && lane == (__ffs(grp) - 1)) atomicAdd(&sh[digit], (unsigned int)__popc(grp));
    };
    if (in_regs) {
#pragma unroll
      for (int q = 0; q < RK; ++q) {
        const long long i = ((long long)q * G + blockIdx.x) * ST + threadIdx.x;
        if (((long long)q * G + blockIdx.x) * ST < n) tally(i < n, kr[q]);   // warp-uniform guard
      }
    } else {
      const long long stride = (long long)G * ST;
      for (long long i0 = (long long)blockIdx.x * ST + threadIdx.x; i0 - threadIdx.x < n_up; i0 += 4 * stride) {
        unsigned long long k4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long i = i0 + u * stride;
          k4[u] = (i < n) ? key_image(a.keys[i]) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long i = i0 + u * stride;
          if (i - threadIdx.x < n_up) tally(i < n, k4[u]);
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += ST)
      if (sh[i]) atomicAdd(&hist[i], sh[i]);
    grid_barrier(bar, G);
    // every CTA finds the digit whose cumulative count reaches `remaining` (same data, same answer)
    constexpr int PER = SEL_BINS / ST;   // 4 consecutive bins per thread
    unsigned int loc[PER];
    int tot = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int bq = threadIdx.x * PER + q;
      loc[q] = (bq < nb) ? __ldcg(&hist[bq]) : 0u;
      tot += (int)loc[q];
    }
    unsigned int before = (unsigned int)block_excl_int(tot, ssc, nullptr);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int bq = threadIdx.x * PER + q;
      if (bq < nb && before < remaining && remaining <= before + loc[q]) {   // exactly one (thread, q)
        s_prefix = prefix | ((unsigned long long)bq << shift);
        s_remaining = remaining - before;
        s_bin = loc[q];
      }
      before += loc[q];
    }
    __syncthreads();
    prefix = s_prefix;
    remaining = s_remaining;
    bin_count = s_bin;
    __syncthreads();
    // ---- early exit ----
    // Log-likelihoods of a pool share their sign and (almost) their exponent, so after the third pass the decided bin
    // typically holds one or two keys and the remaining passes would histogram a handful of keys with a grid barrier
    // each.  Once at most CAND_MAX keys are left they are collected (one more look at the keys in registers, one grid
    // barrier) and every CTA ranks them directly: same threshold, same tie counts, bit for bit.  The candidate list
    // lives in the last pass's histogram, which an early exit leaves unused and the re-arm step zeroes.
    if (pass < NPASS - 1 && bin_count <= (unsigned int)CAND_MAX) {
      unsigned int* cc = a.b.hist + (NPASS - 1) * SEL_BINS;
      unsigned long long* cand = reinterpret_cast<unsigned long long*>(cc + 2);
      auto collect = [&](bool valid, unsigned long long k) {
        if (valid && (k >> shift) == (prefix >> shift)) {
          const unsigned int idx = atomicAdd(cc, 1u);
          if (idx < (unsigned int)CAND_MAX) cand[idx] = k;
        }
      };
      if (in_regs) {
#pragma unroll
        for (int q = 0; q < RK; ++q) {
          const long long i = ((long long)q * G + blockIdx.x) * ST + threadIdx.x;
          collect(i < n, kr[q]);
        }
      } else {
        const long long stride = (long long)G * ST;
        for (long long i = (long long)blockIdx.x * ST + threadIdx.x; i < n; i += stride) collect(true, key_image(a.keys[i]));
      }
      grid_barrier(bar, G);
      const unsigned int nc = bin_count;
      if (threadIdx.x < nc) s_cand[threadIdx.x] = __ldcg(&cand[threadIdx.x]);
      __syncthreads();
      if (threadIdx.x < nc) {
        const unsigned long long me = s_cand[threadIdx.x];
        unsigned int less = 0, eq = 0, eq_before = 0;
        for (unsigned int j = 0; j < nc; ++j) {
          const unsigned long long o = s_cand[j];
          less += (o < me) ? 1u : 0u;
          eq += (o == me) ? 1u : 0u;
          eq_before += (o == me && j < threadIdx.x) ? 1u : 0u;
        }
        if (less < remaining && remaining <= less + eq && eq_before == 0) {   // exactly one thread
          s_prefix = me;
          s_remaining = remaining - less;
          s_bin = eq;
        }
      }
      __syncthreads();
      prefix = s_prefix;
      remaining = s_remaining;
      bin_count = s_bin;
      __syncthreads();
      break;
    }
  }
  // prefix = image of the K-th smallest key; `remaining` of the `bin_count` keys equal to it retire
  const unsigned long long thr = prefix;
  const long long need = remaining;
  const bool all_ties = remaining == bin_count;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.b.scalars[1] = key_value(thr);
    a.b.state->prefix = thr;
    a.b.state->remaining = remaining;
    a.b.state->bin_count = bin_count;
    a.b.state->ties_needed = remaining;
    a.b.state->ties_total = bin_count;
  }

  // ================= per-slice counts + retired values =================
  // A CTA takes its slices four at a time (thread t holds keys [4t, 4t+4) of each): all loads of a batch are in flight
  // together, the four (below, equal) pairs are reduced side by side, and the batch reserves its stretch of the
  // retired-value buffer with ONE atomic (order inside the buffer is irrelevant: it is sorted afterwards).
  const int nslices = a.spr * a.world;
  constexpr int SB = 4, KPT = SEL_SLICE / ST;   // slices per batch; keys per thread per slice
  int* shi = reinterpret_cast<int*>(sh);        // [SB][2][SW] warp partials, then [SB] slice bases
  for (int s0 = blockIdx.x * SB; s0 < nslices; s0 += (int)G * SB) {
    double kv[SB][KPT];
    bool tk[SB][KPT];
    int below[SB], equal[SB];
#pragma unroll
    for (int u = 0; u < SB; ++u) {
      const int s = s0 + u;
      const int r = s / a.spr, i = s % a.spr;
      const long long k0 = (long long)r * a.Pl + (long long)i * SEL_SLICE + (long long)threadIdx.x * KPT;
      const long long k1 = (long long)r * a.Pl + a.Pl;
#pragma unroll
      for (int q = 0; q < KPT; ++q) {
        const bool in = (s < nslices) && (k0 + q < k1) && ((long long)threadIdx.x * KPT + q < SEL_SLICE);
        kv[u][q] = in ? a.keys[k0 + q] : 0.0;
        tk[u][q] = in;
      }
    }
#pragma unroll
    for (int u = 0; u < SB; ++u) {
      below[u] = equal[u] = 0;
#pragma unroll
      for (int q = 0; q < KPT; ++q) {
        const unsigned long long v = key_image(kv[u][q]);
        equal[u] += (tk[u][q] && v == thr) ? 1 : 0;
        tk[u][q] = tk[u][q] && v < thr;
        below[u] += tk[u][q] ? 1 : 0;
      }
    }
    // in-thread -> warp -> CTA exclusive offsets of the values to append (all SB slices of the batch as one run)
    int mine = 0;
#pragma unroll
    for (int u = 0; u < SB; ++u) mine += below[u];
    int run_total = 0;
    const int my_off = block_excl_int(mine, ssc, &run_total);
#pragma unroll
    for (int u = 0; u < SB; ++u) {
      int b = below[u], e = equal[u];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        b += __shfl_xor_sync(0xffffffffu, b, o);
        e += __shfl_xor_sync(0xffffffffu, e, o);
      }
      if (lane == 0) {
        shi[(u * 2 + 0) * SW + wrp] = b;
        shi[(u * 2 + 1) * SW + wrp] = e;
      }
    }
    __syncthreads();
    if (threadIdx.x < SB) {
      const int u = threadIdx.x;
      int bsum = 0, esum = 0;
      for (int w = 0; w < SW; ++w) {
        bsum += shi[(u * 2 + 0) * SW + w];
        esum += shi[(u * 2 + 1) * SW + w];
      }
      if (s0 + u < nslices) a.b.cnt[s0 + u] = make_int2(bsum, esum);
    }
    if (threadIdx.x == 0) s_bin = run_total > 0 ? atomicAdd(a.b.cursor, (unsigned int)run_total) : 0u;
    __syncthreads();
    unsigned int p = s_bin + (unsigned int)my_off;
#pragma unroll
    for (int u = 0; u < SB; ++u)
#pragma unroll
      for (int q = 0; q < KPT; ++q)
        if (tk[u][q]) {
          if (p < (unsigned int)a.K) a.vals[p] = kv[u][q];
          ++p;
        }
    __syncthreads();
  }
  if (blockIdx.x == 0) {   // the ties that go, as values
    __syncthreads();
    if (threadIdx.x == 0) sh[0] = atomicAdd(a.b.cursor, (unsigned int)need);
    __syncthreads();
    const unsigned int base = sh[0];
    const double tv = key_value(thr);
    for (unsigned int t = threadIdx.x; t < (unsigned int)need; t += ST)
      if (base + t < (unsigned int)a.K) a.vals[base + t] = tv;
  }
  grid_barrier(bar, G);

  // ================= retire flags + ordered compaction (local slices) =================
  for (int i = blockIdx.x; i < a.spr; i += (int)G) {
    const int s = a.rank * a.spr + i;
    __syncthreads();
    if (wrp == 0) {
      long long tb = 0;
      int rb = 0;
      if (all_ties) {
        for (int q = lane; q < i; q += 32) {
          const int2 c = __ldcg(&a.b.cnt[a.rank * a.spr + q]);
          rb += c.x + c.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rb += __shfl_xor_sync(0xffffffffu, rb, o);
      } else {
        for (int q = lane; q < s; q += 32) tb += __ldcg(&a.b.cnt[q]).y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tb += __shfl_xor_sync(0xffffffffu, tb, o);
        if (lane == 0) {   // ties taken in earlier local slices depend on each slice's own tie base: serial
          long long t = 0;
          for (int q = 0; q < a.rank * a.spr; ++q) t += __ldcg(&a.b.cnt[q]).y;
          for (int q = 0; q < i; ++q) {
            const int2 c = __ldcg(&a.b.cnt[a.rank * a.spr + q]);
            long long take = need - t;
            take = take < 0 ? 0 : (take > c.y ? c.y : take);
            rb += c.x + (int)take;
            t += c.y;
          }
        }
      }
      if (lane == 0) {
        s_tie_base = tb;
        s_ret_base = rb;
      }
    }
    __syncthreads();
    constexpr int PER = SEL_SLICE / ST;   // 4 consecutive keys per thread
    const long long l0 = (long long)i * SEL_SLICE + (long long)threadIdx.x * PER;
    unsigned long long v[PER];
    int nt = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const long long l = l0 + q;
      v[q] = (l < a.Pl) ? key_image(a.keys[(long long)a.rank * a.Pl + l]) : ~0ull;
      nt += (l < a.Pl && v[q] == thr) ? 1 : 0;
    }
    long long tie_rank = s_tie_base + (all_ties ? 0 : block_excl_int(nt, ssc, nullptr));
    bool ret[PER];
    int nr = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const long long l = l0 + q;
      bool r = false;
      if (l < a.Pl) {
        if (v[q] < thr) r = true;
        else if (v[q] == thr) {
          r = all_ties || (tie_rank < need);
          ++tie_rank;
        }
      }
      ret[q] = r;
      nr += r ? 1 : 0;
    }
    int tot = 0;
    int pos = s_ret_base + block_excl_int(nr, ssc, &tot);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const long long l = l0 + q;
      if (l < a.Pl) {
        a.b.flags[l] = ret[q] ? 1 : 0;
        if (ret[q]) a.b.sidx[pos++] = (int)l;
      }
    }
    if (i == a.spr - 1 && threadIdx.x == 0) a.b.count[0] = s_ret_base + tot;   // retired local points
  }
  // ================= re-arm for the next round (all histogram reads are behind the barriers above) =================
  for (int i = blockIdx.x * ST + threadIdx.x; i < NPASS * SEL_BINS; i += (int)G * ST) a.b.hist[i] = 0u;
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.b.cursor = 0u;
}

}  // namespace

int slq_select_slices(long long Pl) { return (int)((Pl + SEL_SLICE - 1) / SEL_SLICE); }

// keys: the GLOBAL key array (world * Pl entries, rank-major).  Selects the K lowest: threshold -> scalars[1], the retired
// values (unordered) -> vals[0, K), this rank's retire flags / compacted slots / count.  One cooperative launch, no host
// interaction, graph-capturable.
thy_status slq_select_launch(const SelectBuffers& b, const double* keys, long long Pl, int world, int rank, int K,
                             double* vals, cudaStream_t s) {
  SelectArgs a{};
  a.keys = keys;
  a.Pl = Pl;
  a.world = world;
  a.rank = rank;
  a.K = K;
  a.spr = slq_select_slices(Pl);
  a.b = b;
  a.vals = vals;
  void* args[] = {(void*)&a};
  THY_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)k_slq_select, dim3((unsigned)sm_count()), dim3(ST), args, 0, s));
  count_launch();
  return THY_OK;
}

}  // namespace thy
