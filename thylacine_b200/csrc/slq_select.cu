// SLQ threshold step on the device: exact selection of the K lowest live points WITHOUT sorting the pool.
// Reference step being replaced: SlqEngine.scala:139-143 (`keys.min` of the live pool, one point per iteration) inside
// the sampling loop :120-169; here the K = sampleParallelism lowest are retired per round (DESIGN.md "SLQ").
//
// Round 1 sorted every rank's whole local pool with cub::DeviceRadixSort (8 passes over keys AND indices, ~11 kernels,
// 116-150 us at 125 000 keys) although only the K-th smallest VALUE and the SET of points below it are needed on the
// critical path.  This file replaces that with a hand-written MSD radix select:
//   k_sel_pass   6 passes over the keys (11, 11, 11, 11, 11, 9 bits of the order-preserving 64-bit image of the fp64
//                log-likelihood).  Every CTA histograms the keys that still match the decided prefix in shared memory,
//                merges into a global histogram, and the LAST CTA to finish (atomic ticket) scans it, extends the
//                prefix by one digit and re-arms the histogram.  After the last pass the prefix IS the K-th smallest
//                key, bit for bit, and the remainder says how many of the keys equal to it still have to go.
//   k_sel_count  per 2048-key slice: how many keys are below the threshold / equal to it.
//   k_sel_flag   local slices only: retire flags and an index-ordered (deterministic) compaction of the retired slots.
//                Ties at the threshold are taken in global index order (rank-major), so exactly K points retire
//                world-wide even when walkers that never moved left exact duplicates in the pool (the reference nudges
//                duplicates by one ulp instead, SlqEngine.scala:107-118).
// Across GPUs the keys of all ranks are all-gathered first (8 MB at 10^6 live points); every rank then runs the same
// selection on the same array and derives its own retirees -- no second collective, identical thresholds everywhere.
// The keys are read from L2 after the first pass (8 MB << 126 MB), so a pass costs a few microseconds.
#include "slq_internal.h"

namespace thy {

namespace {

constexpr int SEL_THREADS = 256;
constexpr int SEL_BINS = 2048;
constexpr int NPASS = 6;
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

// One MSD pass.  st->prefix holds the digits decided so far (bits above shift + bits), st->remaining the rank still to
// be resolved inside that prefix (1-based: the `remaining`-th smallest of the matching keys).
__global__ void __launch_bounds__(SEL_THREADS) k_sel_pass(const double* __restrict__ keys, long long n, int pass, int K,
                                                          SelectState* __restrict__ st, unsigned int* __restrict__ hist,
                                                          unsigned int* __restrict__ ticket, double* __restrict__ scalars) {
  __shared__ unsigned int sh[SEL_BINS];
  __shared__ unsigned int wsum[SEL_THREADS / 32];
  __shared__ int s_last;
  const int shift = c_shift[pass], bits = c_bits[pass], nb = 1 << bits;
  for (int i = threadIdx.x; i < SEL_BINS; i += SEL_THREADS) sh[i] = 0u;
  __syncthreads();
  const unsigned long long prefix = (pass == 0) ? 0ull : st->prefix;
  const int hi = shift + bits;   // bits [hi, 64) are decided
  // log-likelihoods of a live pool share their leading bits, so whole warps hit one bin in the early passes:
  // aggregate equal digits inside the warp (match.any) and let one lane add the group's count
  const long long n_up = (n + 31) / 32 * 32;
  for (long long i = (long long)blockIdx.x * SEL_THREADS + threadIdx.x; i < n_up; i += (long long)gridDim.x * SEL_THREADS) {
    unsigned int digit = 0xffffffffu;
    if (i < n) {
      const unsigned long long k = key_image(keys[i]);
      if ((hi >= 64) || ((k >> hi) == (prefix >> hi))) digit = (unsigned int)(k >> shift) & (nb - 1);
    }
    const unsigned int grp = __match_any_sync(0xffffffffu, digit);
    if (digit != 0xffffffffu && (threadIdx.x & 31) == (__ffs(grp) - 1)) atomicAdd(&sh[digit], (unsigned int)__popc(grp));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += SEL_THREADS)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- last CTA: find the digit whose cumulative count reaches `remaining` ----
  const unsigned int remaining = (pass == 0) ? (unsigned int)K : st->remaining;
  constexpr int PER = SEL_BINS / SEL_THREADS;   // 8 consecutive bins per thread
  unsigned int loc[PER], tot = 0;
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int b = threadIdx.x * PER + q;
    loc[q] = (b < nb) ? __ldcg(&hist[b]) : 0u;
    tot += loc[q];
  }
  // exclusive block scan of tot
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  unsigned int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[wrp] = incl;
  __syncthreads();
  unsigned int base = 0;
  for (int w = 0; w < wrp; ++w) base += wsum[w];
  unsigned int before = base + incl - tot;   // keys in bins below this thread's first bin
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int b = threadIdx.x * PER + q;
    if (b < nb && before < remaining && remaining <= before + loc[q]) {
      // exactly one (thread, q) satisfies this
      st->prefix = prefix | ((unsigned long long)b << shift);
      st->remaining = remaining - before;
      st->bin_count = loc[q];
      if (pass == NPASS - 1) {
        const unsigned long long kk = prefix | ((unsigned long long)b << shift);
        scalars[1] = key_value(kk);          // the threshold: the K-th smallest log-likelihood
        st->ties_needed = remaining - before;  // keys equal to the threshold that retire (>= 1)
        st->ties_total = loc[q];               // keys equal to the threshold in the whole pool
      }
    }
    before += loc[q];
  }
  __syncthreads();
  // re-arm for the next pass / next round
  for (int i = threadIdx.x; i < SEL_BINS; i += SEL_THREADS) hist[i] = 0u;
  if (threadIdx.x == 0) *ticket = 0u;
}

// Per-slice counts over the GLOBAL key array: cnt[s] = (# keys < thr, # keys == thr) of slice s.  Slices are
// rank-aligned: rank r owns slices [r * spr, (r + 1) * spr), slice i of a rank covers its keys [i * SEL_SLICE, ...).
// Non-local slices are only needed to order ties; when every key equal to the threshold retires they are skipped.
__global__ void __launch_bounds__(SEL_THREADS) k_sel_count(const double* __restrict__ keys, long long Pl, int spr, int rank,
                                                           const SelectState* __restrict__ st,
                                                           const double* __restrict__ scalars, int2* __restrict__ cnt) {
  __shared__ int sb[SEL_THREADS / 32], se[SEL_THREADS / 32];
  const int s = blockIdx.x, r = s / spr, i = s % spr;
  const bool local = (r == rank);
  if (!local && st->ties_needed == st->ties_total) return;   // no tie ordering needed: nobody reads this slice's counts
  const unsigned long long thr = st->prefix;   // image of the threshold: all comparisons are on the integer images, so
                                               // they agree bit for bit with the radix passes (-0.0 / +0.0, NaNs)
  const long long k0 = (long long)r * Pl + (long long)i * SEL_SLICE;
  const long long k1 = min((long long)r * Pl + Pl, k0 + SEL_SLICE);
  int below = 0, equal = 0;
  for (long long k = k0 + threadIdx.x; k < k1; k += SEL_THREADS) {
    const unsigned long long v = key_image(keys[k]);
    below += (v < thr) ? 1 : 0;
    equal += (v == thr) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    below += __shfl_xor_sync(0xffffffffu, below, o);
    equal += __shfl_xor_sync(0xffffffffu, equal, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sb[threadIdx.x >> 5] = below;
    se[threadIdx.x >> 5] = equal;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int b = 0, e = 0;
    for (int w = 0; w < SEL_THREADS / 32; ++w) {
      b += sb[w];
      e += se[w];
    }
    cnt[s] = make_int2(b, e);
  }
}

// Local slices: retire flags + index-ordered compaction of the retired slots into sidx[0, c), c -> count[0].
__global__ void __launch_bounds__(SEL_THREADS) k_sel_flag(const double* __restrict__ keys, long long Pl, int spr, int rank,
                                                          const SelectState* __restrict__ st,
                                                          const double* __restrict__ scalars, const int2* __restrict__ cnt,
                                                          unsigned char* __restrict__ flags, int* __restrict__ sidx,
                                                          int* __restrict__ count) {
  __shared__ long long s_tie_base;   // keys == thr at lower global index than this slice
  __shared__ int s_ret_base;         // retired points in earlier local slices
  __shared__ int wsum[SEL_THREADS / 32];
  const int i = blockIdx.x, s = rank * spr + i;
  const unsigned long long thr = st->prefix;
  const long long need = st->ties_needed;
  const bool all_ties = st->ties_needed == st->ties_total;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  if (wrp == 0) {
    // one warp walks the earlier slices in order (a few hundred entries)
    long long tb = 0;
    if (!all_ties)
      for (int q = lane; q < s; q += 32) tb += cnt[q].y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tb += __shfl_xor_sync(0xffffffffu, tb, o);
    // retired in earlier LOCAL slices: below + the ties taken there (needs each slice's own tie base: serial, lane 0)
    int rb = 0;
    if (all_ties) {
      for (int q = lane; q < i; q += 32) rb += cnt[rank * spr + q].x + cnt[rank * spr + q].y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rb += __shfl_xor_sync(0xffffffffu, rb, o);
    } else if (lane == 0) {
      long long t = 0;
      for (int q = 0; q < rank * spr; ++q) t += cnt[q].y;
      for (int q = 0; q < i; ++q) {
        const int2 c = cnt[rank * spr + q];
        long long take = need - t;
        take = take < 0 ? 0 : (take > c.y ? c.y : take);
        rb += c.x + (int)take;
        t += c.y;
      }
    }
    if (lane == 0) {
      s_tie_base = tb;
      s_ret_base = rb;
    }
  }
  __syncthreads();
  constexpr int PER = SEL_SLICE / SEL_THREADS;   // 8 consecutive keys per thread
  const long long l0 = (long long)i * SEL_SLICE + (long long)threadIdx.x * PER;   // local index of this thread's first key
  unsigned long long v[PER];
  int nt = 0;
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const long long l = l0 + q;
    v[q] = (l < Pl) ? key_image(keys[(long long)rank * Pl + l]) : ~0ull;
    nt += (l < Pl && v[q] == thr) ? 1 : 0;
  }
  // exclusive scan of the tie counts -> which ties are taken
  auto block_excl = [&](int x) -> int {
    int incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    __syncthreads();
    if (lane == 31) wsum[wrp] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < wrp; ++w) base += wsum[w];
    return base + incl - x;
  };
  long long tie_rank = s_tie_base + (all_ties ? 0 : block_excl(nt));
  bool ret[PER];
  int nr = 0;
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const long long l = l0 + q;
    bool r = false;
    if (l < Pl) {
      if (v[q] < thr) r = true;
      else if (v[q] == thr) {
        r = all_ties || (tie_rank < need);
        ++tie_rank;
      }
    }
    ret[q] = r;
    nr += r ? 1 : 0;
  }
  int pos = s_ret_base + block_excl(nr);
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const long long l = l0 + q;
    if (l < Pl) {
      flags[l] = ret[q] ? 1 : 0;
      if (ret[q]) sidx[pos++] = (int)l;
    }
  }
  if (i == spr - 1 && threadIdx.x == SEL_THREADS - 1) count[0] = pos;   // the last thread of the last slice ends at c
}

// The retired VALUES of the round (every key below the threshold plus ties_needed copies of it), unordered: they are
// sorted afterwards on the evidence stream, off the critical path.  One warp-aggregated atomic per warp.
__global__ void __launch_bounds__(SEL_THREADS) k_sel_gather_values(const double* __restrict__ keys, long long n,
                                                                   const SelectState* __restrict__ st,
                                                                   const double* __restrict__ scalars, int Kr,
                                                                   double* __restrict__ out, unsigned int* __restrict__ cursor) {
  const unsigned long long thrk = st->prefix;
  const double thr = scalars[1];
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * SEL_THREADS;
  const long long n_up = (n + 31) / 32 * 32;
  for (long long i = (long long)blockIdx.x * SEL_THREADS + threadIdx.x; i < n_up; i += stride) {
    const bool take = (i < n) && (key_image(keys[i]) < thrk);
    const unsigned int m = __ballot_sync(0xffffffffu, take);
    if (m) {
      unsigned int base = 0;
      if (lane == 0) base = atomicAdd(cursor, (unsigned int)__popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (take) {
        const unsigned int p = base + __popc(m & ((1u << lane) - 1u));
        if (p < (unsigned int)Kr) out[p] = keys[i];
      }
    }
  }
  if (blockIdx.x == 0) {
    const unsigned int need = st->ties_needed;
    unsigned int base = 0;
    if (threadIdx.x == 0) base = atomicAdd(cursor, need);
    __shared__ unsigned int sb;
    if (threadIdx.x == 0) sb = base;
    __syncthreads();
    base = sb;
    for (unsigned int t = threadIdx.x; t < need; t += SEL_THREADS)
      if (base + t < (unsigned int)Kr) out[base + t] = thr;
  }
}

__global__ void k_sel_reset_cursor(unsigned int* cursor) { *cursor = 0u; }

}  // namespace

int slq_select_slices(long long Pl) { return (int)((Pl + SEL_SLICE - 1) / SEL_SLICE); }

// keys: the GLOBAL key array (world * Pl entries, rank-major); selects the K lowest, writes the threshold to scalars[1],
// this rank's retire flags / compacted slots / count.  8 launches, no host interaction, graph-capturable.
thy_status slq_select_launch(const SelectBuffers& b, const double* keys, long long Pl, int world, int rank, int K,
                             cudaStream_t s) {
  const long long n = Pl * world;
  int grid = sm_count() * 2;
  const long long need = (n + SEL_THREADS - 1) / SEL_THREADS;
  if (grid > need) grid = (int)need;
  if (grid < 1) grid = 1;
  for (int pass = 0; pass < NPASS; ++pass)
    k_sel_pass<<<grid, SEL_THREADS, 0, s>>>(keys, n, pass, K, b.state, b.hist, b.ticket, b.scalars);
  const int spr = slq_select_slices(Pl);
  k_sel_count<<<spr * world, SEL_THREADS, 0, s>>>(keys, Pl, spr, rank, b.state, b.scalars, b.cnt);
  k_sel_flag<<<spr, SEL_THREADS, 0, s>>>(keys, Pl, spr, rank, b.state, b.scalars, b.cnt, b.flags, b.sidx, b.count);
  count_launch(NPASS + 2);
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

// Unordered retired values of the round into out[0, Kr) (evidence stream).
thy_status slq_gather_retired_values(const SelectBuffers& b, const double* keys, long long n, int Kr, double* out,
                                     cudaStream_t s) {
  k_sel_reset_cursor<<<1, 1, 0, s>>>(b.cursor);
  int grid = sm_count();
  const long long need = (n + SEL_THREADS - 1) / SEL_THREADS;
  if (grid > need) grid = (int)need;
  if (grid < 1) grid = 1;
  k_sel_gather_values<<<grid, SEL_THREADS, 0, s>>>(keys, n, b.state, b.scalars, Kr, out, b.cursor);
  count_launch(2);
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

}  // namespace thy
