// Pieces of the SLQ engine shared between slq.cu (round control), slq_select.cu (threshold selection) and
// slq_evidence.cu (prior-mass simulation, quadrature, resampling).
#pragma once
#include "kernels.h"
#include "philox.cuh"

namespace thy {

constexpr int SEL_SLICE = 2048;   // keys per slice of the count / flag kernels

struct SelectState {
  unsigned long long prefix;   // decided digits of the K-th smallest key's image; after the last pass: the image itself
  unsigned int remaining;      // rank still to resolve inside the prefix (1-based)
  unsigned int bin_count;      // keys in the chosen bin of the last finished pass
  unsigned int ties_needed;    // keys equal to the threshold that retire this round
  unsigned int ties_total;     // keys equal to the threshold in the whole pool
};

struct SelectBuffers {
  SelectState* state;
  unsigned int* hist;     // [6][2048] one histogram per radix pass, zero between rounds
  unsigned int* ticket;   // [1] grid-barrier counter (monotonic)
  unsigned int* cursor;   // [1] append cursor of the retired values, zero between rounds
  int2* cnt;              // [world * slices_per_rank] (below, equal) per slice
  unsigned char* flags;   // [Pl] 1 = retired this round
  int* sidx;              // [K] retired local slots, ascending
  int* count;             // [1] retired local points
  double* scalars;        // engine scalars; [1] = threshold
};

#ifdef __CUDACC__
// A uniformly chosen SURVIVOR of the round (a live point that is not being retired): rejection sampling on the retire
// flags with the counter RNG (at most half of a rank's pool retires per round, so two draws suffice on average); the
// probing fallback after 32 draws is unreachable in practice and only keeps the function total.  Shared by the fused
// walk kernel and the per-step spawn kernel, which therefore start every walker from the same point.
__device__ __forceinline__ long long slq_pick_survivor(uint64_t seed, uint64_t key, uint64_t round, long long Pl,
                                                       const unsigned char* __restrict__ flags) {
  long long a = 0;
  for (uint32_t t = 0; t < 16; ++t) {
    const RngDraw u = rng_uniform2(seed, key, t, RNG_SLQ_PICK, round);
    a = (long long)(u.u0 * (double)Pl);
    if (a >= Pl) a = Pl - 1;
    if (!flags[a]) return a;
    a = (long long)(u.u1 * (double)Pl);
    if (a >= Pl) a = Pl - 1;
    if (!flags[a]) return a;
  }
  for (long long k = 0; k < Pl; ++k) {
    a = (a + 1 == Pl) ? 0 : a + 1;
    if (!flags[a]) break;
  }
  return a;
}
#endif

// ---- NVLink peer-memory exchange of the round's new keys (slq_p2p.cu) ----
// Every rank owns one exchange buffer that all ranks map (CUDA IPC / peer access):
//   pairs   [2 parities][world sources][cap] x 16 bytes  (global key index as int64 bits, new log-likelihood)
//   counts  [2 parities][world sources] x int64           how many pairs a source staged
//   flags   [world sources] x uint64                      last round whose walk that source has finished, + 1
struct P2pView {
  const unsigned long long* peer_base;   // device array [world]: rank r's exchange buffer as mapped on THIS rank (null: off)
  int world, rank;
  long long cap;
};
__host__ __device__ inline size_t p2p_pair_off(const P2pView& v, int parity, int src, long long w) {
  return (((size_t)parity * v.world + src) * (size_t)v.cap + (size_t)w) * 16;
}
__host__ __device__ inline size_t p2p_count_off(const P2pView& v, int parity, int src) {
  return (size_t)2 * v.world * (size_t)v.cap * 16 + ((size_t)parity * v.world + src) * 8;
}
__host__ __device__ inline size_t p2p_flag_off(const P2pView& v, int src) {
  return (size_t)2 * v.world * (size_t)v.cap * 16 + (size_t)2 * v.world * 8 + (size_t)src * 8;
}
inline size_t p2p_bytes(int world, long long cap) { return (size_t)2 * world * (size_t)cap * 16 + (size_t)3 * world * 8; }
thy_status slq_p2p_wait_and_apply(const P2pView& v, unsigned char* my_base, const long long* ctl, double* scalars,
                                  double* keys_all, cudaStream_t s);
// push this rank's new keys (sidx / ll after the walk) to every rank, then publish count and flag
thy_status slq_p2p_signal(const P2pView& v, long long Pl, const int* sidx, const double* ll, const int* count_dev,
                          const long long* ctl, cudaStream_t s);

int slq_select_slices(long long Pl);
thy_status slq_select_launch(const SelectBuffers& b, const double* keys, long long Pl, int world, int rank, int K,
                             double* vals, cudaStream_t s);

// ---- exact sort of a round's retired values (slq_sort.cu) ----
constexpr int SLQ_SORT_MAX_BUCKETS = 16384;
struct BucketSortBuffers {
  unsigned long long* minmax;   // [2] rest state (~0, 0)
  unsigned int* count;          // [SLQ_SORT_MAX_BUCKETS] rest state 0
  unsigned int* cursor;         // [SLQ_SORT_MAX_BUCKETS] rest state 0
  unsigned int* offset;         // [SLQ_SORT_MAX_BUCKETS + 1]
  double* tmp;                  // [n_max]
};
thy_status slq_bucket_sort_init(const BucketSortBuffers& b, cudaStream_t s);   // puts the counters in their rest state
// out[0, n) = in[0, n) ascending (in and out distinct); six plain launches on s, no host interaction
thy_status slq_bucket_sort(const BucketSortBuffers& b, const double* in, double* out, int n, cudaStream_t s);

// ---- prior-mass simulation and quadrature (slq_evidence.cu) ----
constexpr int EV_STATE = 12;   // doubles of carried state per abscissa
// state: [0] logX assigned to the next point, [1] logX_{i-2}, [2] logX_{i-1}, [3] l_{i-1}, [4] pending points (0, 1, 2+),
//        [5] running max m, [6] sum exp(term - m), [7] sum exp(term - m) * l, [8] ln t_{i-2}, [9] ln t_{i-1}
struct EvidenceBuffers {
  double* lt;        // [A][stride] ln t_j of the current round
  double* segsum;    // [A][nseg_max] per-segment sums of ln t
  double* part;      // [A][nseg_max][3] per-segment (m, s, sl) partial log-sum-exp
  unsigned int* tickets;  // [A]
  double* state;     // [A][EV_STATE]
  double* out;       // [A][2] (ln Z, H)
  long long stride;
  int nseg_max;
};
// live points when the i-th retired point (0-based, global retirement index) leaves: live-phase rounds retire K at a
// time from a pool of P (the j-th of a round sees P - j), the drain removes one at a time.
__host__ __device__ inline long long slq_n_live(long long i, long long P, long long K, long long live) {
  return (i < live) ? (P - (i % K)) : (P - (i - live));
}
// Adds the Kr retired log-likelihoods r[0, Kr) (ascending; global retirement indices it0 ...) to every abscissa's
// quadrature.  live = number of points retired in the live phase (LLONG_MAX while it is still running).  it0_dev, when
// not null, overrides it0 with a device-resident counter (the launch then replays from a CUDA graph).
thy_status slq_evidence_launch(const EvidenceBuffers& e, int A, int Kr, long long P, long long K, long long live,
                               long long it0, const long long* it0_dev, uint64_t seed, const double* r, int finalize,
                               cudaStream_t s);
// ln X_i of abscissa a for retired points [i0, i0 + n) into out_dev (device), given ln X_{i0} in *carry_dev (updated).
thy_status slq_logx_launch(const EvidenceBuffers& e, int a, long long i0, int n, long long P, long long K, long long live,
                           uint64_t seed, double* carry_dev, double* out_dev, cudaStream_t s);

}  // namespace thy
