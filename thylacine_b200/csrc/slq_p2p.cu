// SLQ across GPUs without a per-round collective: new log-likelihoods travel over NVLink peer memory.
//
// The selection step needs every rank to know every live point's log-likelihood (slq_select.cu).  Round 2 first did that
// with one NCCL all-gather of all P keys per round (8 MB at 10^6 live points: 57-81 us on 2-8 B200s, a sixth of a round at
// 8 ranks) although only the K replaced points change.  Here each rank keeps its own full copy `keys_all`, and the walk
// kernel that creates a replacement writes (global index, new log-likelihood) straight into a staging area in EVERY
// rank's memory (16-byte stores over NVLink from the thread that holds the value; slq_walk.cu commit step).  A round is
//   k_p2p_wait    spin (bounded) until every peer's flag says its walk of the previous round is finished
//   k_p2p_apply   scatter the staged pairs of all sources into the local keys_all
//   k_slq_select  ... k_slq_walk (stages the new keys on all ranks) ...
//   k_p2p_signal  system-scope fence, publish the staged count, then the flag, on every rank
// Staging is double-buffered by round parity: a rank can only reach the walk of round R + 2 (which reuses round R's
// parity) after every peer has signalled round R + 1, i.e. after every peer has applied round R's pairs.  The waiter and
// the signaller run on different GPUs; nothing here waits on a kernel of the same device.  The reference has no
// counterpart (single JVM process); the step being distributed is SlqEngine.scala:139-143 (`keys.min`).
#include "slq_internal.h"

namespace thy {

namespace {

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void k_p2p_wait(P2pView v, unsigned char* my_base, const long long* __restrict__ ctl, double* __restrict__ scalars) {
  const int r = threadIdx.x;
  if (r >= v.world || r == v.rank) return;
  const unsigned long long target = (unsigned long long)ctl[1];
  const volatile unsigned long long* flag = reinterpret_cast<const volatile unsigned long long*>(my_base + p2p_flag_off(v, r));
  const unsigned long long t0 = global_timer_ns();
  while (*flag < target) {
    __nanosleep(200);
    if (global_timer_ns() - t0 > 5000000000ull) {   // 5 s: a peer died; record it and let the host report
      scalars[7] = 1.0;
      break;
    }
  }
  __threadfence_system();
}

__global__ void __launch_bounds__(256) k_p2p_apply(P2pView v, const unsigned char* my_base, const long long* __restrict__ ctl,
                                                   double* __restrict__ keys_all) {
  const int parity = (int)(ctl[1] & 1);
  const int src = blockIdx.y;
  const long long count = __ldcg(reinterpret_cast<const long long*>(my_base + p2p_count_off(v, parity, src)));
  const double2* pairs = reinterpret_cast<const double2*>(my_base + p2p_pair_off(v, parity, src, 0));
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < count; w += (long long)gridDim.x * blockDim.x) {
    const double2 pr = __ldcg(pairs + w);
    keys_all[__double_as_longlong(pr.x)] = pr.y;
  }
}

// (sidx[w], ll[sidx[w]]) for w < c are the round's replaced slots and their new log-likelihoods.
__global__ void __launch_bounds__(256) k_p2p_push(P2pView v, long long Pl, const int* __restrict__ sidx, const double* __restrict__ ll,
                                                  const int* __restrict__ count_dev, const long long* __restrict__ ctl) {
  const int c = *count_dev;
  const int parity_next = (int)((ctl[1] + 1) & 1);
  const int r = blockIdx.y;
  double2* dst = reinterpret_cast<double2*>(reinterpret_cast<unsigned char*>(v.peer_base[r]) +
                                           p2p_pair_off(v, parity_next, v.rank, 0));
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < c; w += gridDim.x * blockDim.x) {
    const int slot = sidx[w];
    dst[w] = make_double2(__longlong_as_double((long long)v.rank * Pl + slot), ll[slot]);
  }
}

__global__ void k_p2p_signal(P2pView v, const int* __restrict__ count_dev, const long long* __restrict__ ctl) {
  const int r = threadIdx.x;
  if (r >= v.world) return;
  const long long round = ctl[1];
  const int parity_next = (int)((round + 1) & 1);
  unsigned char* base = reinterpret_cast<unsigned char*>(v.peer_base[r]);
  __threadfence_system();
  *reinterpret_cast<volatile long long*>(base + p2p_count_off(v, parity_next, v.rank)) = (long long)*count_dev;
  __threadfence_system();
  *reinterpret_cast<volatile unsigned long long*>(base + p2p_flag_off(v, v.rank)) = (unsigned long long)(round + 1);
}

}  // namespace

thy_status slq_p2p_wait_and_apply(const P2pView& v, unsigned char* my_base, const long long* ctl, double* scalars,
                                  double* keys_all, cudaStream_t s) {
  k_p2p_wait<<<1, 32, 0, s>>>(v, my_base, ctl, scalars);
  const unsigned gx = (unsigned)std::min<long long>(64, (v.cap + 255) / 256);
  k_p2p_apply<<<dim3(gx, (unsigned)v.world), 256, 0, s>>>(v, my_base, ctl, keys_all);
  count_launch(2);
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

thy_status slq_p2p_signal(const P2pView& v, long long Pl, const int* sidx, const double* ll, const int* count_dev,
                          const long long* ctl, cudaStream_t s) {
  const unsigned gx = (unsigned)std::min<long long>(32, (v.cap + 255) / 256);
  k_p2p_push<<<dim3(gx, (unsigned)v.world), 256, 0, s>>>(v, Pl, sidx, ll, count_dev, ctl);
  k_p2p_signal<<<1, 32, 0, s>>>(v, count_dev, ctl);
  count_launch(2);
  THY_CUDA_CHECK(cudaGetLastError());
  return THY_OK;
}

}  // namespace thy
