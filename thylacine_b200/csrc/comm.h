// Multi-GPU plumbing: one process per GPU, an NCCL communicator built from a unique id the host broadcasts
// (torch.distributed store / any out-of-band channel).  libnccl is loaded lazily with dlopen so that single-GPU users
// need no NCCL at all; in a torch process the already-loaded (torch-bundled) libnccl.so.2 is picked up.
#pragma once
#include "common.cuh"

struct thy_comm_s {
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
};

namespace thy {
// Collectives on the given stream (no-ops when comm is null / world == 1 except all_gather, which copies).
thy_status comm_all_gather_f64(thy_comm_s* c, const double* send, double* recv, size_t count, cudaStream_t s);
thy_status comm_all_reduce_f64(thy_comm_s* c, double* buf, size_t count, int op /*0 sum, 2 max, 3 min*/, cudaStream_t s);
thy_status comm_all_reduce_i64(thy_comm_s* c, long long* buf, size_t count, int op, cudaStream_t s);
}  // namespace thy
