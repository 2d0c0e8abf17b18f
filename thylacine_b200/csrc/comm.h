// Multi-GPU plumbing: one process per GPU, an NCCL communicator built from a unique id the host broadcasts
// (torch.distributed store / any out-of-band channel).  libnccl is loaded lazily with dlopen so that single-GPU users
// need no NCCL at all; in a torch process the already-loaded (torch-bundled) libnccl.so.2 is picked up.
#pragma once
#include "common.cuh"
#include <vector>

struct thy_comm_s {
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
};

namespace thy {
// Collectives on the given stream (no-ops when comm is null / world == 1 except all_gather, which copies).
thy_status comm_all_gather_f64(thy_comm_s* c, const double* send, double* recv, size_t count, cudaStream_t s);
thy_status comm_all_reduce_f64(thy_comm_s* c, double* buf, size_t count, int op /*0 sum, 2 max, 3 min*/, cudaStream_t s);
bool comm_is_live(const thy_comm_s* c);   // false once thy_comm_destroy has run on the handle
thy_status comm_all_reduce_i64(thy_comm_s* c, long long* buf, size_t count, int op, cudaStream_t s);
// Maps one cudaMalloc'd buffer of EVERY rank into this rank's address space: CUDA IPC between processes, plain peer
// access between the threads of one process.  Collective (every rank calls it with its own buffer).  out_ptrs[r] is rank
// r's buffer as seen from this rank (out_ptrs[rank] == local); opened[] lists the IPC mappings to close later.  *ok is
// the collective verdict: false on every rank if any rank could not map a peer (no peer access, IPC refused).
thy_status comm_map_peers(thy_comm_s* c, void* local, std::vector<void*>& out_ptrs, std::vector<void*>& opened, bool* ok,
                          cudaStream_t s);
}  // namespace thy
