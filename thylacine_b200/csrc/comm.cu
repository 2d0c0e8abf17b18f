// NCCL over NVLink 5 / NVSwitch, used only where the path has a real exchange step (DESIGN.md "Multi-GPU"):
// the SLQ global lowest-logL selection (all-gather of each rank's K smallest) and sample-moment sums (all-reduce).
// The reference has no collectives at all (single JVM process).
#include "comm.h"
#include <dlfcn.h>
#include <unistd.h>
#include <cstring>
#include <vector>
#include <mutex>
#include <set>

namespace thy {

namespace {
typedef int ncclResult_t;
struct ncclUniqueId { char internal[128]; };
typedef void* ncclComm_t;
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
constexpr int kNcclInt64 = 4, kNcclFloat64 = 8;

NcclApi& api() {
  static NcclApi a;
  if (a.lib) return a;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.lib) break;
  }
  if (!a.lib) return a;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.lib, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.lib, "ncclCommInitRank");
  a.CommInitAll = (decltype(a.CommInitAll))dlsym(a.lib, "ncclCommInitAll");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
  a.AllGather = (decltype(a.AllGather))dlsym(a.lib, "ncclAllGather");
  a.AllReduce = (decltype(a.AllReduce))dlsym(a.lib, "ncclAllReduce");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
  a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.AllReduce && a.GetErrorString;
  return a;
}

std::mutex g_live_mutex;
std::set<const thy_comm_s*> g_live_comms;   // handles between create and destroy (engines check before a late collective)

thy_status nccl_fail(ncclResult_t r, const char* what) {
  return fail(THY_ERR_NCCL, std::string(what) + ": " + (api().GetErrorString ? api().GetErrorString(r) : "nccl error"));
}
}  // namespace

bool comm_is_live(const thy_comm_s* c) {
  std::lock_guard<std::mutex> lock(g_live_mutex);
  return c && g_live_comms.count(c) != 0;
}
static void comm_register(const thy_comm_s* c, bool live) {
  std::lock_guard<std::mutex> lock(g_live_mutex);
  if (live) g_live_comms.insert(c); else g_live_comms.erase(c);
}

thy_status comm_all_gather_f64(thy_comm_s* c, const double* send, double* recv, size_t count, cudaStream_t s) {
  if (!c || c->world == 1) {
    if (send != recv) THY_CUDA_CHECK(cudaMemcpyAsync(recv, send, count * sizeof(double), cudaMemcpyDeviceToDevice, s));
    return THY_OK;
  }
  ncclResult_t r = api().AllGather(send, recv, count, kNcclFloat64, c->nccl_comm, s);
  return r == 0 ? THY_OK : nccl_fail(r, "ncclAllGather");
}

thy_status comm_all_reduce_f64(thy_comm_s* c, double* buf, size_t count, int op, cudaStream_t s) {
  if (!c || c->world == 1) return THY_OK;
  ncclResult_t r = api().AllReduce(buf, buf, count, kNcclFloat64, op, c->nccl_comm, s);
  return r == 0 ? THY_OK : nccl_fail(r, "ncclAllReduce");
}

thy_status comm_all_reduce_i64(thy_comm_s* c, long long* buf, size_t count, int op, cudaStream_t s) {
  if (!c || c->world == 1) return THY_OK;
  ncclResult_t r = api().AllReduce(buf, buf, count, kNcclInt64, op, c->nccl_comm, s);
  return r == 0 ? THY_OK : nccl_fail(r, "ncclAllReduce");
}

// ---- peer mapping (NVLink peer memory for the product's own exchange kernels) ----------------------------------------
namespace {
struct PeerRecord {             // 96 bytes = 12 doubles, exchanged with one NCCL all-gather
  long long pid;
  long long device;
  unsigned long long ptr;
  cudaIpcMemHandle_t handle;    // 64 bytes
  long long pad;
};
static_assert(sizeof(PeerRecord) == 96, "PeerRecord layout");
}  // namespace

thy_status comm_map_peers(thy_comm_s* c, void* local, std::vector<void*>& out_ptrs, std::vector<void*>& opened, bool* ok,
                          cudaStream_t s) {
  *ok = false;
  if (!c || c->world <= 1) {
    out_ptrs.assign(1, local);
    *ok = true;
    return THY_OK;
  }
  const int world = c->world, rank = c->rank;
  int dev = 0;
  THY_CUDA_CHECK(cudaGetDevice(&dev));
  PeerRecord mine{};
  mine.pid = (long long)getpid();
  mine.device = dev;
  mine.ptr = (unsigned long long)local;
  bool good = cudaIpcGetMemHandle(&mine.handle, local) == cudaSuccess;
  if (!good) cudaGetLastError();
  DevBuf<double> box;
  THY_TRY(box.reserve((size_t)world * 12));
  THY_CUDA_CHECK(cudaMemcpyAsync(box.ptr + (size_t)rank * 12, &mine, sizeof(mine), cudaMemcpyHostToDevice, s));
  THY_TRY(comm_all_gather_f64(c, box.ptr + (size_t)rank * 12, box.ptr, 12, s));
  std::vector<PeerRecord> all(world);
  THY_CUDA_CHECK(cudaMemcpyAsync(all.data(), box.ptr, (size_t)world * sizeof(PeerRecord), cudaMemcpyDeviceToHost, s));
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  out_ptrs.assign(world, nullptr);
  for (int r = 0; r < world && good; ++r) {
    if (r == rank) {
      out_ptrs[r] = local;
    } else if (all[r].pid == mine.pid) {   // another thread of this process: plain peer access
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, dev, (int)all[r].device) != cudaSuccess || !can) {
        good = false;
        break;
      }
      const cudaError_t e = cudaDeviceEnablePeerAccess((int)all[r].device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) good = false;
      cudaGetLastError();
      out_ptrs[r] = (void*)all[r].ptr;
    } else {
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        good = false;
        break;
      }
      out_ptrs[r] = p;
      opened.push_back(p);
    }
  }
  // the verdict must be the same everywhere: min over ranks
  DevBuf<long long> flag;
  THY_TRY(flag.reserve(1));
  const long long mineok = good ? 1 : 0;
  THY_CUDA_CHECK(cudaMemcpyAsync(flag.ptr, &mineok, sizeof(long long), cudaMemcpyHostToDevice, s));
  THY_TRY(comm_all_reduce_i64(c, flag.ptr, 1, 3 /* min */, s));
  long long allok = 0;
  THY_CUDA_CHECK(cudaMemcpyAsync(&allok, flag.ptr, sizeof(long long), cudaMemcpyDeviceToHost, s));
  THY_CUDA_CHECK(cudaStreamSynchronize(s));
  if (!allok) {
    for (void* p : opened) cudaIpcCloseMemHandle(p);
    opened.clear();
    out_ptrs.clear();
    cudaGetLastError();
  }
  *ok = allok != 0;
  return THY_OK;
}

}  // namespace thy

using namespace thy;

extern "C" {

thy_status thy_comm_unique_id(unsigned char out_id[128]) {
  if (!out_id) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  if (!api().ok) return fail(THY_ERR_NCCL, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  ncclResult_t r = api().GetUniqueId(&id);
  if (r != 0) return nccl_fail(r, "ncclGetUniqueId");
  std::memcpy(out_id, id.internal, 128);
  return THY_OK;
}

thy_status thy_comm_create(int rank, int world, const unsigned char unique_id[128], thy_comm_t* out) {
  if (!out) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  if (world < 1 || rank < 0 || rank >= world) return fail(THY_ERR_INVALID_ARGUMENT, "bad rank / world");
  auto* c = new (std::nothrow) thy_comm_s();
  if (!c) return fail(THY_ERR_INVALID_ARGUMENT, "allocation failed");
  c->rank = rank;
  c->world = world;
  if (world > 1) {
    if (require_device() != THY_OK) {
      delete c;
      return THY_ERR_NO_DEVICE;
    }
    if (!unique_id) {
      delete c;
      return fail(THY_ERR_INVALID_ARGUMENT, "unique id required for world > 1");
    }
    if (!api().ok) {
      delete c;
      return fail(THY_ERR_NCCL, "libnccl.so.2 could not be loaded");
    }
    ncclUniqueId id;
    std::memcpy(id.internal, unique_id, 128);
    ncclComm_t nc = nullptr;
    ncclResult_t r = api().CommInitRank(&nc, world, id, rank);
    if (r != 0) {
      delete c;
      return nccl_fail(r, "ncclCommInitRank");
    }
    c->nccl_comm = nc;
  }
  comm_register(c, true);
  *out = c;
  return THY_OK;
}

thy_status thy_comm_create_all(int n_devices, const int* devices, thy_comm_t* out) {
  if (!out || n_devices < 1) return fail(THY_ERR_INVALID_ARGUMENT, "bad arguments");
  THY_TRY(require_device());
  int have = 0;
  THY_CUDA_CHECK(cudaGetDeviceCount(&have));
  std::vector<int> devs(n_devices);
  for (int i = 0; i < n_devices; ++i) {
    devs[i] = devices ? devices[i] : i;
    if (devs[i] < 0 || devs[i] >= have) return fail(THY_ERR_INVALID_ARGUMENT, "device index out of range");
  }
  std::vector<ncclComm_t> comms(n_devices, nullptr);
  if (n_devices > 1) {
    if (!api().ok || !api().CommInitAll) return fail(THY_ERR_NCCL, "libnccl.so.2 could not be loaded");
    int cur = 0;
    cudaGetDevice(&cur);
    ncclResult_t r = api().CommInitAll(comms.data(), n_devices, devs.data());
    cudaSetDevice(cur);
    if (r != 0) return nccl_fail(r, "ncclCommInitAll");
  }
  for (int i = 0; i < n_devices; ++i) {
    auto* c = new (std::nothrow) thy_comm_s();
    if (!c) return fail(THY_ERR_INVALID_ARGUMENT, "allocation failed");
    c->rank = i;
    c->world = n_devices;
    c->nccl_comm = comms[i];
    comm_register(c, true);
    out[i] = c;
  }
  return THY_OK;
}

thy_status thy_device_count(int* out_count) {
  if (!out_count) return fail(THY_ERR_INVALID_ARGUMENT, "null pointer");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *out_count = n;
  return THY_OK;
}

thy_status thy_set_device(int device) {
  THY_TRY(require_device());
  THY_CUDA_CHECK(cudaSetDevice(device));
  return THY_OK;
}

thy_status thy_comm_destroy(thy_comm_t c) {
  if (c) comm_register(c, false);
  if (c && c->nccl_comm && api().ok) api().CommDestroy((ncclComm_t)c->nccl_comm);
  delete c;
  return THY_OK;
}

thy_status thy_comm_rank(thy_comm_t c, int* out_rank, int* out_world) {
  if (!c) return fail(THY_ERR_INVALID_ARGUMENT, "null handle");
  if (out_rank) *out_rank = c->rank;
  if (out_world) *out_world = c->world;
  return THY_OK;
}

/* Sum of a device fp64 buffer over all ranks (sample moments: count, sum x, sum x x^T). */
thy_status thy_comm_all_reduce_sum(thy_comm_t c, double* buf_dev, int64_t count, void* cuda_stream) {
  THY_TRY(comm_all_reduce_f64(c, buf_dev, (size_t)count, 0, (cudaStream_t)cuda_stream));
  return THY_OK;
}

}  // extern "C"
