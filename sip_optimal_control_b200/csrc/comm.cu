// Multi-GPU side of the C ABI (include/sipoc.h, "several devices"): batch shards and the
// ONE exchange the path has -- the per-iteration all-reduce of the four statistics
// {sum of squared residual norms, max residual norm, #failed, #problems}.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the engine itself has no link
// dependency on it, and a single-GPU user never loads it.  One all-gather of 4 doubles per
// rank followed by a one-warp fold kernel gives the sum / max / sum / sum combination in a
// single collective (a SUM and a MAX all-reduce would be two latency-bound round trips).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdint>
#include <cstdio>
#include <mutex>
#include <new>
#include <string>

#include "../../include/sipoc.h"

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {getenv("SIPOC_NCCL_LIBRARY"), "libnccl.so.2", "libnccl.so"};
    for (const char *name : names) {
      if (name == nullptr || *name == '\0') continue;
      api.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (api.handle != nullptr) break;
    }
    if (api.handle == nullptr) return;
#define BIND(field, symbol)                                                     \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, symbol)); \
  if (api.field == nullptr) return;
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommInitAll, "ncclCommInitAll");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(AllGather, "ncclAllGather");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
    BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    api.ok = true;
  });
  return api;
}

// gathered: [world][4] -> stats[4] = {sum, max, sum, sum} over the ranks.
__global__ void fold_stats_kernel(const double *gathered, int world, double *stats) {
  const int slot = threadIdx.x;
  if (slot >= 4) return;
  double acc = gathered[slot];
  for (int r = 1; r < world; ++r) {
    const double v = gathered[r * 4 + slot];
    acc = slot == 1 ? fmax(acc, v) : acc + v;
  }
  stats[slot] = acc;
}

}  // namespace

struct sipoc_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  double *gathered = nullptr;  // device [world][4]
  std::string last_error;
};

namespace {
sipoc_error finish_create(sipoc_comm *c) {
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  const cudaError_t err = cudaMalloc(&c->gathered, sizeof(double) * 4 * c->world);
  if (prev >= 0) cudaSetDevice(prev);
  return err == cudaSuccess ? SIPOC_OK : SIPOC_OUT_OF_MEMORY;
}
}  // namespace

extern "C" {

sipoc_error sipoc_shard_range(int64_t total, int rank, int world, int64_t *begin, int64_t *end) {
  if (world <= 0 || rank < 0 || rank >= world || total < 0 || !begin || !end)
    return SIPOC_INVALID_ARGUMENT;
  const int64_t base = total / world, extra = total % world;
  *begin = rank * base + (rank < extra ? rank : extra);
  *end = *begin + base + (rank < extra ? 1 : 0);
  return SIPOC_OK;
}

sipoc_error sipoc_comm_unique_id(void *id_out) {
  if (id_out == nullptr) return SIPOC_INVALID_ARGUMENT;
  NcclApi &n = nccl();
  if (!n.ok) return SIPOC_UNSUPPORTED;
  ncclUniqueId id;
  if (n.GetUniqueId(&id) != ncclSuccess) return SIPOC_CUDA_ERROR;
  static_assert(sizeof(id) == SIPOC_COMM_ID_BYTES, "NCCL unique id size");
  memcpy(id_out, &id, sizeof(id));
  return SIPOC_OK;
}

sipoc_error sipoc_comm_create(const void *id_in, int rank, int world, int device,
                              sipoc_comm **out) {
  if (out == nullptr) return SIPOC_INVALID_ARGUMENT;
  *out = nullptr;
  if (id_in == nullptr || world <= 0 || rank < 0 || rank >= world) return SIPOC_INVALID_ARGUMENT;
  NcclApi &n = nccl();
  if (!n.ok) return SIPOC_UNSUPPORTED;
  sipoc_comm *c = new (std::nothrow) sipoc_comm();
  if (c == nullptr) return SIPOC_OUT_OF_MEMORY;
  if (device < 0) cudaGetDevice(&device);
  c->rank = rank;
  c->world = world;
  c->device = device;
  ncclUniqueId id;
  memcpy(&id, id_in, sizeof(id));
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  const ncclResult_t rc = n.CommInitRank(&c->comm, world, id, rank);
  if (prev >= 0) cudaSetDevice(prev);
  if (rc != ncclSuccess) {
    std::fprintf(stderr, "sipoc_comm_create: %s\n", n.GetErrorString(rc));
    delete c;
    return SIPOC_CUDA_ERROR;
  }
  const sipoc_error e = finish_create(c);
  if (e != SIPOC_OK) {
    n.CommDestroy(c->comm);
    delete c;
    return e;
  }
  *out = c;
  return SIPOC_OK;
}

sipoc_error sipoc_comm_create_all(const int *device_ids, int n_dev, sipoc_comm **out) {
  if (device_ids == nullptr || out == nullptr || n_dev <= 0 || n_dev > 64)
    return SIPOC_INVALID_ARGUMENT;
  NcclApi &n = nccl();
  if (!n.ok) return SIPOC_UNSUPPORTED;
  ncclComm_t comms[64];
  const ncclResult_t rc = n.CommInitAll(comms, n_dev, device_ids);
  if (rc != ncclSuccess) {
    std::fprintf(stderr, "sipoc_comm_create_all: %s\n", n.GetErrorString(rc));
    return SIPOC_CUDA_ERROR;
  }
  for (int i = 0; i < n_dev; ++i) {
    sipoc_comm *c = new (std::nothrow) sipoc_comm();
    if (c == nullptr) return SIPOC_OUT_OF_MEMORY;
    c->comm = comms[i];
    c->rank = i;
    c->world = n_dev;
    c->device = device_ids[i];
    const sipoc_error e = finish_create(c);
    if (e != SIPOC_OK) return e;
    out[i] = c;
  }
  return SIPOC_OK;
}

void sipoc_comm_destroy(sipoc_comm *c) {
  if (c == nullptr) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  if (c->comm != nullptr && nccl().ok) nccl().CommDestroy(c->comm);
  cudaFree(c->gathered);
  if (prev >= 0) cudaSetDevice(prev);
  delete c;
}

int sipoc_comm_rank(const sipoc_comm *c) { return c == nullptr ? -1 : c->rank; }
int sipoc_comm_size(const sipoc_comm *c) { return c == nullptr ? 0 : c->world; }

sipoc_error sipoc_comm_group_begin(void) {
  return nccl().ok && nccl().GroupStart() == ncclSuccess ? SIPOC_OK : SIPOC_CUDA_ERROR;
}
sipoc_error sipoc_comm_group_end(void) {
  return nccl().ok && nccl().GroupEnd() == ncclSuccess ? SIPOC_OK : SIPOC_CUDA_ERROR;
}

// In place on the rank's device double[4]; enqueued on `stream` (graph-capturable).
sipoc_error sipoc_comm_allgather_stats(sipoc_comm *c, const double *stats, void *stream) {
  if (c == nullptr || stats == nullptr) return SIPOC_INVALID_ARGUMENT;
  NcclApi &n = nccl();
  if (!n.ok) return SIPOC_UNSUPPORTED;
  const ncclResult_t rc = n.AllGather(stats, c->gathered, 4, ncclDouble, c->comm,
                                      static_cast<cudaStream_t>(stream));
  if (rc != ncclSuccess) {
    c->last_error = n.GetErrorString(rc);
    return SIPOC_CUDA_ERROR;
  }
  return SIPOC_OK;
}

sipoc_error sipoc_comm_fold_stats(sipoc_comm *c, double *stats, void *stream) {
  if (c == nullptr || stats == nullptr) return SIPOC_INVALID_ARGUMENT;
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != c->device) cudaSetDevice(c->device);
  fold_stats_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(c->gathered, c->world, stats);
  const cudaError_t err = cudaGetLastError();
  if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
  return err == cudaSuccess ? SIPOC_OK : SIPOC_CUDA_ERROR;
}

sipoc_error sipoc_comm_allreduce_stats(sipoc_comm *c, double *stats, void *stream) {
  const sipoc_error rc = sipoc_comm_allgather_stats(c, stats, stream);
  return rc != SIPOC_OK ? rc : sipoc_comm_fold_stats(c, stats, stream);
}

}  // extern "C"
