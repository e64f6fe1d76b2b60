// libias_comm.so: the embedding all-gather of the front end as plain-C entry points over NCCL (NVLink 5 / NVSwitch).
//
// Replaces vicreg.FullGatherLayer (vicreg.py:79-95; intended call site vicreg.py:38-39): forward = ncclAllGather of
// every rank's packed [B_local, 2*D] embeddings in rank order, backward = ncclReduceScatter(sum) of the gradients
// w.r.t. the gathered tensor (the reference's all_reduce + slice).  One communicator per process (one process per
// GPU); the unique id is created on rank 0 and distributed by the host (ias_b200/dist.py uses torch.distributed).
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ias_b200.h"

namespace {
char* cerr_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
int cerr(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(cerr_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace

#define IAS_NCCL(call)                                                                              \
  do {                                                                                              \
    ncclResult_t r__ = (call);                                                                      \
    if (r__ != ncclSuccess) return cerr(IAS_ERR_NCCL, "%s failed: %s", #call, ncclGetErrorString(r__)); \
  } while (0)

static_assert(sizeof(ncclUniqueId) == IAS_NCCL_ID_BYTES, "ncclUniqueId size changed");

extern "C" const char* ias_comm_last_error(void) { return cerr_buf(); }

extern "C" int ias_comm_unique_id(void* id128_host) {
  if (!id128_host) return cerr(IAS_ERR_INVALID, "ias_comm_unique_id: NULL");
  ncclUniqueId id;
  IAS_NCCL(ncclGetUniqueId(&id));
  memcpy(id128_host, &id, sizeof(id));
  return IAS_OK;
}

extern "C" int ias_comm_init(const void* id128_host, int rank, int world, void** comm) {
  if (!id128_host || !comm || world < 1 || rank < 0 || rank >= world)
    return cerr(IAS_ERR_INVALID, "ias_comm_init: rank=%d world=%d", rank, world);
  ncclUniqueId id;
  memcpy(&id, id128_host, sizeof(id));
  ncclComm_t c;
  IAS_NCCL(ncclCommInitRank(&c, world, id, rank));
  *comm = reinterpret_cast<void*>(c);
  return IAS_OK;
}

extern "C" int ias_comm_destroy(void* comm) {
  if (!comm) return IAS_OK;
  IAS_NCCL(ncclCommDestroy(reinterpret_cast<ncclComm_t>(comm)));
  return IAS_OK;
}

extern "C" int ias_comm_allgather(void* comm, const float* local, float* all, size_t count, ias_stream_t stream) {
  if (!comm || !local || !all || count == 0) return cerr(IAS_ERR_INVALID, "ias_comm_allgather: bad arguments");
  IAS_NCCL(ncclAllGather(local, all, count, ncclFloat, reinterpret_cast<ncclComm_t>(comm),
                         reinterpret_cast<cudaStream_t>(stream)));
  return IAS_OK;
}

extern "C" int ias_comm_reduce_scatter(void* comm, const float* all, float* local, size_t count, ias_stream_t stream) {
  if (!comm || !local || !all || count == 0) return cerr(IAS_ERR_INVALID, "ias_comm_reduce_scatter: bad arguments");
  IAS_NCCL(ncclReduceScatter(all, local, count, ncclFloat, ncclSum, reinterpret_cast<ncclComm_t>(comm),
                             reinterpret_cast<cudaStream_t>(stream)));
  return IAS_OK;
}
