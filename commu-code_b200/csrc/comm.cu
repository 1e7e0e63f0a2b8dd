// Gradient exchange for data-parallel training: ONE sum all-reduce of the flat fp32 gradient arena
// per optimizer step, issued directly on NCCL (NVLink 5 / NVSwitch) from this library.  NCCL is
// resolved at run time with dlopen so the library also loads on machines without it.
// Replaces the DDP bucketed all-reduce of the reference (train.py:155, 467-473), which fires once
// per micro-batch; accumulating locally and reducing once is mathematically identical.
#include <dlfcn.h>
#include <string.h>
#include "api_common.h"

// The handful of NCCL types / constants this file needs, declared locally (ABI-stable since NCCL 2.0) so that the
// library BUILDS without the NCCL development headers; the functions themselves are resolved with dlsym.
extern "C" {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclFloat32 = 7 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
}

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_api;
ncclComm_t g_comm = nullptr;
int g_world = 1, g_rank = 0;

int load_api(const char* path) {
  if (g_api.handle) return 0;
  const char* cands[3] = {path, "libnccl.so.2", "libnccl.so"};
  for (int i = 0; i < 3 && !g_api.handle; ++i)
    if (cands[i] && cands[i][0]) g_api.handle = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
  if (!g_api.handle) return cb_host::fail(COMMU_ERR_NCCL, "cannot dlopen libnccl: %s", dlerror());
#define LOAD(sym)                                                                      \
  g_api.sym = reinterpret_cast<decltype(g_api.sym)>(dlsym(g_api.handle, "nccl" #sym)); \
  if (!g_api.sym) return cb_host::fail(COMMU_ERR_NCCL, "libnccl lacks nccl" #sym);
  LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(AllReduce) LOAD(CommDestroy) LOAD(GetErrorString)
#undef LOAD
  return 0;
}
#define CB_CHECK_NCCL(expr)                                                               \
  do {                                                                                    \
    ncclResult_t _r = (expr);                                                             \
    if (_r != ncclSuccess)                                                                \
      return cb_host::fail(COMMU_ERR_NCCL, "%s failed: %s", #expr, g_api.GetErrorString(_r)); \
  } while (0)
}  // namespace

extern "C" {

// Fills `id_out` (128 bytes) on the calling rank; broadcast it to the other ranks by any means.
int commu_comm_unique_id(const char* nccl_path, void* id_out) {
  int rc = load_api(nccl_path);
  if (rc) return rc;
  ncclUniqueId id;
  CB_CHECK_NCCL(g_api.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

int commu_comm_init(const char* nccl_path, const void* id_in, int rank, int world) {
  int rc = load_api(nccl_path);
  if (rc) return rc;
  CB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm_init: bad rank %d / world %d", rank, world);
  if (g_comm) return cb_host::fail(COMMU_ERR_INVALID, "comm_init: communicator already initialised");
  ncclUniqueId id;
  memcpy(&id, id_in, sizeof(id));
  CB_CHECK_NCCL(g_api.CommInitRank(&g_comm, world, id, rank));
  g_world = world;
  g_rank = rank;
  return 0;
}

int commu_allreduce_sum_f32(float* buf, int64_t n, void* stream) {
  CB_REQUIRE(buf && n > 0, "allreduce: bad args");
  if (g_world == 1 && !g_comm) return 0;
  CB_REQUIRE(g_comm != nullptr, "allreduce: communicator not initialised");
  CB_CHECK_NCCL(g_api.AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, g_comm, (cudaStream_t)stream));
  return 0;
}

int commu_comm_destroy(void) {
  if (g_comm) {
    CB_CHECK_NCCL(g_api.CommDestroy(g_comm));
    g_comm = nullptr;
    g_world = 1;
  }
  return 0;
}

}  // extern "C"
