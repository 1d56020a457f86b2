// Data-parallel entry points of the C ABI (include/fov360.h): one process per GPU, one NCCL
// communicator per process, one summed allreduce of the flat fp32 gradient bucket per training step over
// NVLink 5 / NVSwitch.  The reference has no distributed code at all; these calls are what a maintainer binds
// next to model.fit.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): inside a PyTorch process that resolves to the copy torch
// already loaded, elsewhere to the system library; libfov360.so itself carries no link-time NCCL dependency, so it
// loads on boxes without NCCL and fov_dp_init reports FOV_ERR_UNSUPPORTED there.
#include <dlfcn.h>
#include <string.h>

#include "fov_common.cuh"
#include "fov_internal.h"

namespace {

// the few NCCL types / enums this file needs (nccl.h: ncclUniqueId is 128 opaque bytes; ncclFloat32 = 7, ncclSum = 0)
struct NcclId { char internal[128]; };
typedef struct ncclComm* NcclComm;
enum { kNcclFloat = 7, kNcclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

NcclApi g_nccl;
NcclComm g_comm = nullptr;     // the library's only long-lived state besides the debug switches
int g_rank = 0, g_world = 1;

int load_nccl() {
  if (g_nccl.handle) return FOV_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    fov_set_error("fov_dp: libnccl.so.2 not found (%s)", dlerror());
    return FOV_ERR_UNSUPPORTED;
  }
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = (int (*)(NcclId*))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (int (*)(NcclComm*, int, NcclId, int))dlsym(h, "ncclCommInitRank");
  a.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllReduce");
  a.Broadcast = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(h, "ncclBroadcast");
  a.CommDestroy = (int (*)(NcclComm))dlsym(h, "ncclCommDestroy");
  a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.Broadcast || !a.CommDestroy) {
    fov_set_error("fov_dp: libnccl is missing a required symbol");
    dlclose(h);
    return FOV_ERR_UNSUPPORTED;
  }
  g_nccl = a;
  return FOV_OK;
}

int nccl_check(int rc, const char* what) {
  if (rc == 0) return FOV_OK;
  fov_set_error("fov_dp: %s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error");
  return FOV_ERR_CUDA;
}

}  // namespace

extern "C" int fov_dp_unique_id_bytes(void) { return (int)sizeof(NcclId); }

extern "C" int fov_dp_get_unique_id(void* id_out) {
  FOV_CHECK_ARG(id_out != nullptr, "id_out is NULL");
  int rc = load_nccl();
  if (rc) return rc;
  NcclId id;
  if ((rc = nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId"))) return rc;
  memcpy(id_out, &id, sizeof(id));
  return FOV_OK;
}

extern "C" int fov_dp_init(const void* unique_id, int rank, int world) {
  FOV_CHECK_ARG(unique_id != nullptr && world >= 1 && rank >= 0 && rank < world, "bad rank / world / id");
  FOV_CHECK_ARG(g_comm == nullptr, "already initialised (fov_dp_destroy first)");
  int rc = load_nccl();
  if (rc) return rc;
  NcclId id;
  memcpy(&id, unique_id, sizeof(id));
  NcclComm comm = nullptr;
  if ((rc = nccl_check(g_nccl.CommInitRank(&comm, world, id, rank), "ncclCommInitRank"))) return rc;
  g_comm = comm; g_rank = rank; g_world = world;
  return FOV_OK;
}

extern "C" int fov_dp_world(void) { return g_comm ? g_world : 1; }
extern "C" int fov_dp_rank(void) { return g_comm ? g_rank : 0; }

extern "C" int fov_dp_allreduce(float* flat, size_t n, void* stream) {
  FOV_CHECK_ARG(flat != nullptr && n > 0, "bad args");
  FOV_CHECK_ARG(g_comm != nullptr, "fov_dp_init has not been called");
  return nccl_check(g_nccl.AllReduce(flat, flat, n, kNcclFloat, kNcclSum, g_comm, (cudaStream_t)stream), "ncclAllReduce");
}

extern "C" int fov_dp_broadcast(float* flat, size_t n, int root, void* stream) {
  FOV_CHECK_ARG(flat != nullptr && n > 0 && root >= 0, "bad args");
  FOV_CHECK_ARG(g_comm != nullptr, "fov_dp_init has not been called");
  FOV_CHECK_ARG(root < g_world, "root outside the communicator");
  return nccl_check(g_nccl.Broadcast(flat, flat, n, kNcclFloat, root, g_comm, (cudaStream_t)stream), "ncclBroadcast");
}

extern "C" int fov_dp_destroy(void) {
  if (!g_comm) return FOV_OK;
  const int rc = nccl_check(g_nccl.CommDestroy(g_comm), "ncclCommDestroy");
  g_comm = nullptr; g_rank = 0; g_world = 1;
  return rc;
}
