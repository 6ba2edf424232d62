// nccl_dyn.cc -- NCCL bound at run time with dlopen, so libslamgpu.so has no link-time
// dependency on a particular libnccl: inside a torch process the already loaded
// (torch-bundled) libnccl.so.2 is reused, elsewhere the system one is opened.
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "internal.h"

namespace {
struct Api {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
Api g_api;

int load(std::string *err) {
  if (g_api.h) return SLAMGPU_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    *err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
    return SLAMGPU_E_NCCL;
  }
  g_api.GetUniqueId = (decltype(g_api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_api.CommInitRank = (decltype(g_api.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_api.AllGather = (decltype(g_api.AllGather))dlsym(h, "ncclAllGather");
  g_api.CommDestroy = (decltype(g_api.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_api.GetErrorString = (decltype(g_api.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_api.GetUniqueId || !g_api.CommInitRank || !g_api.AllGather || !g_api.CommDestroy || !g_api.GetErrorString) {
    *err = "libnccl is missing a required symbol";
    return SLAMGPU_E_NCCL;
  }
  g_api.h = h;
  return SLAMGPU_OK;
}

int check(ncclResult_t r, const char *what, std::string *err) {
  if (r == ncclSuccess) return SLAMGPU_OK;
  *err = std::string(what) + ": " + g_api.GetErrorString(r);
  return SLAMGPU_E_NCCL;
}
}  // namespace

static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");

int sg_nccl_unique_id(void *id128, std::string *err) {
  if (!id128) { *err = "id buffer is NULL"; return SLAMGPU_E_INVALID; }
  int r = load(err);
  if (r != SLAMGPU_OK) return r;
  ncclUniqueId id;
  r = check(g_api.GetUniqueId(&id), "ncclGetUniqueId", err);
  if (r == SLAMGPU_OK) memcpy(id128, &id, 128);
  return r;
}

int sg_nccl_init(int nranks, int rank, const void *id128, void **comm, std::string *err) {
  int r = load(err);
  if (r != SLAMGPU_OK) return r;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t c = nullptr;
  r = check(g_api.CommInitRank(&c, nranks, id, rank), "ncclCommInitRank", err);
  if (r == SLAMGPU_OK) *comm = c;
  return r;
}

int sg_nccl_allgather(void *comm, const void *send, void *recv, size_t bytes, cudaStream_t s, std::string *err) {
  if (!comm) { *err = "no NCCL communicator on this ctx"; return SLAMGPU_E_STATE; }
  return check(g_api.AllGather(send, recv, bytes, ncclChar, (ncclComm_t)comm, s), "ncclAllGather", err);
}

void sg_nccl_destroy(void *comm) {
  if (comm && g_api.CommDestroy) g_api.CommDestroy((ncclComm_t)comm);
}
