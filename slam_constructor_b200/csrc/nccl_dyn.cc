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
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
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
  g_api.Send = (decltype(g_api.Send))dlsym(h, "ncclSend");
  g_api.Recv = (decltype(g_api.Recv))dlsym(h, "ncclRecv");
  g_api.GroupStart = (decltype(g_api.GroupStart))dlsym(h, "ncclGroupStart");
  g_api.GroupEnd = (decltype(g_api.GroupEnd))dlsym(h, "ncclGroupEnd");
  g_api.CommDestroy = (decltype(g_api.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_api.GetErrorString = (decltype(g_api.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_api.GetUniqueId || !g_api.CommInitRank || !g_api.AllGather || !g_api.CommDestroy || !g_api.GetErrorString ||
      !g_api.Send || !g_api.Recv || !g_api.GroupStart || !g_api.GroupEnd) {
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

// point-to-point transfers of one group: {peer, device pointer, bytes}; sends and receives are posted together
int sg_nccl_exchange(void *comm, const SgXfer *sends, int n_sends, const SgXfer *recvs, int n_recvs, cudaStream_t s, std::string *err) {
  if (!comm) { *err = "no NCCL communicator on this ctx"; return SLAMGPU_E_STATE; }
  if (n_sends + n_recvs == 0) return SLAMGPU_OK;
  int r = check(g_api.GroupStart(), "ncclGroupStart", err);
  for (int k = 0; k < n_sends && r == SLAMGPU_OK; ++k)
    r = check(g_api.Send(sends[k].ptr, sends[k].bytes, ncclChar, sends[k].peer, (ncclComm_t)comm, s), "ncclSend", err);
  for (int k = 0; k < n_recvs && r == SLAMGPU_OK; ++k)
    r = check(g_api.Recv(recvs[k].ptr, recvs[k].bytes, ncclChar, recvs[k].peer, (ncclComm_t)comm, s), "ncclRecv", err);
  int e = check(g_api.GroupEnd(), "ncclGroupEnd", err);
  return r != SLAMGPU_OK ? r : e;
}

void sg_nccl_destroy(void *comm) {
  if (comm && g_api.CommDestroy) g_api.CommDestroy((ncclComm_t)comm);
}
