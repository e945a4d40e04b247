// abi_comm.cu -- ci_comm_* / ci_allgather of include/ci_b200.h: the ONE collective of the path
// (an all-gather of the per-draw result rows at the end of a sharded fit) for
// callers that bind the C ABI without torch.distributed.  A thin wrapper over NCCL, which is
// resolved AT RUN TIME with dlopen("libnccl.so.2"): a process that already loaded NCCL (torch
// ships one) shares that copy, a process that never creates a ci_comm needs no NCCL at all, and
// the library has no link-time dependency on it.  The reference has no counterpart (single
// process, single device: causalimpact_lib.py:342-345).
#include <dlfcn.h>

#include "ci_host.cuh"

namespace {

struct nccl_uid { char internal[CI_COMM_ID_BYTES]; };      // == ncclUniqueId (NCCL_UNIQUE_ID_BYTES 128)
typedef struct ncclComm* nccl_comm_t;
enum { NCCL_SUCCESS = 0, NCCL_INT8 = 0 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(nccl_uid*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char* names[] = {getenv("CI_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return nullptr;
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle, "ncclAllGather"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
  api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(dlsym(api.handle, "ncclGetVersion"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy) {
    dlclose(api.handle);
    api.handle = nullptr;
    return nullptr;
  }
  return &api;
}

int nccl_fail(NcclApi* a, const char* what, int rc) {
  return fail(CI_ERR_CUDA, "%s failed: %s", what,
              (a && a->GetErrorString) ? a->GetErrorString(rc) : "NCCL error");
}

}  // namespace

struct ci_comm {
  nccl_comm_t comm = nullptr;
  int device = -1, rank = 0, nranks = 1;
};

extern "C" {

int ci_comm_get_unique_id(uint8_t* id) {
  if (!id) return fail(CI_ERR_INVALID, "null argument");
  NcclApi* a = nccl();
  if (!a) return fail(CI_ERR_UNSUPPORTED, "NCCL not found (dlopen libnccl.so.2 failed: %s)", dlerror());
  nccl_uid u;
  const int rc = a->GetUniqueId(&u);
  if (rc != NCCL_SUCCESS) return nccl_fail(a, "ncclGetUniqueId", rc);
  memcpy(id, u.internal, CI_COMM_ID_BYTES);
  return CI_OK;
}

int ci_comm_create(ci_ctx* ctx, const uint8_t* id, int rank, int nranks, ci_comm** out) {
  if (!ctx || !id || !out) return fail(CI_ERR_INVALID, "null argument");
  *out = nullptr;
  if (nranks < 1 || rank < 0 || rank >= nranks)
    return fail(CI_ERR_INVALID, "rank %d out of range [0,%d)", rank, nranks);
  NcclApi* a = nccl();
  if (!a) return fail(CI_ERR_UNSUPPORTED, "NCCL not found (dlopen libnccl.so.2 failed: %s)", dlerror());
  CU_TRY(cudaSetDevice(ctx->device));
  nccl_uid u;
  memcpy(u.internal, id, CI_COMM_ID_BYTES);
  ci_comm* cm = new ci_comm();
  cm->device = ctx->device; cm->rank = rank; cm->nranks = nranks;
  const int rc = a->CommInitRank(&cm->comm, nranks, u, rank);
  if (rc != NCCL_SUCCESS) { delete cm; return nccl_fail(a, "ncclCommInitRank", rc); }
  *out = cm;
  return CI_OK;
}

int ci_allgather(ci_comm* cm, const void* send_d, void* recv_d, size_t bytes_per_rank, void* stream) {
  if (!cm || !send_d || !recv_d) return fail(CI_ERR_INVALID, "null argument");
  NcclApi* a = nccl();
  if (!a) return fail(CI_ERR_STATE, "NCCL is not loaded");
  CU_TRY(cudaSetDevice(cm->device));
  const int rc = a->AllGather(send_d, recv_d, bytes_per_rank, NCCL_INT8, cm->comm,
                              static_cast<cudaStream_t>(stream));
  if (rc != NCCL_SUCCESS) return nccl_fail(a, "ncclAllGather", rc);
  return CI_OK;
}

int ci_comm_destroy(ci_comm* cm) {
  if (!cm) return CI_OK;
  NcclApi* a = nccl();
  if (a && cm->comm) {
    cudaSetDevice(cm->device);
    a->CommDestroy(cm->comm);
  }
  delete cm;
  return CI_OK;
}

}  // extern "C"
