// abi_comm.cu -- ci_comm_* / ci_allgather of include/ci_b200.h: the ONE collective of the path
// (an all-gather of the per-draw result rows at the end of a sharded fit) for
// callers that bind the C ABI without torch.distributed.  A thin wrapper over NCCL, which is
// resolved AT RUN TIME with dlopen("libnccl.so.2"): a process that already loaded NCCL (torch
// ships one) shares that copy, a process that never creates a ci_comm needs no NCCL at all, and
// the library has no link-time dependency on it.  The reference has no counterpart (single
// process, single device: causalimpact_lib.py:342-345).
#include <dlfcn.h>

#include "ci_host.cuh"
#include "ci_impact.cuh"

namespace {

struct nccl_uid { char internal[CI_COMM_ID_BYTES]; };      // == ncclUniqueId (NCCL_UNIQUE_ID_BYTES 128)
typedef struct ncclComm* nccl_comm_t;
enum { NCCL_SUCCESS = 0, NCCL_INT8 = 0, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(nccl_uid*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
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
  api.Send = reinterpret_cast<decltype(api.Send)>(dlsym(api.handle, "ncclSend"));
  api.Recv = reinterpret_cast<decltype(api.Recv)>(dlsym(api.handle, "ncclRecv"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(api.handle, "ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(api.handle, "ncclGroupEnd"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.handle, "ncclAllReduce"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
  api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(dlsym(api.handle, "ncclGetVersion"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy || !api.Send ||
      !api.Recv || !api.GroupStart || !api.GroupEnd || !api.AllReduce) {
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
  // ci_impact_sharded_d: exchange windows in peer memory.  Every rank owns winT / winC (the
  // column blocks of ITS time steps) and maps the other ranks' windows through CUDA IPC; the
  // rows kernel of rank a stores block a of rank g's window directly over NVLink.  Capacities
  // are the same on every rank (each can compute every rank's need), so all ranks regrow --
  // and re-exchange the handles -- in the same call.
  size_t capT = 0, capC = 0, cap_parts = 0;
  void* winT = nullptr; void* winC = nullptr; void* parts = nullptr;
  void* peerT[ci::IMP_MAX_RANKS] = {}; void* peerC[ci::IMP_MAX_RANKS] = {};
  bool peer_ok = true;              // false: IPC unavailable -> grouped ncclSend / ncclRecv
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
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
  cudaSetDevice(cm->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < cm->nranks && r < ci::IMP_MAX_RANKS; ++r) {
    if (r == cm->rank) continue;
    if (cm->peerT[r]) cudaIpcCloseMemHandle(cm->peerT[r]);
    if (cm->peerC[r]) cudaIpcCloseMemHandle(cm->peerC[r]);
  }
  if (cm->winT) cudaFree(cm->winT);
  if (cm->winC) cudaFree(cm->winC);
  if (cm->parts) cudaFree(cm->parts);
  if (cm->ev_fork) cudaEventDestroy(cm->ev_fork);
  if (cm->ev_join) cudaEventDestroy(cm->ev_join);
  if (a && cm->comm) a->CommDestroy(cm->comm);
  delete cm;
  return CI_OK;
}

}  // extern "C"

namespace {

// contiguous balanced split of range(n) over ws ranks (the package's shard.split_range)
inline void split_range(int n, int ws, int r, int* start, int* count) {
  const int base = n / ws, extra = n % ws;
  *start = r * base + (r < extra ? r : extra);
  *count = base + (r < extra ? 1 : 0);
}

// CI_B200_TRACE=1: device time of every step of ci_impact_sharded_d, printed by rank 0 to stderr
// (synchronises -- a tuning aid, tools/prof_sharded.py)
struct StepTrace {
  bool on = false;
  cudaStream_t st = nullptr;
  cudaEvent_t ev[12];
  const char* name[12];
  int n = 0;
  StepTrace(cudaStream_t s, bool enable) : on(enable), st(s) {}
  void mark(const char* what) {
    if (!on || n >= 12) return;
    cudaEventCreate(&ev[n]);
    cudaEventRecord(ev[n], st);
    name[n++] = what;
  }
  void report() {
    if (!on) return;
    cudaStreamSynchronize(st);
    for (int i = 1; i < n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      fprintf(stderr, "[ci_impact_sharded_d] %-22s %8.1f us\n", name[i], ms * 1e3f);
    }
    for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
  }
};

// (Re)allocate the exchange windows for `needT` / `needC` bytes (the same numbers on every rank)
// and map every peer's.  Collective; synchronises, but only when a window grows.  On any
// failure on any rank the communicator falls back to ncclSend / ncclRecv for good.
int ensure_windows(ci_comm* cm, NcclApi* api, size_t needT, size_t needC, cudaStream_t st) {
  if (!cm->peer_ok || (needT <= cm->capT && needC <= cm->capC)) return CI_OK;
  const int ws = cm->nranks, me = cm->rank;
  struct Pack { cudaIpcMemHandle_t hT, hC; int status; int pad[3]; };
  Pack mine{};
  Pack* dev = nullptr;
  std::vector<Pack> all(ws);
  CU_TRY(cudaStreamSynchronize(st));
  CU_TRY(cudaMalloc(&dev, sizeof(Pack) * (ws + 1)));
  const bool had_windows = cm->winT != nullptr;
  for (int r = 0; r < ws; ++r) {
    if (r != me && cm->peerT[r]) cudaIpcCloseMemHandle(cm->peerT[r]);
    if (r != me && cm->peerC[r]) cudaIpcCloseMemHandle(cm->peerC[r]);
    cm->peerT[r] = cm->peerC[r] = nullptr;
  }
  if (had_windows) {
    // an exported block must not be freed while another process still maps it: wait until every
    // rank has closed its mappings (a tiny all-gather as the barrier; regrowth is rare)
    CU_TRY(cudaMemcpyAsync(dev + ws, &mine, sizeof(Pack), cudaMemcpyHostToDevice, st));
    const int brc = api->AllGather(dev + ws, dev, sizeof(Pack), NCCL_INT8, cm->comm, st);
    if (brc != NCCL_SUCCESS) { cudaFree(dev); return nccl_fail(api, "ncclAllGather", brc); }
    CU_TRY(cudaStreamSynchronize(st));
  }
  if (cm->winT) cudaFree(cm->winT);
  if (cm->winC) cudaFree(cm->winC);
  cm->winT = cm->winC = nullptr;
  auto grow = [](size_t need) { return ((need + need / 4 + (2u << 20) - 1) >> 21) << 21; };
  cm->capT = grow(needT > cm->capT ? needT : cm->capT);
  cm->capC = grow(needC > cm->capC ? needC : cm->capC);
  mine.status = 0;
  if (cudaMalloc(&cm->winT, cm->capT) != cudaSuccess || cudaMalloc(&cm->winC, cm->capC) != cudaSuccess ||
      cudaIpcGetMemHandle(&mine.hT, cm->winT) != cudaSuccess ||
      cudaIpcGetMemHandle(&mine.hC, cm->winC) != cudaSuccess) {
    mine.status = 1;
    cudaGetLastError();
  }
  CU_TRY(cudaMemcpyAsync(dev + ws, &mine, sizeof(Pack), cudaMemcpyHostToDevice, st));
  int rc = api->AllGather(dev + ws, dev, sizeof(Pack), NCCL_INT8, cm->comm, st);
  if (rc != NCCL_SUCCESS) { cudaFree(dev); return nccl_fail(api, "ncclAllGather", rc); }
  CU_TRY(cudaMemcpyAsync(all.data(), dev, sizeof(Pack) * ws, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  int bad = 0;
  for (int r = 0; r < ws; ++r) bad |= all[r].status;
  if (!bad) {
    for (int r = 0; r < ws; ++r) {
      if (r == me) { cm->peerT[r] = cm->winT; cm->peerC[r] = cm->winC; continue; }
      if (cudaIpcOpenMemHandle(&cm->peerT[r], all[r].hT, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&cm->peerC[r], all[r].hC, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        bad = 1;
        cudaGetLastError();
        break;
      }
    }
  }
  // agree on the outcome: one more tiny all-gather of the status
  mine.status = bad;
  CU_TRY(cudaMemcpyAsync(dev + ws, &mine, sizeof(Pack), cudaMemcpyHostToDevice, st));
  rc = api->AllGather(dev + ws, dev, sizeof(Pack), NCCL_INT8, cm->comm, st);
  if (rc != NCCL_SUCCESS) { cudaFree(dev); return nccl_fail(api, "ncclAllGather", rc); }
  CU_TRY(cudaMemcpyAsync(all.data(), dev, sizeof(Pack) * ws, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  cudaFree(dev);
  for (int r = 0; r < ws; ++r) bad |= all[r].status;
  if (bad || getenv("CI_B200_NO_PEER")) cm->peer_ok = false;
  return CI_OK;
}

template <typename R>
int impact_sharded(ci_ctx* c, ci_comm* cm, NcclApi* api, ci::ImpactDev d, const ci::ShardCounts& sc,
                   int S_total, const void* traj_d, const void* mean_part_d, const double* obs_d,
                   const uint8_t* per_d, void* mean_d, double* out_d, cudaStream_t st) {
  using namespace ci;
  const int ws = cm->nranks, me = cm->rank;
  const int T = d.T, Tc = T - d.t_c0, S_loc = sc.n[me];
  int tb, tn, cb, cn;
  split_range(T, ws, me, &tb, &tn);
  split_range(Tc, ws, me, &cb, &cn);
  const int head_me = me == 0 ? IMP_STATS : 0;
  int me_off = 0;
  for (int r = 0; r < me; ++r) me_off += sc.n[r];
  // window sizes: the largest block any rank receives (rank 0 has the most columns)
  const size_t needT = (size_t)((T + ws - 1) / ws) * S_total * sizeof(R);
  const size_t needC = (size_t)((Tc + ws - 1) / ws + IMP_STATS) * S_total * sizeof(double);
  if (int rc = ensure_windows(cm, api, needT, needC, st)) return rc;
  if (cm->cap_parts < (size_t)ws * T * sizeof(R)) {
    CU_TRY(cudaStreamSynchronize(st));
    if (cm->parts) cudaFree(cm->parts);
    cm->cap_parts = (size_t)ws * T * sizeof(R) * 2;
    CU_TRY(cudaMalloc(&cm->parts, cm->cap_parts));
  }
  if (!cm->ev_fork) {
    CU_TRY(cudaEventCreateWithFlags(&cm->ev_fork, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&cm->ev_join, cudaEventDisableTiming));
  }
  const bool peer = cm->peer_ok;
  R* parts = static_cast<R*>(cm->parts);                      // [ws][T]
  double* series_d = out_d;
  double* summ_d = out_d + (size_t)T * IMP_SERIES_COLS;
  static const bool trace_on = getenv("CI_B200_TRACE") != nullptr;
  StepTrace tr(st, trace_on && me == 0);
  tr.mark("start");
  CU_TRY(cudaMemsetAsync(out_d, 0, ((size_t)T * IMP_SERIES_COLS + IMP_SUMMARY_LEN) * sizeof(double), st));
  const R* blkT; const double* blkC;                           // column blocks of MY time steps
  int rc = NCCL_SUCCESS;
  if (peer) {
    // ---- fused: the rows kernel stores every transposed tile into its owner's window ----
    if (S_loc > 0) {
      PeerDest pd{};
      pd.ws = ws; pd.me_off = me_off; pd.n_me = S_loc;
      pd.T_base = T / ws; pd.T_extra = T % ws; pd.C_base = Tc / ws; pd.C_extra = Tc % ws;
      for (int r = 0; r < ws; ++r) { pd.T[r] = cm->peerT[r]; pd.C[r] = cm->peerC[r]; }
      int row_ctas, nseg;
      impact_rows_grid(S_loc, T, d.t_c0, false, &row_ctas, &nseg);
      k_impact_rows<R><<<row_ctas * nseg, 32 * IMP_WARPS, 0, st>>>(
          static_cast<const R*>(traj_d), nullptr, obs_d, per_d, d, nullptr, nullptr, nullptr, nullptr,
          nullptr, nullptr, row_ctas, impact_seg_len(), pd);
      CU_TRY(cudaGetLastError());
      c->launches++;
    }
    tr.mark("k_impact_rows -> peer windows");
    // the mean parts go to everyone; this collective is also the barrier of the exchange: it
    // completes here only after every rank's rows kernel -- and so its stores into my windows --
    // has completed
    rc = api->AllGather(mean_part_d, parts, (size_t)T * sizeof(R), NCCL_INT8, cm->comm, st);
    if (rc != NCCL_SUCCESS) return nccl_fail(api, "ncclAllGather", rc);
    tr.mark("all-gather of mean parts");
    blkT = static_cast<const R*>(cm->winT);
    blkC = static_cast<const double*>(cm->winC);
  } else {
    // ---- without peer mapping: local transposed arrays, ONE grouped ncclSend / ncclRecv ----
    CU_TRY(c->i_trT.reserve((size_t)T * (S_loc > 0 ? S_loc : 1) * sizeof(R)));
    CU_TRY(c->x_pack.reserve((size_t)(IMP_STATS + Tc) * (S_loc > 0 ? S_loc : 1) * sizeof(double)));
    CU_TRY(c->x_recvT.reserve((size_t)(tn > 0 ? tn : 1) * S_total * sizeof(R)));
    CU_TRY(c->x_recvC.reserve((size_t)(cn + head_me > 0 ? cn + head_me : 1) * S_total * sizeof(double)));
    R* trT = static_cast<R*>(c->i_trT.p);                     // [T][S_loc]
    double* pack = static_cast<double*>(c->x_pack.p);         // [5 + Tc][S_loc]: stats, then cumT
    R* recvT = static_cast<R*>(c->x_recvT.p);
    double* recvC = static_cast<double*>(c->x_recvC.p);
    if (S_loc > 0) {
      int row_ctas, nseg;
      impact_rows_grid(S_loc, T, d.t_c0, false, &row_ctas, &nseg);
      k_impact_rows<R><<<row_ctas * nseg, 32 * IMP_WARPS, 0, st>>>(
          static_cast<const R*>(traj_d), nullptr, obs_d, per_d, d, trT, pack + (size_t)IMP_STATS * S_loc,
          pack, nullptr, nullptr, nullptr, row_ctas, impact_seg_len(), PeerDest{});
      CU_TRY(cudaGetLastError());
      c->launches++;
    }
    tr.mark("k_impact_rows");
    rc = api->GroupStart();
    if (rc != NCCL_SUCCESS) return nccl_fail(api, "ncclGroupStart", rc);
    size_t offT = 0, offC = 0;
    for (int g = 0; g < ws && rc == NCCL_SUCCESS; ++g) {
      int gtb, gtn, gcb, gcn;
      split_range(T, ws, g, &gtb, &gtn);
      split_range(Tc, ws, g, &gcb, &gcn);
      const size_t sendT = (size_t)gtn * S_loc * sizeof(R), recvTb = (size_t)tn * sc.n[g] * sizeof(R);
      const int ghead = g == 0 ? IMP_STATS : 0;
      const size_t sendC = (size_t)(gcn + ghead) * S_loc * sizeof(double);
      const size_t recvCb = (size_t)(cn + head_me) * sc.n[g] * sizeof(double);
      if (sendT) rc = api->Send(trT + (size_t)gtb * S_loc, sendT, NCCL_INT8, g, cm->comm, st);
      if (rc == NCCL_SUCCESS && recvTb)
        rc = api->Recv(reinterpret_cast<char*>(recvT) + offT, recvTb, NCCL_INT8, g, cm->comm, st);
      if (rc == NCCL_SUCCESS && sendC)
        rc = api->Send(pack + (size_t)(g == 0 ? 0 : IMP_STATS + gcb) * S_loc, sendC, NCCL_INT8, g,
                       cm->comm, st);
      if (rc == NCCL_SUCCESS && recvCb)
        rc = api->Recv(reinterpret_cast<char*>(recvC) + offC, recvCb, NCCL_INT8, g, cm->comm, st);
      if (rc == NCCL_SUCCESS)
        rc = api->Send(mean_part_d, (size_t)T * sizeof(R), NCCL_INT8, g, cm->comm, st);
      if (rc == NCCL_SUCCESS)
        rc = api->Recv(parts + (size_t)g * T, (size_t)T * sizeof(R), NCCL_INT8, g, cm->comm, st);
      offT += recvTb; offC += recvCb;
    }
    const int rc2 = api->GroupEnd();
    if (rc != NCCL_SUCCESS) return nccl_fail(api, "ncclSend/ncclRecv", rc);
    if (rc2 != NCCL_SUCCESS) return nccl_fail(api, "ncclGroupEnd", rc2);
    tr.mark("grouped send/recv");
    blkT = recvT; blkC = recvC;
  }
  // the mean over all draws, then (rank 0, on the context's own stream, beside the column jobs)
  // the mean-derived columns: the mean CTA of k_impact_rows alone
  k_mean_combine<R><<<(T + 255) / 256, 256, 0, st>>>(parts, sc, S_total, T, static_cast<R*>(mean_d));
  c->launches++;
  ImpactDev dall = d;
  dall.S = S_total;
  const bool fork = me == 0 && c->stream != st;
  if (me == 0) {
    cudaStream_t ms = fork ? c->stream : st;
    if (fork) {
      CU_TRY(cudaEventRecord(cm->ev_fork, st));
      CU_TRY(cudaStreamWaitEvent(ms, cm->ev_fork, 0));
    }
    k_impact_rows<R><<<1, 32 * IMP_WARPS, 0, ms>>>(
        nullptr, static_cast<const R*>(mean_d), obs_d, per_d, dall, nullptr, nullptr, nullptr,
        series_d, summ_d, nullptr, 1, impact_seg_len(), PeerDest{});
    c->launches++;
    if (fork) CU_TRY(cudaEventRecord(cm->ev_join, ms));
  }
  size_t bytes; int in_smem, nt;
  select_launch_cfg(c, S_total, sizeof(double), &nt, &bytes, &in_smem);
  const int gx = (S_total + 255) / 256;
  const R* colT = blkT; const double* colC = blkC; const double* stats = nullptr;
  ColBlocks blocks{};
  if (in_smem) {
    // the column jobs read the blocks as they arrived; only the 5 statistics rows of rank 0 are
    // laid side by side (impact_summary_block walks them whole)
    blocks.ws = ws; blocks.rowsT = tn; blocks.rowsC = cn + head_me; blocks.headC = head_me;
    for (int r = 0; r < ws; ++r) blocks.n[r] = sc.n[r];
    if (me == 0) {
      CU_TRY(c->x_allC.reserve((size_t)IMP_STATS * S_total * sizeof(double)));
      k_merge_blocks<double><<<dim3(gx, IMP_STATS), 256, 0, st>>>(
          blkC, static_cast<double*>(c->x_allC.p), cn + head_me, sc, S_total);
      c->launches++;
      stats = static_cast<const double*>(c->x_allC.p);
    }
  } else {
    // more draws than the shared-memory select holds: contiguous columns for the global-memory select
    CU_TRY(c->x_allT.reserve((size_t)(tn > 0 ? tn : 1) * S_total * sizeof(R)));
    CU_TRY(c->x_allC.reserve((size_t)(cn + head_me > 0 ? cn + head_me : 1) * S_total * sizeof(double)));
    if (tn > 0) {
      k_merge_blocks<R><<<dim3(gx, tn), 256, 0, st>>>(blkT, static_cast<R*>(c->x_allT.p), tn, sc, S_total);
      c->launches++;
    }
    if (cn + head_me > 0) {
      k_merge_blocks<double><<<dim3(gx, cn + head_me), 256, 0, st>>>(
          blkC, static_cast<double*>(c->x_allC.p), cn + head_me, sc, S_total);
      c->launches++;
    }
    colT = static_cast<const R*>(c->x_allT.p);
    colC = static_cast<const double*>(c->x_allC.p) + (size_t)head_me * S_total;
    stats = me == 0 ? static_cast<const double*>(c->x_allC.p) : nullptr;
  }
  CU_TRY(cudaGetLastError());
  tr.mark("mean combine (+ merge)");
  const ImpactCols jc{tb, tn, cb, cn, me == 0 ? 1 : 0};
  const int jobs = cn + tn + (me == 0 ? IMP_STATS + 1 : 0);
  if (jobs > 0) {
    auto kern = k_impact_jobs<R>;
    CU_TRY(set_smem(kern, (uint32_t)bytes));
    kern<<<jobs, nt, bytes, st>>>(colT, colC, stats, obs_d, dall, series_d, summ_d, in_smem, nullptr,
                                  jc, blocks);
    c->launches++;
  }
  CU_TRY(cudaGetLastError());
  if (fork) CU_TRY(cudaStreamWaitEvent(st, cm->ev_join, 0));
  tr.mark("k_impact_jobs (+ mean row)");
  // every entry of out was written by exactly one rank (zeros elsewhere): the sum is a gather
  rc = api->AllReduce(out_d, out_d, (size_t)T * IMP_SERIES_COLS + IMP_SUMMARY_LEN, NCCL_FLOAT64,
                      NCCL_SUM, cm->comm, st);
  if (rc != NCCL_SUCCESS) return nccl_fail(api, "ncclAllReduce", rc);
  tr.mark("all-reduce");
  tr.report();
  return CI_OK;
}

}  // namespace

extern "C" int ci_impact_sharded_d(ci_ctx* c, ci_comm* cm, const ci_impact_args* a,
                                   const int32_t* counts, const void* traj_d,
                                   const void* mean_part_d, const double* observed,
                                   const uint8_t* period, void* mean_d, double* out_d,
                                   void* stream) {
  if (!c || !cm || !a || !counts || !mean_part_d || !observed || !period || !mean_d || !out_d)
    return fail(CI_ERR_INVALID, "null argument");
  NcclApi* api = nccl();
  if (!api) return fail(CI_ERR_STATE, "NCCL is not loaded");
  if (cm->device != c->device) return fail(CI_ERR_INVALID, "the communicator lives on another device");
  if (cm->nranks > ci::IMP_MAX_RANKS)
    return fail(CI_ERR_UNSUPPORTED, "at most %d ranks", ci::IMP_MAX_RANKS);
  ci::ShardCounts sc{};
  sc.ws = cm->nranks;
  long long total = 0;
  for (int r = 0; r < cm->nranks; ++r) {
    if (counts[r] < 0) return fail(CI_ERR_INVALID, "counts[%d] is negative", r);
    sc.n[r] = counts[r];
    total += counts[r];
  }
  if (counts[0] < 1) return fail(CI_ERR_INVALID, "rank 0 must hold at least one draw");
  if (total > 0x7fffffff) return fail(CI_ERR_INVALID, "too many draws");
  if (a->S != counts[cm->rank]) return fail(CI_ERR_INVALID, "args->S must equal counts[rank]");
  if (a->S > 0 && !traj_d) return fail(CI_ERR_INVALID, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ci_impact_args chk = *a;
  if (chk.S < 1) chk.S = 1;                   // a rank without draws still takes part
  ci::ImpactDev d{};
  const double* obs_d; const uint8_t* per_d;
  if (int rc = cih_impact_prepare(c, &chk, observed, period, st, &d, &obs_d, &per_d)) return rc;
  d.S = a->S;
  const int S = (int)total;
  if (a->dtype == CI_F64)
    return impact_sharded<double>(c, cm, api, d, sc, S, traj_d, mean_part_d, obs_d, per_d, mean_d, out_d, st);
  return impact_sharded<float>(c, cm, api, d, sc, S, traj_d, mean_part_d, obs_d, per_d, mean_d, out_d, st);
}
