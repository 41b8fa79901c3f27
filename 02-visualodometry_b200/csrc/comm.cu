// comm.cu — optional NCCL communicator (one process per GPU, NVLink/NVSwitch).
//
// The only exchange on the hot path is the all-reduce of the 32 H/b/chi terms per
// Gauss-Newton round. libnccl.so.2 is resolved with dlopen so the library has no
// link-time NCCL dependency and, inside a torch process, shares torch's already-loaded NCCL.
#include "vo_common.cuh"

#include <dlfcn.h>

namespace {

typedef struct { char internal[128]; } nccl_unique_id;  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef int (*fn_get_unique_id)(nccl_unique_id*);
typedef int (*fn_comm_init_rank)(void**, int, nccl_unique_id, int);
typedef int (*fn_comm_destroy)(void*);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_get_error_string)(int);

struct NcclApi {
  void* handle = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_get_error_string error_string = nullptr;
  bool tried = false;
};

NcclApi g_nccl;
const int kNcclFloat64 = 8;  // ncclFloat64
const int kNcclSum = 0;      // ncclSum

bool load_nccl() {
  if (g_nccl.tried) return g_nccl.handle != nullptr;
  g_nccl.tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) return false;
  g_nccl.get_unique_id = (fn_get_unique_id)dlsym(g_nccl.handle, "ncclGetUniqueId");
  g_nccl.comm_init_rank = (fn_comm_init_rank)dlsym(g_nccl.handle, "ncclCommInitRank");
  g_nccl.comm_destroy = (fn_comm_destroy)dlsym(g_nccl.handle, "ncclCommDestroy");
  g_nccl.all_reduce = (fn_all_reduce)dlsym(g_nccl.handle, "ncclAllReduce");
  g_nccl.error_string = (fn_get_error_string)dlsym(g_nccl.handle, "ncclGetErrorString");
  if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.comm_destroy || !g_nccl.all_reduce) {
    dlclose(g_nccl.handle);
    g_nccl.handle = nullptr;
    return false;
  }
  return true;
}

const char* nccl_err(int r) { return g_nccl.error_string ? g_nccl.error_string(r) : "nccl failure"; }

}  // namespace

extern "C" {

int vo_comm_unique_id(uint8_t id[128]) {
  if (!id) return VO_ERR_INVALID;
  if (!load_nccl()) return VO_ERR_NCCL;
  nccl_unique_id u;
  int r = g_nccl.get_unique_id(&u);
  if (r != 0) return VO_ERR_NCCL;
  memcpy(id, u.internal, 128);
  return VO_OK;
}

int vo_ctx_comm_init(vo_ctx* ctx, int n_ranks, int rank, const uint8_t id[128]) {
  if (!ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return VO_ERR_INVALID;
  if (ctx->nccl_comm) return vo_set_error(ctx, VO_ERR_STATE, "vo_ctx_comm_init", "communicator already set");
  if (!load_nccl()) return vo_set_error(ctx, VO_ERR_NCCL, "dlopen(libnccl.so.2)", dlerror());
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  nccl_unique_id u;
  memcpy(u.internal, id, 128);
  void* comm = nullptr;
  int r = g_nccl.comm_init_rank(&comm, n_ranks, u, rank);
  if (r != 0) return vo_set_error(ctx, VO_ERR_NCCL, "ncclCommInitRank", nccl_err(r));
  ctx->nccl_comm = comm;
  ctx->n_ranks = n_ranks;
  ctx->rank = rank;
  return VO_OK;
}

int vo_ctx_comm_destroy(vo_ctx* ctx) {
  if (!ctx) return VO_ERR_INVALID;
  if (ctx->nccl_comm && g_nccl.comm_destroy) {
    cudaStreamSynchronize(ctx->stream);
    g_nccl.comm_destroy(ctx->nccl_comm);
  }
  ctx->nccl_comm = nullptr;
  ctx->n_ranks = 1;
  ctx->rank = 0;
  return VO_OK;
}

int vo_ctx_comm_size(const vo_ctx* ctx) { return ctx ? ctx->n_ranks : 0; }

// ---- fused exchange over peer memory (CUDA IPC handles; the caller moves the 64-byte handles around)
int vo_ctx_peer_export(vo_ctx* ctx, uint8_t handle[VO_IPC_HANDLE_BYTES]) {
  if (!ctx || !handle) return VO_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == VO_IPC_HANDLE_BYTES, "CUDA IPC handle size");
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  if (!ctx->mailbox) {
    VO_CUDA(ctx, cudaMalloc(&ctx->mailbox, sizeof(VoMailbox)));
    VO_CUDA(ctx, cudaMemset(ctx->mailbox, 0, sizeof(VoMailbox)));
  }
  cudaIpcMemHandle_t h;
  VO_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->mailbox));
  memcpy(handle, &h, VO_IPC_HANDLE_BYTES);
  return VO_OK;
}

int vo_ctx_peer_attach(vo_ctx* ctx, int n_ranks, int rank, const uint8_t* handles) {
  if (!ctx || !handles || n_ranks < 1 || n_ranks > VO_MAX_PEERS || rank < 0 || rank >= n_ranks) return VO_ERR_INVALID;
  if (!ctx->mailbox) return vo_set_error(ctx, VO_ERR_STATE, "vo_ctx_peer_attach", "export first");
  if (ctx->peer_n) return vo_set_error(ctx, VO_ERR_STATE, "vo_ctx_peer_attach", "already attached");
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  for (int r = 0; r < n_ranks; ++r) {
    if (r == rank) {
      ctx->peer_mailbox[r] = ctx->mailbox;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * VO_IPC_HANDLE_BYTES, VO_IPC_HANDLE_BYTES);
    cudaError_t e = cudaIpcOpenMemHandle(&ctx->peer_mailbox[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int q = 0; q < r; ++q)
        if (q != rank && ctx->peer_mailbox[q]) cudaIpcCloseMemHandle(ctx->peer_mailbox[q]);
      for (int q = 0; q < VO_MAX_PEERS; ++q) ctx->peer_mailbox[q] = nullptr;
      cudaGetLastError();
      return vo_set_error(ctx, VO_ERR_CUDA, "cudaIpcOpenMemHandle", cudaGetErrorString(e));
    }
  }
  ctx->peer_n = n_ranks;
  ctx->peer_rank = rank;
  return VO_OK;
}

int vo_ctx_peer_detach(vo_ctx* ctx) {
  if (!ctx) return VO_ERR_INVALID;
  if (ctx->peer_n) {
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < ctx->peer_n; ++r)
      if (r != ctx->peer_rank && ctx->peer_mailbox[r]) cudaIpcCloseMemHandle(ctx->peer_mailbox[r]);
  }
  for (int q = 0; q < VO_MAX_PEERS; ++q) ctx->peer_mailbox[q] = nullptr;
  ctx->peer_n = 0;
  return VO_OK;
}

int vo_ctx_peer_active(const vo_ctx* ctx) { return (ctx && ctx->peer_n > 1) ? 1 : 0; }

}  // extern "C"

int vo_comm_allreduce_f64(vo_ctx* ctx, double* d_buf, int n) {
  if (!ctx->nccl_comm) return VO_OK;
  int r = g_nccl.all_reduce(d_buf, d_buf, (size_t)n, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream);
  if (r != 0) return vo_set_error(ctx, VO_ERR_NCCL, "ncclAllReduce", nccl_err(r));
  return VO_OK;
}

// element-wise MAX of n int32 in place over the communicator (the sharded matcher's exchange); no-op without one
int vo_comm_allreduce_max_i32(vo_ctx* ctx, int32_t* d_buf, long long n) {
  if (!ctx->nccl_comm) return VO_OK;
  const int kNcclInt32 = 2, kNcclMax = 2;  // ncclInt32, ncclMax
  int r = g_nccl.all_reduce(d_buf, d_buf, (size_t)n, kNcclInt32, kNcclMax, ctx->nccl_comm, ctx->stream);
  if (r != 0) return vo_set_error(ctx, VO_ERR_NCCL, "ncclAllReduce", nccl_err(r));
  return VO_OK;
}
