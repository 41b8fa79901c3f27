// vo_common.cuh — context, error plumbing and small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <new>
#include <string.h>

#include "../../include/vo_b200.h"

struct vo_nccl;  // resolved at run time, see comm.cu

struct vo_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int64_t launches = 0;
  char err[256] = {0};
  // scratch arena (device) reused by the stateless entry points; grows monotonically
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  // device staging of the host-buffer entry points (vo_match): a second arena, because the *_dev call underneath
  // carves its own scratch while the staged inputs are still in use
  void* stage = nullptr;
  size_t stage_bytes = 0;
  // pinned host staging for small synchronous read-backs
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  // communicator (nullptr on a single GPU)
  void* nccl_comm = nullptr;
  int n_ranks = 1;
  int rank = 0;
  // fused peer exchange: my mailbox (device memory, IPC-exported) and the peers' mapped mailboxes
  void* mailbox = nullptr;
  void* peer_mailbox[VO_MAX_PEERS] = {nullptr};
  int peer_n = 0;     // 0: not attached
  int peer_rank = 0;
  int match_path = 0;  // VO_MATCH_PATH_*: which of the (all bit-exact) matcher paths vo_match takes
};

// mailbox layout: [2 parities][VO_MAX_PEERS ranks][32 terms] x two 8-byte words {32 payload bits | sequence number << 32}
// (a payload and its flag travel in ONE 8-byte store, which NVLink delivers atomically: no fence, no separate
// flag write - the "LL" protocol), + the local sequence counter
struct VoMailbox {
  unsigned long long ll[2][VO_MAX_PEERS][32][2];
  unsigned seq;       // rounds exchanged so far (advanced by the kernel)
  unsigned timeout;   // set by a kernel whose wait for a peer expired
};

int vo_set_error(vo_ctx* ctx, int status, const char* what, const char* detail);

#define VO_CUDA(ctx, call)                                                              \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) return vo_set_error((ctx), VO_ERR_CUDA, #call, cudaGetErrorString(e__)); \
  } while (0)

#define VO_CHECK_LAUNCH(ctx, name)                                                      \
  do {                                                                                  \
    (ctx)->launches++;                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) return vo_set_error((ctx), VO_ERR_CUDA, name, cudaGetErrorString(e__)); \
  } while (0)

#define VO_REQUIRE(ctx, cond, msg)                                         \
  do {                                                                     \
    if (!(cond)) return vo_set_error((ctx), VO_ERR_INVALID, msg, #cond);   \
  } while (0)

// device scratch of at least `bytes` (256-B aligned), valid until the next vo_scratch call
int vo_scratch(vo_ctx* ctx, size_t bytes, void** out);
int vo_stage(vo_ctx* ctx, size_t bytes, void** out);
int vo_pinned(vo_ctx* ctx, size_t bytes, void** out);
int vo_ctx_activate(vo_ctx* ctx);

// all-reduce (sum) of n doubles in place on the context stream; no-op without a communicator
int vo_comm_allreduce_f64(vo_ctx* ctx, double* d_buf, int n);
int vo_comm_allreduce_max_i32(vo_ctx* ctx, int32_t* d_buf, long long n);

// in-place exclusive scan of per-block counts by one CTA (fixed order); *d_total = grand total
int vo_scan_block_counts(vo_ctx* ctx, int* d_counts, long long n_blocks, long long* d_total);

static inline size_t vo_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

#ifdef __CUDACC__
// streaming 128-bit load: read-only path, do not allocate in L1 (data is touched once per launch)
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// 32 values per lane -> lane L returns the warp total of value L.  Recursive halving: 31 shuffles instead of
// 32 x 5; every total is built over the same butterfly tree (partners L^16, L^8, ..., L^1) as warp_sum, so the
// two give bit-identical sums.
__device__ __forceinline__ float warp_sum32_scatter(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = upper ? v[i] : v[i + o];
      const float keep = upper ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif
