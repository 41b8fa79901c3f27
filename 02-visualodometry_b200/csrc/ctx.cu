// ctx.cu — context lifecycle, scratch arenas, error strings, host-side isometry helpers.
#include "vo_common.cuh"

#include <stdlib.h>

int vo_set_error(vo_ctx* ctx, int status, const char* what, const char* detail) {
  if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what ? what : "", detail ? detail : "");
  return status;
}

extern "C" {

const char* vo_status_str(int status) {
  switch (status) {
    case VO_OK: return "ok";
    case VO_ERR_INVALID: return "invalid argument";
    case VO_ERR_CUDA: return "CUDA error";
    case VO_ERR_NCCL: return "NCCL error";
    case VO_ERR_NOMEM: return "out of memory";
    case VO_ERR_STATE: return "invalid call order";
    case VO_ERR_CAPACITY: return "output buffer too small";
    default: return "unknown status";
  }
}

int vo_version(void) { return VO_B200_VERSION; }

int vo_device_count(int* n) {
  if (!n) return VO_ERR_INVALID;
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  *n = (e == cudaSuccess) ? c : 0;
  return e == cudaSuccess ? VO_OK : VO_ERR_CUDA;
}

int vo_ctx_create(int device, void* cuda_stream, vo_ctx** out) {
  if (!out) return VO_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return VO_ERR_CUDA;  // no CPU fallback
  if (device < 0 || device >= count) return VO_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return VO_ERR_CUDA;
  vo_ctx* c = new (std::nothrow) vo_ctx();
  if (!c) return VO_ERR_NOMEM;
  c->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete c;
    return VO_ERR_CUDA;
  }
  c->sm_count = prop.multiProcessorCount;
  if (cuda_stream) {
    c->stream = (cudaStream_t)cuda_stream;
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete c;
      return VO_ERR_CUDA;
    }
    c->own_stream = true;
  }
  *out = c;
  return VO_OK;
}

int vo_ctx_destroy(vo_ctx* ctx) {
  if (!ctx) return VO_OK;
  cudaSetDevice(ctx->device);
  vo_ctx_comm_destroy(ctx);
  vo_ctx_peer_detach(ctx);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->mailbox) cudaFree(ctx->mailbox);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->stage) cudaFree(ctx->stage);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return VO_OK;
}

int vo_ctx_sync(vo_ctx* ctx) {
  if (!ctx) return VO_ERR_INVALID;
  VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VO_OK;
}

const char* vo_last_error(const vo_ctx* ctx) { return ctx ? ctx->err : "null context"; }
int64_t vo_ctx_kernel_launches(const vo_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* vo_ctx_stream(const vo_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// Eigen Isometry3f::inverse(): R^T and -(R^T) t, coefficient products reduced as x0+(x1+x2)
// (exec/icp_test.cpp:79,114). Host arithmetic; this TU is compiled with -ffp-contract=off.
void vo_pose_inverse(const float T[12], float out[12]) {
  float r[12];
  for (int i = 0; i < 3; ++i) {
    float a = -T[i], b = -T[4 + i], c = -T[8 + i];
    r[4 * i + 0] = T[i];
    r[4 * i + 1] = T[4 + i];
    r[4 * i + 2] = T[8 + i];
    float x0 = a * T[3], x1 = b * T[7], x2 = c * T[11];
    r[4 * i + 3] = x0 + (x1 + x2);
  }
  memcpy(out, r, sizeof(r));
}

// Isometry3f * Isometry3f: R = Ra Rb, t = Ra tb + ta
void vo_pose_mul(const float A[12], const float B[12], float out[12]) {
  float r[12];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 4; ++j) {
      float x0 = A[4 * i] * B[j], x1 = A[4 * i + 1] * B[4 + j], x2 = A[4 * i + 2] * B[8 + j];
      r[4 * i + j] = x0 + (x1 + x2);
    }
    r[4 * i + 3] = r[4 * i + 3] + A[4 * i + 3];
  }
  memcpy(out, r, sizeof(r));
}

}  // extern "C"

int vo_ctx_activate(vo_ctx* ctx) {
  if (!ctx) return VO_ERR_INVALID;
  VO_CUDA(ctx, cudaSetDevice(ctx->device));
  return VO_OK;
}

int vo_scratch(vo_ctx* ctx, size_t bytes, void** out) {
  if (bytes > ctx->scratch_bytes) {
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = vo_align_up(bytes + bytes / 4, 1 << 20);
    cudaError_t e = cudaMalloc(&ctx->scratch, want);
    if (e != cudaSuccess) return vo_set_error(ctx, VO_ERR_NOMEM, "cudaMalloc(scratch)", cudaGetErrorString(e));
    ctx->scratch_bytes = want;
  }
  *out = ctx->scratch;
  return VO_OK;
}

int vo_stage(vo_ctx* ctx, size_t bytes, void** out) {
  if (bytes > ctx->stage_bytes) {
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->stage) cudaFree(ctx->stage);
    ctx->stage = nullptr;
    ctx->stage_bytes = 0;
    size_t want = vo_align_up(bytes + bytes / 4, 1 << 20);
    cudaError_t e = cudaMalloc(&ctx->stage, want);
    if (e != cudaSuccess) return vo_set_error(ctx, VO_ERR_NOMEM, "cudaMalloc(stage)", cudaGetErrorString(e));
    ctx->stage_bytes = want;
  }
  *out = ctx->stage;
  return VO_OK;
}

int vo_pinned(vo_ctx* ctx, size_t bytes, void** out) {
  if (bytes > ctx->pinned_bytes) {
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
    size_t want = vo_align_up(bytes, 1 << 16);
    cudaError_t e = cudaMallocHost(&ctx->pinned, want);
    if (e != cudaSuccess) return vo_set_error(ctx, VO_ERR_NOMEM, "cudaMallocHost", cudaGetErrorString(e));
    ctx->pinned_bytes = want;
  }
  *out = ctx->pinned;
  return VO_OK;
}
