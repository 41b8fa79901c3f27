// exp/gram_bench.cu — the PICP H / b accumulation as a tensor-core Gram [J|e]^T [J|e] (mma.sync m16n8k8 TF32) against
// the packed-FFMA2 form the product uses, in isolation on sm_100a (not part of the product).
//
// Every lane owns correspondences whose augmented Jacobian rows r0 = [J0 | e0 | 0], r1 = [J1 | e1 | 0] (8 wide) are
// already in registers (a cheap per-iteration update keeps the compiler from folding them).  Four modes:
//   0  update only (the cost both forms share)
//   1  SIMT: 40 MACs per correspondence (21 H + 6 b slots, the two structural zeros of the pinhole Jacobian skipped),
//      two correspondences per lane packed in f32x2 = 40 FFMA2 per pair per lane, as csrc/picp.cu pair_back does
//   2  tensor core, fp32-grade: rows through shared memory (the fragment of lane (g,t) is feature g of rows t, t+4, i.e.
//      of ONE correspondence's two rows: a transpose), split x = hi + lo, three MMAs per 4 correspondences
//      (hi*hi + hi*lo + lo*hi)
//   3  tensor core, one TF32 MMA per 4 correspondences (operands truncated to 10 mantissa bits)
// Prints ns per 32 correspondences per warp-step and the largest relative deviation of the 8x8 Gram from a float64 sum.
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float a, float b) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(f2 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2 ffma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

__device__ __forceinline__ void mma_tf32(float (&d)[4], unsigned a0, unsigned a2, unsigned b0, unsigned b1) {
  // rows 8..15 of the 16 x 8 A tile are the zero padding of the 8-row Gram: a1 = a3 = 0
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}

constexpr int kThreads = 384, kWarps = kThreads / 32, kStride = 24;  // 24-float rows: conflict-free fragment reads

// rows of lane `lane`, correspondence c (0/1) at iteration it: deterministic, cheap, O(1) magnitudes
__device__ __forceinline__ void make_rows(int gtid, int it, float (&r0)[2][8], float (&r1)[2][8]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const float s = 1.f + 0.001f * (float)((gtid * 2 + c) & 1023), u = 0.5f + 0.0007f * (float)(it & 1023);
#pragma unroll
    for (int f = 0; f < 7; ++f) {
      r0[c][f] = fmaf(s, 0.3f + 0.1f * f, -u);
      r1[c][f] = fmaf(u, 0.2f + 0.05f * f, s * 0.25f);
    }
    r0[c][1] = 0.f; r1[c][0] = 0.f;  // pinhole structural zeros
    r0[c][7] = 0.f; r1[c][7] = 0.f;
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) gram_kernel(float* out /* [grid][warps][64] */, int iters) {
  extern __shared__ __align__(16) float s_dyn[];  // [warp][correspondence][J0|e0|0 at +0, J1|e1|0 at +8, pad to 24]
  float (*s_rows)[64][kStride] = reinterpret_cast<float (*)[64][kStride]>(s_dyn);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int gtid = blockIdx.x * kThreads + tid;
  f2 acc[36];  // SIMT: upper triangle of the 8x8 Gram, two correspondences packed
#pragma unroll
  for (int i = 0; i < 36; ++i) acc[i] = 0ull;
  float d[4] = {0.f, 0.f, 0.f, 0.f};  // MMA accumulators: lane (g,t) holds G[g][2t], G[g][2t+1] (+ the padding rows)
  float sink = 0.f;
  for (int it = 0; it < iters; ++it) {
    float r0[2][8], r1[2][8];
    make_rows(gtid, it, r0, r1);
    if (MODE == 0) {
#pragma unroll
      for (int f = 0; f < 7; ++f) sink += r0[0][f] + r1[0][f] + r0[1][f] + r1[1][f];
    }
    if (MODE == 1) {
      f2 p0[7], p1[7];
#pragma unroll
      for (int f = 0; f < 7; ++f) { p0[f] = pk(r0[0][f], r0[1][f]); p1[f] = pk(r1[0][f], r1[1][f]); }
      int k = 0;
#pragma unroll
      for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int j = i; j < 7; ++j, ++k) {
          if (i == 6 && j == 6) continue;                                   // chi is summed elsewhere in the product
          if (!(i == 1 || j == 1)) acc[k] = ffma2(p0[i], p0[j], acc[k]);   // J0[1] = 0
          if (!(i == 0 || j == 0)) acc[k] = ffma2(p1[i], p1[j], acc[k]);   // J1[0] = 0
        }
    }
    if (MODE >= 2) {
      // transpose through shared memory: the owner writes its two correspondences' rows, lane (g,t) reads feature g of
      // correspondence 4s+t (rows t and t+4 of step s)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float* dst = s_rows[warp][2 * lane + c];
        *reinterpret_cast<float4*>(dst) = make_float4(r0[c][0], r0[c][1], r0[c][2], r0[c][3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(r0[c][4], r0[c][5], r0[c][6], r0[c][7]);
        *reinterpret_cast<float4*>(dst + 8) = make_float4(r1[c][0], r1[c][1], r1[c][2], r1[c][3]);
        *reinterpret_cast<float4*>(dst + 12) = make_float4(r1[c][4], r1[c][5], r1[c][6], r1[c][7]);
      }
      __syncwarp();
#pragma unroll
      for (int s = 0; s < 16; ++s) {  // 64 correspondences per warp-iteration, 4 per MMA step
        const float* src = s_rows[warp][4 * s + t];
        const float x0 = src[g], x1 = src[8 + g];
        if (MODE == 3) {
          mma_tf32(d, __float_as_uint(x0), __float_as_uint(x1), __float_as_uint(x0), __float_as_uint(x1));
        } else {
          const unsigned h0 = __float_as_uint(x0) & 0xffffe000u, h1 = __float_as_uint(x1) & 0xffffe000u;
          const unsigned l0 = __float_as_uint(x0 - __uint_as_float(h0)), l1 = __float_as_uint(x1 - __uint_as_float(h1));
          mma_tf32(d, h0, h1, h0, h1);
          mma_tf32(d, h0, h1, l0, l1);
          mma_tf32(d, l0, l1, h0, h1);
        }
      }
      __syncwarp();
    }
  }
  // the warp's 8x8 Gram (upper triangle meaningful) to out
  float* o = out + ((size_t)blockIdx.x * kWarps + warp) * 64;
  if (MODE == 1) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
      for (int j = i; j < 7; ++j, ++k) {
        float a, b;
        upk(acc[k], a, b);
        float v = a + b;
        for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) o[i * 8 + j] = v;
      }
  } else if (MODE >= 2) {
    o[g * 8 + 2 * t] = d[0];
    o[g * 8 + 2 * t + 1] = d[1];
  } else if (sink == 1.2345f) {
    o[0] = sink;
  }
}

static void reference(int grid, int iters, int block, int warp, double (&G)[8][8]) {
  for (auto& row : G) for (auto& v : row) v = 0;
  for (int lane = 0; lane < 32; ++lane) {
    const int gtid = block * kThreads + warp * 32 + lane;
    for (int it = 0; it < iters; ++it)
      for (int c = 0; c < 2; ++c) {
        const float s = 1.f + 0.001f * (float)((gtid * 2 + c) & 1023), u = 0.5f + 0.0007f * (float)(it & 1023);
        float r0[8], r1[8];
        for (int f = 0; f < 7; ++f) {
          r0[f] = fmaf(s, 0.3f + 0.1f * f, -u);
          r1[f] = fmaf(u, 0.2f + 0.05f * f, s * 0.25f);
        }
        r0[1] = 0; r1[0] = 0; r0[7] = 0; r1[7] = 0;
        for (int i = 0; i < 8; ++i)
          for (int j = 0; j < 8; ++j) G[i][j] += (double)r0[i] * r0[j] + (double)r1[i] * r1[j];
      }
  }
  (void)grid;
}

template <int MODE>
static double run(const char* name, float* d_out, int grid, int iters, double base_ms) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaMemset(d_out, 0, (size_t)grid * kWarps * 64 * sizeof(float));
  const size_t smem = (size_t)kWarps * 64 * kStride * sizeof(float);
  cudaFuncSetAttribute(gram_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  gram_kernel<MODE><<<grid, kThreads, smem>>>(d_out, iters);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  gram_kernel<MODE><<<grid, kThreads, smem>>>(d_out, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  double worst = 0;
  if (MODE >= 1) {
    std::vector<float> h((size_t)grid * kWarps * 64);
    cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    double G[8][8];
    reference(grid, iters, grid - 1, kWarps - 1, G);
    const float* o = h.data() + ((size_t)(grid - 1) * kWarps + (kWarps - 1)) * 64;
    for (int i = 0; i < 7; ++i)
      for (int j = i; j < 7; ++j) {
        if (i == 6 && j == 6 && MODE == 1) continue;
        if (G[i][j] == 0) continue;
        worst = fmax(worst, fabs(o[i * 8 + j] - G[i][j]) / fabs(G[i][j]));
      }
  }
  // one warp-iteration = 64 correspondences; per SM kWarps warps share 4 schedulers
  const double ns_per_32 = (ms * 1e6) / iters / 2.0;
  printf("%-34s %8.3f ms   %7.2f ns per 32 correspondences of a warp   (+%6.2f over the shared update)   max rel dev %.2e\n", name, ms,
         ns_per_32, (ms - base_ms) * 1e6 / iters / 2.0, worst);
  return ms;
}

int main() {
  const int grid = 148, iters = 2048;
  float* d_out;
  cudaMalloc(&d_out, (size_t)grid * kWarps * 64 * sizeof(float));
  const double base = run<0>("update only", d_out, grid, iters, 0.0);
  run<1>("SIMT FFMA2 (product form)", d_out, grid, iters, base);
  run<2>("mma.sync TF32 x3 (hi/lo split)", d_out, grid, iters, base);
  run<3>("mma.sync TF32 x1 (truncated)", d_out, grid, iters, base);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
