// exp/hacc_bench.cu — the H/b accumulation pattern of the PICP kernel in isolation (not part of the product):
// 12 packed Jacobian values + 2 packed errors -> 27 packed accumulators, the pinhole structural zeros skipped,
// once with FFMA2 (as the kernel does) and once with two scalar FFMA per packed op.  Reports cycles per
// correspondence pair at 1..4 warps per scheduler.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float a, float b){ f2 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r; }
__device__ __forceinline__ void upk(f2 v, float& a, float& b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ f2 ffma2(f2 a, f2 b, f2 c){ f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d; }
__device__ __forceinline__ f2 fmul2(f2 a, f2 b){ f2 d; asm("mul.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d; }

template <int MODE> __global__ void __launch_bounds__(512) k(float* out, int iters, float x) {
  f2 acc[27];
  float sa[27], sb[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) { acc[i] = 0ull; sa[i] = 0.f; sb[i] = 0.f; }
  float seed = x + threadIdx.x * 1e-3f;
  for (int it = 0; it < iters; ++it) {
    // stand-in for J: cheap dependent values (a few packed ops), different every iteration
    f2 J0[6], J1[6], e0, e1;
    const f2 s = pk(seed, seed * 0.5f);
#pragma unroll
    for (int i = 0; i < 6; ++i) { J0[i] = fmul2(s, pk(1.f + i, 2.f + i)); J1[i] = fmul2(s, pk(3.f + i, 0.5f + i)); }
    e0 = fmul2(s, pk(0.1f, 0.2f)); e1 = fmul2(s, pk(0.3f, 0.4f));
    J0[1] = 0ull; J1[0] = 0ull;
    seed = seed * 1.0000001f + 1e-7f;
    int kk = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const bool z0 = i == 1, z1 = i == 0;
#pragma unroll
      for (int j = i; j < 6; ++j, ++kk) {
        const bool y0 = z0 || j == 1, y1 = z1 || j == 0;
        if (MODE == 0) {
          if (!y0) acc[kk] = ffma2(J0[i], J0[j], acc[kk]);
          if (!y1) acc[kk] = ffma2(J1[i], J1[j], acc[kk]);
        } else {
          float a0, a1, b0, b1;
          if (!y0) { upk(J0[i], a0, a1); upk(J0[j], b0, b1); sa[kk] = fmaf(a0, b0, sa[kk]); sb[kk] = fmaf(a1, b1, sb[kk]); }
          if (!y1) { upk(J1[i], a0, a1); upk(J1[j], b0, b1); sa[kk] = fmaf(a0, b0, sa[kk]); sb[kk] = fmaf(a1, b1, sb[kk]); }
        }
      }
      if (MODE == 0) {
        if (!z0) acc[21 + i] = ffma2(J0[i], e0, acc[21 + i]);
        if (!z1) acc[21 + i] = ffma2(J1[i], e1, acc[21 + i]);
      } else {
        float a0, a1, b0, b1;
        if (!z0) { upk(J0[i], a0, a1); upk(e0, b0, b1); sa[21 + i] = fmaf(a0, b0, sa[21 + i]); sb[21 + i] = fmaf(a1, b1, sb[21 + i]); }
        if (!z1) { upk(J1[i], a0, a1); upk(e1, b0, b1); sa[21 + i] = fmaf(a0, b0, sa[21 + i]); sb[21 + i] = fmaf(a1, b1, sb[21 + i]); }
      }
    }
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < 27; ++i) { float u, v; upk(acc[i], u, v); r += u + v + sa[i] + sb[i]; }
  if (r == 1.2345f) out[0] = r;
}
template <int MODE> void run(const char* name, float* out, int threads) {
  int iters = 20000; cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148, threads>>>(out, iters, 1.0001f); cudaDeviceSynchronize();
  cudaEventRecord(a); k<MODE><<<148, threads>>>(out, iters, 1.0001f); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double cyc = ms * 1e-3 * 1.965e9;
  const double per_pair_per_smsp = cyc / ((double)iters * (threads / 32) / 4.0);  // scheduler cycles per warp-level pair
  printf("%-6s %3d threads (%.2f warps/scheduler): %8.3f ms, %6.1f scheduler cycles per warp-pair (40 packed FMA + 14 packed MUL)\n",
         name, threads, threads / 128.0, ms, per_pair_per_smsp);
}
int main() {
  float* out; cudaMalloc(&out, 4);
  for (int t : {128, 256, 352, 512}) { run<0>("FFMA2", out, t); run<1>("FFMA", out, t); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
