// exp/pipe_mix_bench.cu — do the FMA pipe (packed FFMA2 / scalar FFMA) and the ALU pipe (FSETP / FSEL) of sm_100a
// overlap?  (not part of the product)   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o exp/pipe_mix_bench exp/pipe_mix_bench.cu
// Each mode runs NF packed FFMA2 or 2*NF scalar FFMA plus NA compare+select pairs per iteration on independent
// registers; reported: scheduler cycles per iteration per warp at 1..4 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float a, float b){ f2 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r; }
__device__ __forceinline__ void upk(f2 v, float& a, float& b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ f2 ffma2(f2 a, f2 b, f2 c){ f2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d; }
__device__ __forceinline__ float ffma1(float a, float b, float c){ float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;":"=f"(d):"f"(a),"f"(b),"f"(c)); return d; }
// one FSETP + one FSEL
__device__ __forceinline__ float cmpsel(float a, float t, float z){ float d; asm volatile("{.reg .pred p; setp.gt.f32 p, %1, %2; selp.f32 %0, %1, %3, p;}":"=f"(d):"f"(a),"f"(t),"f"(z)); return d; }

template <int PACKED, int NF, int NA> __global__ void __launch_bounds__(512) k(float* out, int iters, float x) {
  f2 p[8]; float a[16]; float y[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = pk(x + i, x - i);
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = x + i + threadIdx.x; y[i] = x * i + threadIdx.x; }
  const f2 m = pk(x, x * 0.5f), c = pk(0.25f, 0.125f);
  const float t = x * 3.f, z = x * 0.25f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PACKED == 1 && i < NF) p[i] = ffma2(p[i], m, c);
      if (PACKED == 0 && i < NF) { a[2 * i] = ffma1(a[2 * i], x, z); a[2 * i + 1] = ffma1(a[2 * i + 1], x, z); }
      if (i < NA) y[i] = cmpsel(y[i], t, z);
      if (i + 8 < NA) y[i + 8] = cmpsel(y[i + 8], t, z);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float u, v; upk(p[i], u, v); s += u + v; }
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i] + y[i];
  if (s == 1.2345f) out[0] = s;
}
template <int PACKED, int NF, int NA> void run(const char* name, float* out) {
  printf("%-34s", name);
  for (int threads : {128, 256, 384, 512}) {
    int iters = 1 << 16; cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<PACKED, NF, NA><<<148, threads>>>(out, iters, 1.0001f); cudaDeviceSynchronize();
    cudaEventRecord(a); k<PACKED, NF, NA><<<148, threads>>>(out, iters, 1.0001f); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double cyc = ms * 1e-3 * 1.965e9;
    printf("  %d w/s: %6.1f", threads / 128, cyc / ((double)iters * (threads / 128)));
  }
  printf("   (cycles per iteration per warp)\n");
}
int main() {
  float* out; cudaMalloc(&out, 4);
  run<1, 8, 0>("8 FFMA2", out);
  run<0, 8, 0>("16 FFMA", out);
  run<1, 0, 8>("8 (FSETP+FSEL)", out);
  run<1, 0, 16>("16 (FSETP+FSEL)", out);
  run<1, 8, 8>("8 FFMA2 + 8 (FSETP+FSEL)", out);
  run<1, 8, 16>("8 FFMA2 + 16 (FSETP+FSEL)", out);
  run<0, 8, 8>("16 FFMA + 8 (FSETP+FSEL)", out);
  run<0, 8, 16>("16 FFMA + 16 (FSETP+FSEL)", out);
  run<0, 4, 8>("8 FFMA + 8 (FSETP+FSEL)", out);
  run<1, 4, 8>("4 FFMA2 + 8 (FSETP+FSEL)", out);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
