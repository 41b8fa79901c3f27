// exp/fma2_bench.cu — issue-rate microbenchmark for packed f32x2 ops on sm_100a (not part of the product)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b){ unsigned long long r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c){ unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d; }
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b){ unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d; }
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b){ unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d; }

template<int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float x){
  float a[16]; unsigned long long p[8];
  for(int i=0;i<16;++i) a[i]=x+i+threadIdx.x;
  for(int i=0;i<8;++i) p[i]=pk(a[2*i],a[2*i+1]);
  unsigned long long m=pk(x,x*0.5f), c=pk(0.25f,0.125f);
  for(int it=0;it<iters;++it){
    if(MODE==0){ _Pragma("unroll") for(int i=0;i<16;++i) a[i]=fmaf(a[i],x,0.25f); }
    if(MODE==1){ _Pragma("unroll") for(int i=0;i<8;++i) p[i]=ffma2(p[i],m,c); }
    if(MODE==2){ _Pragma("unroll") for(int i=0;i<16;++i) a[i]=__fmul_rn(a[i],x); }
    if(MODE==3){ _Pragma("unroll") for(int i=0;i<8;++i) p[i]=fmul2(p[i],m); }
    if(MODE==4){ _Pragma("unroll") for(int i=0;i<16;++i) a[i]=__fadd_rn(a[i],x); }
    if(MODE==5){ _Pragma("unroll") for(int i=0;i<8;++i) p[i]=fadd2(p[i],m); }
  }
  float s=0; for(int i=0;i<16;++i) s+=a[i]; for(int i=0;i<8;++i){ float u,v; upk(p[i],u,v); s+=u+v; }
  if(s==1.2345f) out[0]=s;
}
template<int MODE> void run(const char* name, float* out){
  int iters=4096; cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148*8,256>>>(out,iters,1.0001f); cudaDeviceSynchronize();
  cudaEventRecord(a); k<MODE><<<148*8,256>>>(out,iters,1.0001f); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b);
  double lane_ops = 148.0*8*256*(double)iters*16;   // scalar-equivalent float ops
  printf("%-8s %8.3f ms  %7.2f T scalar-op/s  (%.1f per clk per SM at 1.965 GHz)\n", name, ms, lane_ops/ms/1e9, lane_ops/(ms*1e-3)/148/1.965e9);
}
int main(){ float* out; cudaMalloc(&out,4); run<0>("FFMA",out); run<1>("FFMA2",out); run<2>("FMUL",out); run<3>("FMUL2",out); run<4>("FADD",out); run<5>("FADD2",out); printf("%s\n", cudaGetErrorString(cudaGetLastError())); return 0; }
