// exp/stream_bench.cu — read-bandwidth microbenchmark (not part of the product): how fast can 5 SoA
// planes of 10.5M floats (210 MB) be streamed with (a) LDG.128 grid-stride, (b) cp.async.bulk ring?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bench stream_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)

__device__ __forceinline__ float4 ldg4(const float4* p){ float4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];":"=f"(r.x),"=f"(r.y),"=f"(r.z),"=f"(r.w):"l"(p)); return r; }

template<int WORK>
__device__ __forceinline__ float burn(float4 a, float4 b, float4 c, float4 d, float4 e){
  float s = a.x+a.y+a.z+a.w+b.x+b.y+b.z+b.w+c.x+c.y+c.z+c.w+d.x+d.y+d.z+d.w+e.x+e.y+e.z+e.w;
#pragma unroll
  for(int i=0;i<WORK;++i) s = fmaf(s, 1.0001f, a.x);
  return s;
}

template<int WORK, int UNROLL>
__global__ void __launch_bounds__(256) k_ldg(const float* pk, long long stride, long long n, float* out){
  const float4* p0=(const float4*)pk; const float4* p1=(const float4*)(pk+stride); const float4* p2=(const float4*)(pk+2*stride);
  const float4* p3=(const float4*)(pk+3*stride); const float4* p4=(const float4*)(pk+4*stride);
  long long nq=n/4, step=(long long)gridDim.x*256; float s=0;
  for(long long q=(long long)blockIdx.x*256+threadIdx.x; q<nq; q+=step*UNROLL){
    float4 a[UNROLL],b[UNROLL],c[UNROLL],d[UNROLL],e[UNROLL];
#pragma unroll
    for(int u=0;u<UNROLL;++u){ long long qq=q+u*step; if(qq<nq){ a[u]=ldg4(p0+qq); b[u]=ldg4(p1+qq); c[u]=ldg4(p2+qq); d[u]=ldg4(p3+qq); e[u]=ldg4(p4+qq);} }
#pragma unroll
    for(int u=0;u<UNROLL;++u){ long long qq=q+u*step; if(qq<nq) s+=burn<WORK>(a[u],b[u],c[u],d[u],e[u]); }
  }
  if(s==123.456f) out[0]=s;
}

__device__ __forceinline__ unsigned su32(const void* p){ return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(unsigned long long* b, unsigned c){ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"::"r"(su32(b)),"r"(c)); }
__device__ __forceinline__ void mb_arrive(unsigned long long* b){ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];"::"r"(su32(b)):"memory"); }
__device__ __forceinline__ void mb_expect(unsigned long long* b, unsigned bytes){ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(su32(b)),"r"(bytes):"memory"); }
__device__ __forceinline__ void mb_wait(unsigned long long* b, unsigned parity){
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"::"r"(su32(b)),"r"(parity):"memory"); }
__device__ __forceinline__ void bulk(void* dst, const void* src, unsigned bytes, unsigned long long* b){
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"::"r"(su32(dst)),"l"(src),"r"(bytes),"r"(su32(b)):"memory"); }

template<int WORK, int STAGES, int TILE>   // TILE floats per plane per stage, 256 consumer threads (+32 producer)
__global__ void __launch_bounds__(288) k_tma(const float* pk, long long stride, long long n, float* out){
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) unsigned long long full[STAGES], empty[STAGES];
  int warp=threadIdx.x>>5, lane=threadIdx.x&31;
  if(threadIdx.x==0){ for(int s=0;s<STAGES;++s){ mb_init(&full[s],1); mb_init(&empty[s],8);} asm volatile("fence.mbarrier_init.release.cluster;":::"memory"); }
  __syncthreads();
  long long nt=n/TILE; float s=0;
  if(warp==8){ if(lane==0){ int it=0; for(long long t=blockIdx.x;t<nt;t+=gridDim.x,++it){ int st=it%STAGES; unsigned ph=(it/STAGES)&1; mb_wait(&empty[st],ph^1);
        float* dst=sm+(size_t)st*5*TILE; mb_expect(&full[st],5*TILE*4);
        for(int p=0;p<5;++p) bulk(dst+p*TILE, pk+p*stride+t*TILE, TILE*4, &full[st]); } } }
  else { int it=0; for(long long t=blockIdx.x;t<nt;t+=gridDim.x,++it){ int st=it%STAGES; unsigned ph=(it/STAGES)&1; mb_wait(&full[st],ph);
        const float4* tl=(const float4*)(sm+(size_t)st*5*TILE);
        for(int q=threadIdx.x;q<TILE/4;q+=256){ float4 a=tl[q],b=tl[TILE/4+q],c=tl[2*TILE/4+q],d=tl[3*TILE/4+q],e=tl[4*TILE/4+q]; s+=burn<WORK>(a,b,c,d,e);} 
        __syncwarp(); if(lane==0) mb_arrive(&empty[st]); } }
  if(s==123.456f) out[0]=s;
}

template<class F> float timeit(F f, int reps){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); f(); CK(cudaDeviceSynchronize()); cudaEventRecord(a); for(int i=0;i<reps;++i) f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms,a,b); return ms/reps*1e3f; }

int main(){
  long long n=10485760; float *pk,*out; CK(cudaMalloc(&pk,n*5*4)); CK(cudaMalloc(&out,4)); CK(cudaMemset(pk,0,n*5*4));
  double mb=n*20.0/1e6; int sms; cudaDeviceGetAttribute(&sms,cudaDevAttrMultiProcessorCount,0);
  printf("SMs %d, %.1f MB per pass\n",sms,mb);
#define RUN_LDG(W,U,CPS) { float us=timeit([&]{ k_ldg<W,U><<<sms*CPS,256>>>(pk,n,n,out); },20); printf("ldg  work %3d unroll %d ctas/sm %d : %7.1f us  %6.2f TB/s\n",W,U,CPS,us,mb/us/1e0*1e-6*1e6/1e6); }
  RUN_LDG(0,1,2) RUN_LDG(0,2,2) RUN_LDG(0,4,2) RUN_LDG(0,1,4) RUN_LDG(0,2,4) RUN_LDG(0,1,8) RUN_LDG(0,2,8)
  RUN_LDG(100,1,2) RUN_LDG(100,2,2) RUN_LDG(100,2,4) RUN_LDG(400,2,2) RUN_LDG(400,2,4)
#define RUN_TMA(W,S,T,CPS) { size_t smb=(size_t)S*5*T*4; CK(cudaFuncSetAttribute(k_tma<W,S,T>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smb)); float us=timeit([&]{ k_tma<W,S,T><<<sms*CPS,288,smb>>>(pk,n,n,out); },20); CK(cudaGetLastError()); printf("tma  work %3d stages %d tile %5d ctas/sm %d smem %6zu : %7.1f us  %6.2f TB/s\n",W,S,T,CPS,smb,us,mb/us); }
  RUN_TMA(0,4,1024,2) RUN_TMA(0,2,1024,2) RUN_TMA(0,4,2048,1) RUN_TMA(0,4,2048,2) RUN_TMA(0,8,1024,1) RUN_TMA(0,4,512,4) RUN_TMA(0,3,1024,3)
  RUN_TMA(100,4,1024,2) RUN_TMA(400,4,1024,2) RUN_TMA(400,4,2048,1) RUN_TMA(400,3,1024,3)
  return 0;
}
