// exp/mbar_pending_test.cu — what does mbarrier.pending_count report for the state returned by mbarrier.arrive?
// nvcc -gencode arch=compute_100a,code=sm_100a -o exp/mbar_pending_test exp/mbar_pending_test.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned* out) {
  __shared__ __align__(8) unsigned long long bar;
  const unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(3u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < 7; ++i) {
      unsigned long long st;
      unsigned cnt;
      asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(b) : "memory");
      asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(cnt) : "l"(st));
      out[i] = cnt;
    }
  }
}
int main() {
  unsigned* d; cudaMalloc(&d, 64); k<<<1, 32>>>(d); unsigned h[7]; cudaMemcpy(h, d, 28, cudaMemcpyDeviceToHost);
  printf("init count 3; pending_count of the state returned by 7 successive arrives:");
  for (int i = 0; i < 7; ++i) printf(" %u", h[i]);
  printf("\n%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
