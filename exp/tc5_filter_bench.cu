// exp/tc5_filter_bench.cu — the matcher's lower-bound filter tile (128 rows x 128 columns x K = 16, bf16 -> fp32,
// epilogue = harvest the sign bit of every accumulator into survivor masks) once with tcgen05.mma + TMEM + tcgen05.ld and
// once with mma.sync.m16n8k16, operands already resident (shared memory / registers), one CTA of 128 threads per SM.
// Not part of the product: it answers "would tcgen05 beat mma.sync for a K = 16 contraction with a compare-only
// epilogue?" with a measurement (profiles/r02_tc5_filter_bench.md).  The operand VALUES are irrelevant here (the
// shared-memory tiles hold a fixed pattern in the canonical no-swizzle K-major layout); only instruction throughput is
// measured.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp/tc5_filter_bench exp/tc5_filter_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait_bounded(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  for (long long spins = 0; !ok; ++spins) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spins > (1ll << 24)) asm volatile("trap;");
  }
}
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(unsigned taddr, unsigned (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ unsigned harvest16(const unsigned (&r)[16], unsigned m) {
#pragma unroll
  for (int i = 0; i < 16; ++i) m = __funnelshift_l(r[i], m, 1);  // m = (m << 1) | sign(r[i]): one SHF per accumulator
  return m;
}

constexpr int kM = 128, kN = 128, kK = 16;

// MODE 0: tcgen05.mma + tcgen05.ld + epilogue   1: tcgen05.mma only (no read-out)   2: tcgen05.mma + tcgen05.ld, no epilogue
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_tc5(unsigned* out, int tiles) {
  __shared__ __align__(1024) __nv_bfloat16 sA[kM * kK], sB[kN * kK];  // 8-row x 16-byte core matrices, K groups adjacent
  constexpr int kStagesT = 4;  // accumulator stages in TMEM (4 x 128 columns = all 512)
  __shared__ __align__(8) unsigned long long s_bar[kStagesT];
  __shared__ unsigned s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kM * kK; i += 128) { sA[i] = __float2bfloat16(0.01f * (i % 37) - 0.2f); sB[i] = __float2bfloat16(0.02f * (i % 29) - 0.3f); }
  if (tid == 0) {
    for (int s = 0; s < kStagesT; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores to sA / sB -> async proxy (the MMA)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const unsigned tmem = s_tmem;
  // shared-memory matrix descriptors: no swizzle, K-major; core matrices 128 B apart along K, 256 B apart along M / N
  auto desc = [](unsigned addr) {
    return (unsigned long long)((addr >> 4) & 0x3FFF) | ((unsigned long long)(128 >> 4) << 16) | ((unsigned long long)(256 >> 4) << 32) |
           (1ull << 46);
  };
  const unsigned long long dA = desc(smem_u32(sA)), dB = desc(smem_u32(sB));
  // instruction descriptor: D = F32, A = B = BF16, both K-major, N = 128, M = 128
  const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(kN >> 3) << 17) | ((unsigned)(kM >> 4) << 24);
  auto issue = [&](int t) {
    const unsigned d = tmem + (unsigned)(t % kStagesT) * kN;
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(dA),
                 "l"(dB), "r"(idesc), "r"(0u)
                 : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&s_bar[t % kStagesT])) : "memory");
  };
  unsigned check = 0;
  if (tid == 0)
    for (int t = 0; t < kStagesT - 1 && t < tiles; ++t) issue(t);
  for (int t = 0; t < tiles; ++t) {
    asm volatile("tcgen05.fence::after_thread_sync;");
    // three tiles ahead: that stage's last readers passed the barrier at the end of the iteration before
    if (tid == 0 && t + kStagesT - 1 < tiles) issue(t + kStagesT - 1);
    mbar_wait_bounded(smem_u32(&s_bar[t % kStagesT]), (unsigned)(t / kStagesT) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (MODE != 1) {
      const unsigned taddr = tmem + (unsigned)(t % kStagesT) * kN + ((unsigned)(32 * warp) << 16);
      unsigned m[4] = {0, 0, 0, 0};
      unsigned r[8][16];
#pragma unroll
      for (int c = 0; c < 8; ++c) tmem_ld16(taddr + 16 * c, r[c]);  // all eight loads in flight, one wait
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (MODE == 0) m[c >> 1] = harvest16(r[c], m[c >> 1]);
        else m[c >> 1] ^= r[c][c];
      }
      check += (m[0] | m[1]) ^ (m[2] + m[3]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  if (check == 0x12345678u) out[0] = check;
}

// the same tile with mma.sync: 4 warps x (32 rows x 128 columns) = 2 x 16 m16n8k16 per warp, fragments in registers
// MODE 0: HMMA + epilogue   1: HMMA only   3: epilogue only (accumulators = a changing register pattern)
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_sync(unsigned* out, int tiles) {
  const int tid = threadIdx.x;
  unsigned a[2][4], b[16][2];
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) a[i][j] = 0x3c003c00u + tid * 7 + i * 3 + j;
  for (int i = 0; i < 16; ++i) for (int j = 0; j < 2; ++j) b[i][j] = 0x3a003a00u + tid * 5 + i * 11 + j;
  unsigned check = 0;
  for (int t = 0; t < tiles; ++t) {
    unsigned m[4] = {0, 0, 0, 0};
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
      for (int ni = 0; ni < 16; ++ni) {
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
        if (MODE != 3) {
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
                       : "r"(a[mi][0]), "r"(a[mi][1]), "r"(a[mi][2]), "r"(a[mi][3]), "r"(b[ni][0] + t), "r"(b[ni][1]));
        } else {
          c0 = __uint_as_float(b[ni][0] + t); c1 = __uint_as_float(b[ni][1] ^ t); c2 = __uint_as_float(a[mi][0] + t); c3 = __uint_as_float(a[mi][1] - t);
        }
        if (MODE != 1) {
          unsigned& mm = m[(mi * 16 + ni) >> 3];
          mm = __funnelshift_l(__float_as_uint(c0), mm, 1);
          mm = __funnelshift_l(__float_as_uint(c1), mm, 1);
          mm = __funnelshift_l(__float_as_uint(c2), mm, 1);
          mm = __funnelshift_l(__float_as_uint(c3), mm, 1);
        } else {
          m[mi] ^= __float_as_uint(c0) + __float_as_uint(c3);
        }
      }
    }
    check += (m[0] | m[1]) ^ (m[2] + m[3]);
  }
  if (check == 0x12345678u) out[0] = check;
}

template <class F> void run(const char* name, F kern, unsigned* out) {
  const int tiles = 1 << 15;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  kern<<<148, 128>>>(out, 64); cudaDeviceSynchronize();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("%-44s FAILED: %s\n", name, cudaGetErrorString(e)); return; }
  cudaEventRecord(a); kern<<<148, 128>>>(out, tiles); cudaEventRecord(b); cudaEventSynchronize(b);
  e = cudaGetLastError();
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double cyc = ms * 1e-3 * 1.965e9 / tiles;
  printf("%-44s %8.3f ms  %7.1f SM cycles per 128x128x16 tile  (%s)\n", name, ms, cyc, cudaGetErrorString(e));
}
int main() {
  unsigned* out; cudaMalloc(&out, 64);
  run("mma.sync  HMMA + sign harvest", k_sync<0>, out);
  run("mma.sync  HMMA only", k_sync<1>, out);
  run("          sign harvest only (no MMA)", k_sync<3>, out);
  run("tcgen05   MMA + tcgen05.ld + sign harvest", k_tc5<0>, out);
  run("tcgen05   MMA + tcgen05.ld (no harvest)", k_tc5<2>, out);
  run("tcgen05   MMA only (no read-out)", k_tc5<1>, out);
  return 0;
}
