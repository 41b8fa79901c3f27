// match.cu — descriptor matching on sm_100a.
//
// Replaces the header template match_points<P1,P2> (reference src/my_utilities.h:70-120):
// for every query row i, best / second-best squared L2 distance over all rows j of the other
// set, lowest index on ties, accepted iff best < 0.2 && best/second < 0.8.
//
// Layout: descriptors packed row-major float[N][D] (the shim gathers them out of the
// reference's per-record heap VectorXf).
//
// Rounding contract: every distance that decides a result is evaluated in float32 with explicit
// round-to-nearest sub/mul/add in the order Eigen's SSE squaredNorm redux uses for VectorXf
// (SURVEY App. A.7); nothing is contracted to FMA, so best, second-best, indices and accept flags
// are bit-exact against the oracle on every path.
//
// Three paths (vo_match_dev picks by size; all end in the same merge / accept test / stable compaction):
//   1. match_scan_kernel<D>, D = 1..16, any size: one query row per thread (its D floats in registers),
//      grid.y = column splits sized so the grid fills whole waves, columns streamed through shared memory
//      in tiles; exact merge of the per-split (best, second, idx) triples in split order.  This is what
//      the frame loop uses (500 x 490 descriptors).
//   2. match_scan10_kernel, D = 10, >= 8192 rows: rows visited in Morton order, two columns per packed
//      f32x2 instruction, and an exact monotone lower bound (first half of Eigen's reduction tree) voted
//      across the warp before the second half is evaluated.
//   3. match_scan10_mma_kernel (+ match_scan10_indexed_kernel as its exact twin), D = 10, >= 8192
//      columns and enough work: rows and columns sorted on one Morton curve, 128-column tiles with
//      bounding boxes in two levels, tiles skipped by a per-row point-to-box bound, and inside a visited
//      tile a bf16 tensor-core filter (mma.sync m16n8k16: |a|^2 + |b|^2 - 2 a.b - bound, scaled so that it
//      is a guaranteed lower bound of the reference's float distance) that leaves a handful of columns
//      per row for the exact fp32 evaluation.  Derivation and the error budget: at the kernel.
#include "vo_device.cuh"

#include <float.h>
#include <stdlib.h>

#include <cub/device/device_radix_sort.cuh>

namespace {

#ifndef VO_MATCH_THREADS
#define VO_MATCH_THREADS 32
#endif
#ifndef VO_MATCH_TILE
#define VO_MATCH_TILE 128
#endif
// rows per CTA / columns per tile, measured on B200 for the indexed scan (exp/match_time.py, 1M x 1M and 131072 x 1M, ms):
// 128/256: 161 / 44.9   64/256: 134 / 43.6   64/128: 133 / 45.4   32/256: 145 / 40.7   32/128: 133 / 40.2
constexpr int kMatchThreads = VO_MATCH_THREADS;
constexpr int kTileRows = VO_MATCH_TILE;  // rows of B per shared-memory tile
constexpr int kMaxDim = 16;

template <int DIM>
__global__ void __launch_bounds__(kMatchThreads) match_scan_kernel(
    const float* __restrict__ A, long long row_begin, long long row_end, const float* __restrict__ B,
    long long n2, long long split_size, float* __restrict__ o_best, float* __restrict__ o_second,
    int* __restrict__ o_idx) {
  constexpr int DP = (DIM + 3) / 4 * 4;
  __shared__ __align__(16) float sB[kTileRows * DP];
  const long long rows = row_end - row_begin;
  const long long r = (long long)blockIdx.x * kMatchThreads + threadIdx.x;
  const bool valid = r < rows;
  float a[DIM];
#pragma unroll
  for (int k = 0; k < DIM; ++k) a[k] = valid ? __ldg(A + (row_begin + r) * DIM + k) : 0.f;
  float best = FLT_MAX, second = FLT_MAX;
  int idx = -1;
  const long long j_lo = (long long)blockIdx.y * split_size;
  const long long j_hi = (j_lo + split_size < n2) ? j_lo + split_size : n2;
  for (long long j0 = j_lo; j0 < j_hi; j0 += kTileRows) {
    const int cnt = (int)((j_hi - j0 < kTileRows) ? (j_hi - j0) : kTileRows);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * DIM; t += kMatchThreads) {
      const int jj = t / DIM, k = t - jj * DIM;
      sB[jj * DP + k] = __ldg(B + j0 * DIM + t);
    }
    __syncthreads();
    int jj = 0;
    for (; jj + 4 <= cnt; jj += 4) {
      const float d0 = sqdist_eigen<DIM>(a, sB + (jj + 0) * DP);
      const float d1 = sqdist_eigen<DIM>(a, sB + (jj + 1) * DP);
      const float d2 = sqdist_eigen<DIM>(a, sB + (jj + 2) * DP);
      const float d3 = sqdist_eigen<DIM>(a, sB + (jj + 3) * DP);
      const int j = (int)(j0 + jj);
      update_best(d0, j, best, second, idx);
      update_best(d1, j + 1, best, second, idx);
      update_best(d2, j + 2, best, second, idx);
      update_best(d3, j + 3, best, second, idx);
    }
    for (; jj < cnt; ++jj) update_best(sqdist_eigen<DIM>(a, sB + jj * DP), (int)(j0 + jj), best, second, idx);
  }
  if (valid) {
    const long long o = (long long)blockIdx.y * rows + r;
    o_best[o] = best;
    o_second[o] = second;
    o_idx[o] = idx;
  }
}

// ---------------------------------------------------------------- D = 10 (the reference's descriptor size)
// Packed f32x2 over COLUMN pairs: lane 0 of every packed op is column j, lane 1 is column j+1; the
// query row sits in registers as broadcast pairs (a_k, a_k).  The tile is stored per column pair as
// 20 floats in the order the Eigen redux consumes them: dims (0,4 | 2,6 | 1,5 | 3,7 | 8,9).
//
// Exact pruning: every term is >= 0 and float addition is monotone, so the partial sum
//   lb = (x0 + x4) + (x2 + x6)            (the first half of Eigen's reduction tree, same rounding)
// never exceeds the full distance d.  A column can only change (best, second, idx) when d < second,
// hence when lb < second; if no lane of the warp has such a column the other 6 dimensions and the
// update are skipped.  Results are bit-identical to the unpruned scan.
// (f2 / pack2 / add2 / sq2 / dim_slot10 / kPairFloats: vo_device.cuh, shared with the sequence kernel)

__global__ void __launch_bounds__(kMatchThreads) match_scan10_kernel(
    const float* __restrict__ A, long long row_begin, long long row_end, const float* __restrict__ B,
    long long n2, long long split_size, const unsigned* __restrict__ row_order, float* __restrict__ o_best,
    float* __restrict__ o_second, int* __restrict__ o_idx) {
  constexpr int DIM = 10;
  __shared__ __align__(16) float sB[(kTileRows / 2) * kPairFloats];
  const long long rows = row_end - row_begin;
  const long long slot = (long long)blockIdx.x * kMatchThreads + threadIdx.x;
  const bool valid = slot < rows;
  // row_order (optional): query rows sorted along a Morton curve of the lower-bound dimensions, so the
  // 32 rows of a warp prune the same columns; results are written back at the original row
  const long long r = (valid && row_order) ? (long long)row_order[slot] : slot;
  f2 a[DIM];  // slot order
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    const float v = valid ? __ldg(A + (row_begin + r) * DIM + k) : 0.f;
    a[dim_slot10(k)] = pack2(v, v);
  }
  float best = FLT_MAX, second = FLT_MAX;
  int idx = -1;
  const long long j_lo = (long long)blockIdx.y * split_size;
  const long long j_hi = (j_lo + split_size < n2) ? j_lo + split_size : n2;
  for (long long j0 = j_lo; j0 < j_hi; j0 += kTileRows) {
    const int cnt = (int)((j_hi - j0 < kTileRows) ? (j_hi - j0) : kTileRows);
    const int n_pairs = (cnt + 1) >> 1;
    __syncthreads();
    for (int t = threadIdx.x; t < n_pairs * 2 * DIM; t += kMatchThreads) {
      const int jj = t / DIM, k = t - jj * DIM;
      // an odd tail gets a copy of the last real column: it can only tie, and ties never win (strict <)
      const int src = (jj < cnt) ? jj : cnt - 1;
      sB[(jj >> 1) * kPairFloats + dim_slot10(k) * 2 + (jj & 1)] = __ldg(B + (j0 + src) * DIM + k);
    }
    __syncthreads();
    // two column pairs per step: their lower bounds are independent, which hides the LDS and FP latencies
    auto lower_bound = [&](const float4* rec) {
      const float4 v0 = rec[0], v1 = rec[1];  // dims (0,4) and (2,6) of both columns
      const f2 x0 = sq2(pack2(v0.x, v0.y), a[0]), x4 = sq2(pack2(v0.z, v0.w), a[1]);
      const f2 x2 = sq2(pack2(v1.x, v1.y), a[2]), x6 = sq2(pack2(v1.z, v1.w), a[3]);
      return add2(add2(x0, x4), add2(x2, x6));  // p0[0] + p0[2]
    };
    auto finish = [&](const float4* rec, f2 lb, int p) {
      const float4 v2 = rec[2], v3 = rec[3], v4 = rec[4];
      const f2 x1 = sq2(pack2(v2.x, v2.y), a[4]), x5 = sq2(pack2(v2.z, v2.w), a[5]);
      const f2 x3 = sq2(pack2(v3.x, v3.y), a[6]), x7 = sq2(pack2(v3.z, v3.w), a[7]);
      const f2 x8 = sq2(pack2(v4.x, v4.y), a[8]), x9 = sq2(pack2(v4.z, v4.w), a[9]);
      f2 d = add2(lb, add2(add2(x1, x5), add2(x3, x7)));  // (p0[0]+p0[2]) + (p0[1]+p0[3])
      d = add2(add2(d, x8), x9);
      float d0, d1;
      unpack2(d, d0, d1);
      const int j = (int)(j0 + 2 * p);
      update_best(d0, j, best, second, idx);
      if (2 * p + 1 < cnt) update_best(d1, j + 1, best, second, idx);
    };
    auto may_improve = [&](f2 lb) {
      float lb0, lb1;
      unpack2(lb, lb0, lb1);
      return __any_sync(0xffffffffu, (lb0 < second) || (lb1 < second));
    };
    int p = 0;
    for (; p + 2 <= n_pairs; p += 2) {
      const float4* recA = reinterpret_cast<const float4*>(sB + p * kPairFloats);
      const float4* recB = reinterpret_cast<const float4*>(sB + (p + 1) * kPairFloats);
      const f2 lbA = lower_bound(recA), lbB = lower_bound(recB);
      if (may_improve(lbA)) finish(recA, lbA, p);      // ascending column order keeps the lowest-index tie-break
      if (may_improve(lbB)) finish(recB, lbB, p + 1);
    }
    if (p < n_pairs) {
      const float4* rec = reinterpret_cast<const float4*>(sB + p * kPairFloats);
      const f2 lb = lower_bound(rec);
      if (may_improve(lb)) finish(rec, lb, p);
    }
  }
  if (valid) {
    const long long o = (long long)blockIdx.y * rows + r;  // r = original row
    o_best[o] = best;
    o_second[o] = second;
    o_idx[o] = idx;
  }
}

// ---- query-row ordering for the pruned scan: 4 x 8-bit Morton key over dims 0,4,2,6 (the lower-bound dims)
__device__ __forceinline__ unsigned ordered_u32(float f) {  // monotone float -> uint map
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_u32(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Flat, coalesced pass over rows x 10 floats as float2 (a 40-byte row keeps 8-byte alignment): with a grid stride that
// is a multiple of 5 pairs every thread stays on one pair of dimensions, keeps its four running values in registers
// and merges them once, through shared-memory atomics, at the end.  (The one-thread-per-row form this replaces read
// each 128-byte line ten times: 67 us for 42 MB; this one streams.)
__global__ void __launch_bounds__(256) match_minmax10_kernel(const float* __restrict__ A, long long row_begin, long long rows,
                                                             unsigned* __restrict__ mm /* [10] min, [10] max per dimension, ordered */) {
  __shared__ unsigned s_mm[20];
  if (threadIdx.x < 20) s_mm[threadIdx.x] = threadIdx.x < 10 ? 0xffffffffu : 0u;
  __syncthreads();
  const long long n_threads = (long long)gridDim.x * blockDim.x, stride = n_threads - n_threads % 5;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x, n_pairs = rows * 5;
  const float2* __restrict__ src = reinterpret_cast<const float2*>(A + row_begin * 10);
  unsigned lo0 = 0xffffffffu, hi0 = 0u, lo1 = 0xffffffffu, hi1 = 0u;
  auto take = [](float v, unsigned& lo, unsigned& hi) {
    if (v == v && fabsf(v) <= FLT_MAX) {  // NaN / inf do not stretch the ranges
      const unsigned u = ordered_u32(v);
      lo = min(lo, u);
      hi = max(hi, u);
    }
  };
  if ((reinterpret_cast<unsigned long long>(src) & 7ull) != 0) {  // a caller's view at an odd float offset: scalar loads
    const long long stride1 = n_threads - n_threads % 10, n_vals = rows * 10;
    if (t < stride1) {
      const float* __restrict__ s1 = A + row_begin * 10;
      for (long long p = t; p < n_vals; p += stride1) take(__ldg(s1 + p), lo0, hi0);
      const int d = (int)(t % 10);
      if (lo0 <= hi0) {
        atomicMin(&s_mm[d], lo0);
        atomicMax(&s_mm[10 + d], hi0);
      }
    }
  } else if (t < stride) {
    long long p = t;
    for (; p + 3 * stride < n_pairs; p += 4 * stride) {  // four loads in flight
      const float2 v0 = __ldg(src + p), v1 = __ldg(src + p + stride), v2 = __ldg(src + p + 2 * stride), v3 = __ldg(src + p + 3 * stride);
      take(v0.x, lo0, hi0); take(v0.y, lo1, hi1);
      take(v1.x, lo0, hi0); take(v1.y, lo1, hi1);
      take(v2.x, lo0, hi0); take(v2.y, lo1, hi1);
      take(v3.x, lo0, hi0); take(v3.y, lo1, hi1);
    }
    for (; p < n_pairs; p += stride) {
      const float2 v = __ldg(src + p);
      take(v.x, lo0, hi0);
      take(v.y, lo1, hi1);
    }
    const int d = 2 * (int)(t % 5);
    if (lo0 <= hi0) {
      atomicMin(&s_mm[d], lo0);
      atomicMax(&s_mm[10 + d], hi0);
    }
    if (lo1 <= hi1) {
      atomicMin(&s_mm[d + 1], lo1);
      atomicMax(&s_mm[11 + d], hi1);
    }
  }
  __syncthreads();
  if (threadIdx.x < 10) atomicMin(&mm[threadIdx.x], s_mm[threadIdx.x]);
  else if (threadIdx.x < 20) atomicMax(&mm[threadIdx.x], s_mm[threadIdx.x]);
}

// midpoint of dimension k's finite range over rows and columns: the filter works on centred descriptors
// (distances do not change, but its error bound is relative to the squared norms)
__device__ __forceinline__ float range_center(const unsigned* __restrict__ mm, int k) {
  const unsigned ul = mm[k], uh = mm[10 + k];
  if (ul > uh) return 0.f;  // no finite value in this dimension
  return 0.5f * from_ordered_u32(ul) + 0.5f * from_ordered_u32(uh);
}

__device__ __forceinline__ unsigned spread8(unsigned v) {  // abcdefgh -> a000b000...h (every 4th bit)
  v &= 0xffu;
  v = (v | (v << 12)) & 0x000f000fu;
  v = (v | (v << 6)) & 0x03030303u;
  v = (v | (v << 3)) & 0x11111111u;
  return v;
}

// magnitudes (of the centred descriptors) the bf16 filter's error analysis covers; outside them the exact scan runs
constexpr float kFilterMaxAbs = 1e15f, kFilterMinAbs = 1e-15f;

// range_flag (nullable): also check the row's ten centred values against the filter's magnitude limits - the keys
// pass reads the rows anyway (two separate 42 us passes over rows and columns before)
__global__ void match_keys10_kernel(const float* __restrict__ A, long long row_begin, long long rows,
                                    const unsigned* __restrict__ mm, unsigned* __restrict__ keys,
                                    unsigned* __restrict__ ids, int* __restrict__ range_flag) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (range_flag) {
    bool bad = false;
    if (r < rows) {
#pragma unroll
      for (int k = 0; k < 10; ++k) {
        const float a = fabsf(__ldg(A + (row_begin + r) * 10 + k) - range_center(mm, k));
        bad |= (a <= FLT_MAX) && ((a > kFilterMaxAbs) || (a != 0.f && a < kFilterMinAbs));
      }
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(range_flag, 1);
  }
  if (r >= rows) return;
  const int dims[4] = {0, 4, 2, 6};
  unsigned key = 0;
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const float lo = from_ordered_u32(mm[dims[d]]), hi = from_ordered_u32(mm[10 + dims[d]]);
    const float v = __ldg(A + (row_begin + r) * 10 + dims[d]);
    float q = (hi > lo) ? (v - lo) / (hi - lo) * 255.f : 0.f;
    q = (q == q) ? fminf(fmaxf(q, 0.f), 255.f) : 0.f;
    key |= spread8((unsigned)q) << d;
  }
  keys[r] = key;
  ids[r] = (unsigned)r;
}

// ---- spatial index over the columns (D = 10, large problems): columns sorted along the same Morton curve as
// the query rows, stored as pair records (the shared-memory tile layout, so a tile load is a straight float4
// copy), plus the bounding box of every 256-column tile in the four lower-bound dimensions.
__global__ void match_gather10_kernel(const float* __restrict__ B, const unsigned* __restrict__ order, long long n2,
                                      float* __restrict__ rec, int* __restrict__ orig) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // sorted position
  const long long n2p = (n2 + 1) & ~1ll;
  if (j >= n2p) return;
  const long long src = order[j < n2 ? j : n2 - 1];  // an odd tail is padded with a copy of the last column
  orig[j] = (int)src;
#pragma unroll
  for (int k = 0; k < 10; ++k) rec[(j >> 1) * kPairFloats + dim_slot10(k) * 2 + (j & 1)] = __ldg(B + src * 10 + k);
}

__global__ void __launch_bounds__(kTileRows) match_tilebox10_kernel(const float* __restrict__ rec, long long n2,
                                                                    float* __restrict__ box /* [tiles][8] */) {
  __shared__ float s_lo[kTileRows / 32][4], s_hi[kTileRows / 32][4];
  const long long j = (long long)blockIdx.x * kTileRows + threadIdx.x;
  float lo[4], hi[4];
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const float v = (j < n2) ? rec[(j >> 1) * kPairFloats + d * 2 + (j & 1)] : NAN;  // slots 0..3 = dims 0,4,2,6
    lo[d] = hi[d] = v;  // fminf / fmaxf ignore NaN
  }
#pragma unroll
  for (int d = 0; d < 4; ++d)
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      s_lo[threadIdx.x >> 5][d] = lo[d];
      s_hi[threadIdx.x >> 5][d] = hi[d];
    }
  __syncthreads();
  if (threadIdx.x < 4) {
    float l = s_lo[0][threadIdx.x], h = s_hi[0][threadIdx.x];
    for (int w = 1; w < kTileRows / 32; ++w) {
      l = fminf(l, s_lo[w][threadIdx.x]);
      h = fmaxf(h, s_hi[w][threadIdx.x]);
    }
    box[(size_t)blockIdx.x * 8 + threadIdx.x] = l;
    box[(size_t)blockIdx.x * 8 + 4 + threadIdx.x] = h;
  }
}

#ifndef VO_MATCH_SUPER
#define VO_MATCH_SUPER 16
#endif
constexpr int kSuper = VO_MATCH_SUPER;  // tiles per super-tile (second level of the box hierarchy)

#ifdef VO_MATCH_COUNTERS  // exp/ builds only: how much of the column set the indexed scan really touches
__device__ unsigned long long g_match_counters[4];  // super-tiles entered, tiles scanned, pair records finished, pair records bounded
#define VO_COUNT(i, n) do { if (threadIdx.x == 0) atomicAdd(&g_match_counters[i], (unsigned long long)(n)); } while (0)
#else
#define VO_COUNT(i, n) do { } while (0)
#endif

__global__ void match_superbox10_kernel(const float* __restrict__ box, long long n_tiles, float* __restrict__ sbox) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_super = (n_tiles + kSuper - 1) / kSuper;
  if (g >= n_super) return;
  float lo[4] = {NAN, NAN, NAN, NAN}, hi[4] = {NAN, NAN, NAN, NAN};
  for (long long t = g * kSuper; t < n_tiles && t < (g + 1) * kSuper; ++t)
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      lo[d] = fminf(lo[d], box[t * 8 + d]);
      hi[d] = fmaxf(hi[d], box[t * 8 + 4 + d]);
    }
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    sbox[g * 8 + d] = lo[d];
    sbox[g * 8 + 4 + d] = hi[d];
  }
}

// my_utilities.h:93-99 when the columns are NOT visited in index order: equal distances must still resolve to the
// lowest index, and `second` is the second smallest value of the multiset (both are order independent).
__device__ __forceinline__ void update_best_tie(float d, int j, float& best, float& second, int& idx) {
  const bool lt = (d < best) || (d == best && j < idx);
  const float s2 = (d < second) ? d : second;
  second = lt ? best : s2;
  idx = lt ? j : idx;
  best = lt ? d : best;
}

// Scan with tile skipping. A CTA owns 32 Morton-adjacent query rows (lane = row) and `n_warps` warps that all hold
// the same rows; the column super-tiles are visited outward from the rows' own position on the curve, warp w taking
// every n_warps-th super-tile of that order.  A (super-)tile is skipped when, for every row, the squared distance
// from the row's point to the tile's box (in the lower-bound dimensions) exceeds the row's second-best bound.
//   exactness: lb_float(row, col) >= lb_real * (1 - 4 ulp) >= box_real * (1 - 4 ulp) > box_float * (1 - 1e-5),
//   so box_float * (1 - 1e-5) > bound  =>  d >= lb_float > bound >= final second >= final best: the column can
//   change neither value nor (being strictly farther than the best) the lowest-index tie rule.
//   `bound` = min(own second, s_bound[row]); s_bound is the smallest second-best any warp of the CTA has published:
//   a second-best over a subset of the columns is never below the second-best over all of them.
// The warps never synchronise with one another until the final merge (tile buffers are per warp).
// n_warps > 1 is for row counts too small to fill the machine with one warp per 32 rows: a single warp per scheduler
// runs this loop latency-bound (measured 18 ms for 8192 x 1M rows with 256 lone warps).
constexpr int kTileBytes = (kTileRows / 2) * kPairFloats * 4 + kTileRows * 4;  // pair records + original indices
constexpr int kMaxScanWarps = 16;

#ifndef VO_SCAN_MINB
#define VO_SCAN_MINB 1
#endif
__global__ void __launch_bounds__(32 * kMaxScanWarps, VO_SCAN_MINB) match_scan10_indexed_kernel(
    const float* __restrict__ A, long long row_begin, long long rows, const unsigned* __restrict__ row_order,
    const unsigned* __restrict__ row_keys_sorted, const float* __restrict__ rec, const int* __restrict__ orig,
    const float* __restrict__ box, const float* __restrict__ sbox, const unsigned* __restrict__ col_keys_sorted,
    long long n2, const int* __restrict__ range_flag, float* __restrict__ o_best, float* __restrict__ o_second,
    int* __restrict__ o_idx) {
  if (*range_flag == 0) return;  // the tensor-core filtered scan (below) handles this call
  constexpr int DIM = 10;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ long long s_t0;
  __shared__ int s_bound[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  float* sB = reinterpret_cast<float*>(smem_raw + (size_t)warp * kTileBytes);
  int* sOrig = reinterpret_cast<int*>(sB + (kTileRows / 2) * kPairFloats);
  const long long slot = (long long)blockIdx.x * 32 + lane;
  const bool valid = slot < rows;
  const long long r = valid ? (long long)row_order[slot] : 0;
  f2 a[DIM];
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    const float v = valid ? __ldg(A + (row_begin + r) * DIM + k) : NAN;
    a[dim_slot10(k)] = pack2(v, v);
  }
  if (threadIdx.x == 0) {  // where this CTA's rows sit among the sorted columns
    const long long mid = min((long long)blockIdx.x * 32 + 16, rows - 1);
    const unsigned key = row_keys_sorted[mid];
    long long a0 = 0, a1 = n2;
    while (a0 < a1) {
      const long long m = (a0 + a1) >> 1;
      if (col_keys_sorted[m] < key) a0 = m + 1;
      else a1 = m;
    }
    s_t0 = min(a0, n2 - 1) / (kTileRows * kSuper);  // home super-tile
  }
  if (threadIdx.x < 32) s_bound[threadIdx.x] = __float_as_int(FLT_MAX);
  __syncthreads();
  float best = FLT_MAX, second = FLT_MAX, bound = FLT_MAX;
  int idx = -1;
  const long long n_tiles = (n2 + kTileRows - 1) / kTileRows;
  const long long n_super = (n_tiles + kSuper - 1) / kSuper;
  const long long home = s_t0, n_up = n_super - home, n_both = min(home, n_up);
  // Per lane: squared distance from the lane's own point to the box, shrunk by the rounding margin, against the
  // lane's bound.  (One box around all rows of the CTA is useless for the CTAs that straddle a jump of the curve:
  // they would walk every tile. Measured: 127 ms -> 101 ms at 1M x 1M, 40 -> 26 ms at 131072 x 1M.  Visiting the
  // tiles in rings of increasing box distance instead of curve order was measured slower: 171 ms.)
  float pt[4];
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    float w;
    unpack2(a[d], pt[d], w);
  }
  auto box_can_matter = [&](const float* bx, long long i) {
    const float4 bl = __ldg(reinterpret_cast<const float4*>(bx) + 2 * i), bh = __ldg(reinterpret_cast<const float4*>(bx) + 2 * i + 1);
    const float g0 = fmaxf(0.f, fmaxf(bl.x - pt[0], pt[0] - bh.x)), g1 = fmaxf(0.f, fmaxf(bl.y - pt[1], pt[1] - bh.y));
    const float g2 = fmaxf(0.f, fmaxf(bl.z - pt[2], pt[2] - bh.z)), g3 = fmaxf(0.f, fmaxf(bl.w - pt[3], pt[3] - bh.w));
    const float bb = (g0 * g0 + g1 * g1 + g2 * g2 + g3 * g3) * 0.99999f;
    return __any_sync(0xffffffffu, valid && !(bb > bound));  // NaN bounds never skip
  };
  for (long long step = warp; step < n_super; step += n_warps) {
    // step -> super-tile: home, home-1, home+1, home-2, ... then the rest of the longer side
    long long g;
    if (step < 2 * n_both) g = (step & 1) ? home - ((step + 1) >> 1) : home + (step >> 1);
    else g = (n_up > home) ? home + (step - n_both) : home - 1 - (step - n_both);
    if (n_warps > 1) bound = fminf(second, __int_as_float(s_bound[lane]));
    if (!box_can_matter(sbox, g)) continue;  // 16 tiles at once
    VO_COUNT(0, 1);
    const long long t_end = min(n_tiles, (g + 1) * kSuper);
    for (long long t = g * kSuper; t < t_end; ++t) {
      if (!box_can_matter(box, t)) continue;
      VO_COUNT(1, 1);
      const long long j0 = t * kTileRows;
      const int cnt = (int)((n2 - j0 < kTileRows) ? (n2 - j0) : kTileRows);
      const int n_pairs = (cnt + 1) >> 1;
      __syncwarp();
      {
        const float4* src = reinterpret_cast<const float4*>(rec + (j0 >> 1) * kPairFloats);
        float4* dst = reinterpret_cast<float4*>(sB);
        for (int q = lane; q < n_pairs * (kPairFloats / 4); q += 32) dst[q] = __ldg(src + q);
        for (int q = lane; q < 2 * n_pairs; q += 32) sOrig[q] = __ldg(orig + j0 + q);
      }
      __syncwarp();
      auto lower_bound = [&](const float4* rc) {
        const float4 v0 = rc[0], v1 = rc[1];
        const f2 x0 = sq2(pack2(v0.x, v0.y), a[0]), x4 = sq2(pack2(v0.z, v0.w), a[1]);
        const f2 x2 = sq2(pack2(v1.x, v1.y), a[2]), x6 = sq2(pack2(v1.z, v1.w), a[3]);
        return add2(add2(x0, x4), add2(x2, x6));
      };
      auto finish = [&](const float4* rc, f2 lb, int p) {
        const float4 v2 = rc[2], v3 = rc[3], v4 = rc[4];
        const f2 x1 = sq2(pack2(v2.x, v2.y), a[4]), x5 = sq2(pack2(v2.z, v2.w), a[5]);
        const f2 x3 = sq2(pack2(v3.x, v3.y), a[6]), x7 = sq2(pack2(v3.z, v3.w), a[7]);
        const f2 x8 = sq2(pack2(v4.x, v4.y), a[8]), x9 = sq2(pack2(v4.z, v4.w), a[9]);
        f2 d = add2(lb, add2(add2(x1, x5), add2(x3, x7)));
        d = add2(add2(d, x8), x9);
        float d0, d1;
        unpack2(d, d0, d1);
        // most finished distances still change nothing: only then touch the index table and the running triple
        if (!__any_sync(0xffffffffu, (d0 <= bound) || (d1 <= bound))) return;
        update_best_tie(d0, sOrig[2 * p], best, second, idx);
        if (2 * p + 1 < cnt) update_best_tie(d1, sOrig[2 * p + 1], best, second, idx);
        bound = fminf(bound, second);
      };
      auto may_improve = [&](f2 lb) {
        float lb0, lb1;
        unpack2(lb, lb0, lb1);
        return __any_sync(0xffffffffu, (lb0 <= bound) || (lb1 <= bound));
      };
      int p = 0;
      for (; p + 2 <= n_pairs; p += 2) {
        const float4* recA = reinterpret_cast<const float4*>(sB + p * kPairFloats);
        const float4* recB = reinterpret_cast<const float4*>(sB + (p + 1) * kPairFloats);
        const f2 lbA = lower_bound(recA), lbB = lower_bound(recB);
        const bool mA = may_improve(lbA), mB = may_improve(lbB);
        VO_COUNT(2, (int)mA + (int)mB);
        if (mA) finish(recA, lbA, p);
        if (mB) finish(recB, lbB, p + 1);
      }
      if (p < n_pairs) {
        const float4* rc = reinterpret_cast<const float4*>(sB + p * kPairFloats);
        const f2 lb = lower_bound(rc);
        if (may_improve(lb)) finish(rc, lb, p);
      }
      VO_COUNT(3, n_pairs);
      if (n_warps > 1) {  // publish / pick up the CTA-wide bound once per scanned tile
        if (second < FLT_MAX) atomicMin(&s_bound[lane], __float_as_int(second));  // second >= 0: int order = float order
        bound = fminf(second, __int_as_float(s_bound[lane]));
      }
    }
  }
  if (n_warps > 1) {
    // merge the warps' triples: the bests through the tie-aware rule (their indices decide ties), the seconds as
    // plain values (the column behind a warp's second has a higher index than that warp's best at equal distance)
    __syncthreads();
    float* m_best = reinterpret_cast<float*>(smem_raw);
    float* m_second = m_best + 32 * kMaxScanWarps;
    int* m_idx = reinterpret_cast<int*>(m_second + 32 * kMaxScanWarps);
    m_best[warp * 32 + lane] = best;
    m_second[warp * 32 + lane] = second;
    m_idx[warp * 32 + lane] = idx;
    __syncthreads();
    if (warp != 0) return;
    best = second = FLT_MAX;
    idx = -1;
    float s_min = FLT_MAX;
    for (int w = 0; w < n_warps; ++w) {
      if (m_idx[w * 32 + lane] >= 0) update_best_tie(m_best[w * 32 + lane], m_idx[w * 32 + lane], best, second, idx);
      s_min = fminf(s_min, m_second[w * 32 + lane]);
    }
    second = fminf(second, s_min);
  }
  if (valid) {
    o_best[r] = best;
    o_second[r] = second;
    o_idx[r] = idx;
  }
}

// ---- tensor-core distance filter (D = 10, indexed path) --------------------------------------------------------
// The scan above spends ~40 issue slots per (32 rows x 2 columns) and finds that all but a handful of columns per
// row are farther than the row's second-best.  Deciding THAT does not need the reference's float evaluation order:
// a guaranteed lower bound of the distance is enough, and ||a-b||^2 = |a|^2 + |b|^2 - 2 a.b is one K = 16 bf16 MMA
// per 16 rows x 8 columns (mma.sync m16n8k16: the accumulators land in registers, which is what a compare-only
// epilogue wants; a TMEM round trip per K = 16 tile would be read-bandwidth bound).
//   k = 0..9 : -2 bf16(a_k)        x  bf16(b_k)
//   k = 10,11: hi, lo of (1-eps)|a|^2  x  1
//   k = 12,13: 1                    x  hi, lo of (1-eps)|b|^2
//   k = 14,15: -hi, -lo of the row's current bound (slightly inflated)  x  1
//   v = (1-eps)(|a|^2+|b|^2) - 2 sum bf16(a_k) bf16(b_k)   (+ accumulation error);  output = v - bound
// bf16 carries 8 significant bits: |bf16(x) - x| <= 2^-8 |x| (round to nearest), so
// 2 |sum a^b^ - sum ab| <= (2^-7 + 2^-16) sum 2|a_k b_k| <= (2^-7 + 2^-16)(|a|^2 + |b|^2); the hi/lo splits (2^-16 each),
// the fp32 norms, the accumulator and the reference's own float evaluation add less than 2^-14 (|a|^2+|b|^2) together.
// With eps = 2^-7 + 2^-12 that gives
//   v < d_reference      for every finite pair whose magnitudes pass the range check of match_keys10_kernel (no overflow of the
//                        norms, no underflow of the products; otherwise the exact scan above runs instead),
// (a first version used 2^-8 + 2^-12, taking bf16's unit roundoff for 2^-9: near-duplicate rows, whose rounding errors
// all point the same way, then lost their true second-best - found by exp/match_stress.py, now tests/test_gpu_match.py)
// so v > bound  =>  d > bound >= final second-best: the column cannot change the row's result (strict, ties safe).
// The filter sees descriptors minus the midpoint c of each dimension's range (same distances; fl(x - c) moves a
// distance by < 2^-22 of the centred norms), so a common offset of the data does not loosen the bound.
// Columns with v <= bound are marked in a per-row bit mask and evaluated exactly (reference order, fp32) by the
// row's lane.  NaN/inf rows or columns give v = NaN/inf: never marked, exactly like `d < best` with a NaN/inf d.
constexpr float kFilterEps = 0.0078125f + 0.000244140625f;  // 2^-7 + 2^-12
__device__ __forceinline__ unsigned bf16x2_rn(float lo, float hi) {
  unsigned r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// (-hi, -lo) of a row's bound as two bf16: hi + lo >= bound (the 2^-14 inflation covers the 2^-16 split error),
// clamped below bf16's largest finite value so that "no bound yet" stays finite
__device__ __forceinline__ unsigned filter_bound_word(float bound) {
  const float t = fminf(bound * 1.00006103515625f, 3.0e38f);
  const float hi = __uint_as_float(bf16x2_rn(t, 0.f) << 16);
  return bf16x2_rn(-hi, -(t - hi));
}

// the 16 bf16 of one point (natural dimension order), as 8 words: w[i] = (k = 2i, 2i+1)
__device__ __forceinline__ void filter_words(const float v[10], bool is_row, unsigned w[8]) {
  float n = 0.f;
#pragma unroll
  for (int k = 0; k < 10; ++k) n = fmaf(v[k], v[k], n);
  n *= (1.f - kFilterEps);
  const float hi = __uint_as_float(bf16x2_rn(n, 0.f) << 16);
  const float lo = n - hi;
  const float s = is_row ? -2.f : 1.f;
#pragma unroll
  for (int i = 0; i < 5; ++i) w[i] = bf16x2_rn(s * v[2 * i], s * v[2 * i + 1]);
  const unsigned ones = 0x3F803F80u, norm = bf16x2_rn(hi, lo);
  w[5] = is_row ? norm : ones;
  w[6] = is_row ? ones : norm;
  w[7] = is_row ? filter_bound_word(FLT_MAX) : ones;
}

// per column (sorted position j, padded to whole tiles with NaN): uint2[4], entry t = words (t, t+4) = the
// m16n8k16 B fragment of lane 4*(j%8)+t, so a warp's fragment load for 8 columns is one coalesced 256-byte read
__global__ void match_gatherfrag10_kernel(const float* __restrict__ B, const unsigned* __restrict__ order, long long n2,
                                          long long n_padded, const unsigned* __restrict__ mm, uint4* __restrict__ frag) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_padded) return;
  unsigned w[8];
  if (j < n2) {
    const long long src = order[j];
    float v[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) v[k] = __ldg(B + src * 10 + k) - range_center(mm, k);
    filter_words(v, false, w);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = 0x7FC07FC0u;
  }
  frag[2 * j] = make_uint4(w[0], w[4], w[1], w[5]);
  frag[2 * j + 1] = make_uint4(w[2], w[6], w[3], w[7]);
}

__device__ __forceinline__ void mma_bf16_16816(float d[4], const unsigned a[4], uint2 b) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y), "f"(0.f));
}
__device__ __forceinline__ float min3(float a, float b, float c) {  // FMNMX3; NaN operands are ignored
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

#ifndef VO_MMA_BUFS
#define VO_MMA_BUFS 1
#endif
#ifndef VO_MATCH_DENSE_GUARD
#define VO_MATCH_DENSE_GUARD 1
#endif
constexpr int kDenseMinPerRow = 48;  // survivors of one row in a 128-column tile from which the outright scan is cheaper
constexpr int kDenseSkip = 7;        // tiles evaluated outright, once the data has shown itself filter-proof, before the filter is tried again
constexpr int kDenseStreak = 3;      // consecutive dense tiles that count as filter-proof data (the first tiles of every walk are dense anyway)
constexpr int kFragBufs = VO_MMA_BUFS;             // tile fragment buffers per warp (TMA bulk copies in flight)
constexpr int kFragTileBytes = kTileRows * 32;     // 128 columns x 16 bf16
constexpr int kMmaWarpSmem = 2048 + kFragBufs * kFragTileBytes + 64;

// my_utilities.h:85-91 for one column of a pair record (stride 2 floats), in Eigen's reduction order - the same
// tree, the same single roundings as the packed scan (sq2 / add2)
__device__ __forceinline__ float exact_sqdist10(const float* __restrict__ rc, const float (&a)[10]) {
#define VO_Q(sI) ([&]() { const float df = __fsub_rn(__ldg(rc + 2 * (sI)), a[sI]); return __fmul_rn(df, df); })()
  const float q0 = VO_Q(0), q1 = VO_Q(1), q2 = VO_Q(2), q3 = VO_Q(3), q4 = VO_Q(4);
  const float q5 = VO_Q(5), q6 = VO_Q(6), q7 = VO_Q(7), q8 = VO_Q(8), q9 = VO_Q(9);
#undef VO_Q
  const float d = __fadd_rn(__fadd_rn(__fadd_rn(q0, q1), __fadd_rn(q2, q3)), __fadd_rn(__fadd_rn(q4, q5), __fadd_rn(q6, q7)));
  return __fadd_rn(__fadd_rn(d, q8), q9);
}
// The dense-tile fallback of match_scan10_mma_kernel (see the guard there): every column of a 128-column tile for
// all 32 rows of the warp, two columns per packed instruction, in the exact twin's arithmetic.  Out of line on
// purpose: inlined into the tile loop it cost the common, filtered path 12 % (1M x 1M: 16.3 vs 14.5 ms) through
// register allocation alone.
struct Best3 {
  float best, second, bound;
  int idx;
};
__device__ __noinline__ Best3 dense_tile_scan(const float* __restrict__ rec, const int* __restrict__ orig, long long j0,
                                              int cnt, float a0, float a1, float a2, float a3, float a4, float a5,
                                              float a6, float a7, float a8, float a9, Best3 c) {
  const int n_pairs = (cnt + 1) >> 1;
  const float4* src = reinterpret_cast<const float4*>(rec + (j0 >> 1) * kPairFloats);
  const f2 b0 = pack2(a0, a0), b1 = pack2(a1, a1), b2 = pack2(a2, a2), b3 = pack2(a3, a3), b4 = pack2(a4, a4);
  const f2 b5 = pack2(a5, a5), b6 = pack2(a6, a6), b7 = pack2(a7, a7), b8 = pack2(a8, a8), b9 = pack2(a9, a9);
  for (int p = 0; p < n_pairs; ++p) {
    const float4 v0 = __ldg(src + 5 * p), v1 = __ldg(src + 5 * p + 1), v2 = __ldg(src + 5 * p + 2);
    const float4 v3 = __ldg(src + 5 * p + 3), v4 = __ldg(src + 5 * p + 4);
    const f2 x0 = sq2(pack2(v0.x, v0.y), b0), x4 = sq2(pack2(v0.z, v0.w), b1);
    const f2 x2 = sq2(pack2(v1.x, v1.y), b2), x6 = sq2(pack2(v1.z, v1.w), b3);
    const f2 x1 = sq2(pack2(v2.x, v2.y), b4), x5 = sq2(pack2(v2.z, v2.w), b5);
    const f2 x3 = sq2(pack2(v3.x, v3.y), b6), x7 = sq2(pack2(v3.z, v3.w), b7);
    const f2 x8 = sq2(pack2(v4.x, v4.y), b8), x9 = sq2(pack2(v4.z, v4.w), b9);
    f2 d = add2(add2(add2(x0, x4), add2(x2, x6)), add2(add2(x1, x5), add2(x3, x7)));
    d = add2(add2(d, x8), x9);
    float d0, d1;
    unpack2(d, d0, d1);
    if (!__any_sync(0xffffffffu, (d0 <= c.bound) || (d1 <= c.bound))) continue;
    if (d0 <= c.bound) update_best_tie(d0, __ldg(orig + j0 + 2 * p), c.best, c.second, c.idx);
    if (2 * p + 1 < cnt && d1 <= c.bound) update_best_tie(d1, __ldg(orig + j0 + 2 * p + 1), c.best, c.second, c.idx);
    c.bound = fminf(c.bound, c.second);
  }
  return c;
}

// Same walk, same bounds and same merge as match_scan10_indexed_kernel; the per-tile work is the filter.
#ifndef VO_MMA_LB
#define VO_MMA_LB 32 * kMaxScanWarps
#endif
template <bool MULTI>  // MULTI: more than one warp per 32-row group (the CTA-wide bound and the final merge)
__global__ void __launch_bounds__(VO_MMA_LB, 2) match_scan10_mma_kernel(
    const float* __restrict__ A, long long row_begin, long long rows, const unsigned* __restrict__ row_order,
    const unsigned* __restrict__ row_keys_sorted, const float* __restrict__ rec, const int* __restrict__ orig,
    const uint2* __restrict__ frag, const float* __restrict__ box, const float* __restrict__ sbox,
    const unsigned* __restrict__ col_keys_sorted, long long n2, const unsigned* __restrict__ mm,
    const int* __restrict__ range_flag, float* __restrict__ o_best, float* __restrict__ o_second,
    int* __restrict__ o_idx) {
  if (*range_flag != 0) return;  // magnitudes outside the filter's error analysis: the exact scan runs instead
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ long long s_t0;
  __shared__ int s_bound[32];
  __shared__ float s_wbest[kMaxScanWarps][32];  // every warp's best per row (n_warps > 1)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int g = lane >> 2, tq = lane & 3;
  unsigned* sA = reinterpret_cast<unsigned*>(smem_raw);                         // [32 rows][8 words]
  unsigned char* my_smem = smem_raw + 1024 + warp * kMmaWarpSmem;
  unsigned* mask = reinterpret_cast<unsigned*>(my_smem);            // [32 rows][4 words]: surviving columns of a tile
  float4* sb_stage = reinterpret_cast<float4*>(my_smem + 512);      // boxes of the next 32 super-tiles of the walk
  float4* tb_stage = reinterpret_cast<float4*>(my_smem + 1536);     // boxes of the 16 tiles of one super-tile
  unsigned char* fbuf = my_smem + 2048;                             // [kFragBufs][4 KB] column fragments of a tile
  unsigned long long* fbar = reinterpret_cast<unsigned long long*>(my_smem + 2048 + kFragBufs * kFragTileBytes);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kFragBufs; ++i) mbar_init(&fbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned fphase = 0;  // bit i: parity the next wait on buffer i expects
  const long long slot = (long long)blockIdx.x * 32 + lane;
  const bool valid = slot < rows;
  const long long r = valid ? (long long)row_order[slot] : 0;
  float a[10];  // slot order (pair-record order), fp32: the exact evaluation
  {
    float v[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      v[k] = valid ? __ldg(A + (row_begin + r) * 10 + k) : NAN;
      a[dim_slot10(k)] = v[k];
    }
    if (warp == 0) {
      unsigned w[8];
#pragma unroll
      for (int k = 0; k < 10; ++k) v[k] -= range_center(mm, k);
      filter_words(v, true, w);
#pragma unroll
      for (int i = 0; i < 8; ++i) sA[lane * 8 + i] = w[i];
    }
  }
  *reinterpret_cast<uint4*>(mask + lane * 4) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {  // where this CTA's rows sit among the sorted columns
    const long long mid = min((long long)blockIdx.x * 32 + 16, rows - 1);
    const unsigned key = row_keys_sorted[mid];
    long long a0 = 0, a1 = n2;
    while (a0 < a1) {
      const long long m = (a0 + a1) >> 1;
      if (col_keys_sorted[m] < key) a0 = m + 1;
      else a1 = m;
    }
    s_t0 = min(a0, n2 - 1) / (kTileRows * kSuper);  // home super-tile
  }
  if (threadIdx.x < 32) s_bound[threadIdx.x] = __float_as_int(FLT_MAX);
  if (MULTI) s_wbest[warp][lane] = FLT_MAX;
  __syncthreads();
  unsigned afrag[2][4];
#pragma unroll
  for (int mb = 0; mb < 2; ++mb) {
    afrag[mb][0] = sA[(mb * 16 + g) * 8 + tq];
    afrag[mb][1] = sA[(mb * 16 + g + 8) * 8 + tq];
    afrag[mb][2] = sA[(mb * 16 + g) * 8 + tq + 4];
    afrag[mb][3] = sA[(mb * 16 + g + 8) * 8 + tq + 4];
  }
  float best = FLT_MAX, second = FLT_MAX, bound = FLT_MAX;
  int idx = -1;
  // n_warps > 1: the warps scan disjoint column sets, so the second smallest of {every warp's best} U {every warp's
  // second} is a second-best of the union and bounds every warp's scan - far tighter, early in the walks, than the
  // smallest published `second` alone (two warps that each hold one near neighbour already pin it).  Stale reads
  // are only larger, i.e. conservative.
  auto cta_bound = [&]() {
    float m1 = FLT_MAX, m2 = FLT_MAX;
    for (int w = 0; w < n_warps; ++w) {
      const float b = *(volatile float*)&s_wbest[w][lane];
      m2 = fminf(m2, fmaxf(m1, b));
      m1 = fminf(m1, b);
    }
    return fminf(m2, __int_as_float(*(volatile int*)&s_bound[lane]));
  };
  int dense_skip = 0;  // tiles still to be evaluated outright before the filter is probed again (warp-uniform)
  int dense_streak = 0;  // consecutive filtered tiles that turned out dense
  // The rows' bounds ride in the A operand (k = 14,15, held by the threads with tq == 3): the MMA output is v - bound,
  // a column survives iff its output is <= 0, and a running 3-input minimum over 4 column blocks needs one
  // comparison.  (The bound joins the sum as two more terms; if it dwarfs the others the sign is decided anyway.)
  auto refresh_thr = [&]() {
    const unsigned tw = filter_bound_word(bound);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned w = __shfl_sync(0xffffffffu, tw, g + 8 * i);
      if (tq == 3) afrag[i >> 1][2 + (i & 1)] = w;
    }
  };
  const long long n_tiles = (n2 + kTileRows - 1) / kTileRows;
  const long long n_super = (n_tiles + kSuper - 1) / kSuper;
  const long long home = s_t0, n_up = n_super - home, n_both = min(home, n_up);
  const float pt[4] = {a[0], a[1], a[2], a[3]};
  // boxes are staged through shared memory 32 at a time (one coalesced read instead of one L2 round trip per test)
  auto box_can_matter = [&](const float4* bx) {
    const float4 bl = bx[0], bh = bx[1];
    const float g0 = fmaxf(0.f, fmaxf(bl.x - pt[0], pt[0] - bh.x)), g1 = fmaxf(0.f, fmaxf(bl.y - pt[1], pt[1] - bh.y));
    const float g2 = fmaxf(0.f, fmaxf(bl.z - pt[2], pt[2] - bh.z)), g3 = fmaxf(0.f, fmaxf(bl.w - pt[3], pt[3] - bh.w));
    const float bb = (g0 * g0 + g1 * g1 + g2 * g2 + g3 * g3) * 0.99999f;
    return __any_sync(0xffffffffu, valid && !(bb > bound));  // NaN bounds never skip
  };
  auto step_to_super = [&](long long step) {  // home, home-1, home+1, home-2, ... then the rest of the longer side
    if (step < 2 * n_both) return (step & 1) ? home - ((step + 1) >> 1) : home + (step >> 1);
    return (n_up > home) ? home + (step - n_both) : home - 1 - (step - n_both);
  };
  static_assert(kTileRows == 128 && kSuper == 16, "mask words / fragment / box staging assume 128-column tiles, 16 per super-tile");
  const float4* box4 = reinterpret_cast<const float4*>(box);
  const float4* sbox4 = reinterpret_cast<const float4*>(sbox);
  // The walk alternates sides of `home` by step parity, so with an even number of warps a fixed residue (step = i *
  // n_warps + warp) would pin every warp to ONE side, half of them starting a super-tile away from home (2 warps per
  // group measured slower than 1).  Rotating the residue with i - warp w takes step i * n_warps + ((w + i) mod
  // n_warps) - lets every warp alternate sides like the single-warp walk; the steps are still covered exactly once.
  auto my_step = [&](long long i) { return i * n_warps + (MULTI ? ((warp + i) & (n_warps - 1)) : 0); };  // n_warps: power of 2
  for (long long i0 = 0; i0 * n_warps < n_super; i0 += 32) {
    {  // the next 32 super-tile boxes of this warp's walk
      const long long st = my_step(i0 + lane);
      float4 lo4 = make_float4(0, 0, 0, 0), hi4 = lo4;
      if (st < n_super) {
        const long long sgl = step_to_super(st);
        lo4 = __ldg(sbox4 + 2 * sgl);
        hi4 = __ldg(sbox4 + 2 * sgl + 1);
      }
      __syncwarp();
      sb_stage[2 * lane] = lo4;
      sb_stage[2 * lane + 1] = hi4;
      __syncwarp();
    }
    for (int si = 0; si < 32; ++si) {
      const long long step = my_step(i0 + si);
      if (step >= n_super) break;
      if (MULTI) {
        const float nb = fminf(second, cta_bound());
        if (__any_sync(0xffffffffu, nb < bound)) {
          bound = nb;
          refresh_thr();
        }
      }
      if (!box_can_matter(sb_stage + 2 * si)) continue;  // 16 tiles at once
      VO_COUNT(0, 1);
      const long long sg = step_to_super(step);
      const long long t0 = sg * kSuper;
      const int n_t = (int)min((long long)kSuper, n_tiles - t0);
      __syncwarp();
      tb_stage[lane] = (lane < 2 * n_t) ? __ldg(box4 + 2 * t0 + lane) : make_float4(0, 0, 0, 0);
      __syncwarp();
      // which of the 16 tiles can matter (current bounds); their fragments come through the TMA engine, the next
      // needed tile's copy in flight while this one is filtered
      unsigned need = 0;
      for (int k = 0; k < n_t; ++k) need |= box_can_matter(tb_stage + 2 * k) ? (1u << k) : 0u;
      if (!need) continue;
      auto fetch = [&](int k, int buf) {
        if (lane == 0) {
          mbar_arrive_expect_tx(&fbar[buf], kFragTileBytes);
          bulk_g2s(fbuf + buf * kFragTileBytes, frag + (size_t)(t0 + k) * (kTileRows * 4), kFragTileBytes, &fbar[buf]);
        }
      };
      int k = __ffs(need) - 1, buf = 0;
      need &= need - 1;
      __syncwarp();  // every lane is done reading the buffers
      fetch(k, 0);
      while (k >= 0) {
        int kn = -1;
        if (kFragBufs > 1 && need) {
          kn = __ffs(need) - 1;
          need &= need - 1;
          fetch(kn, buf ^ 1);
        }
        mbar_wait(&fbar[buf], (fphase >> buf) & 1u);
        fphase ^= 1u << buf;
        const long long t = t0 + k;
        const uint2* fb = reinterpret_cast<const uint2*>(fbuf + buf * kFragTileBytes) + lane;
        const int k_done = k;
        if (kFragBufs > 1) {
          k = kn;
          buf ^= 1;
        } else {
          k = need ? __ffs(need) - 1 : -1;
          need &= need - 1;
        }
        // (re-testing the tile's box against the bounds as they are now skips 0.06% of the tiles: not worth 28
        // instructions per tile)
        (void)k_done;
        VO_COUNT(1, 1);
        // Guard against data the filter cannot thin out (columns packed so tightly that a bf16 bound cannot separate
        // them: every column of a tile survives for some row).  A tile in which one row keeps >= kDenseMinPerRow
        // survivors is cheaper to evaluate outright - two columns per packed instruction, all rows together, the
        // exact twin's arithmetic - than survivor by survivor; and the tiles that follow such a tile skip the filter
        // altogether, re-probing every kDenseSkip tiles.  Exactness is unaffected: evaluating a column that the filter
        // would have excluded can change neither value nor index (update_best_tie is order independent).
        bool dense = VO_MATCH_DENSE_GUARD && dense_skip > 0;
        if (dense) --dense_skip;
        if (!dense)
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {  // 4 column blocks = 32 columns = one mask word
          float run = FLT_MAX;
#pragma unroll
          for (int nb = 4 * gq; nb < 4 * gq + 4; ++nb) {
            float c[4], e[4];
            const uint2 bq = fb[nb * 32];
            mma_bf16_16816(c, afrag[0], bq);
            mma_bf16_16816(e, afrag[1], bq);
            run = min3(min3(c[0], c[1], c[2]), min3(c[3], e[0], e[1]), min3(e[2], e[3], run));
          }
          // some column of the 32 survives for some row: redo the blocks and mark (warp-uniform: mma.sync inside)
          if (__any_sync(0xffffffffu, run <= 0.f)) {
#pragma unroll 1
            for (int nb = 4 * gq; nb < 4 * gq + 4; ++nb) {
              float c[4], e[4];
              const uint2 bq = fb[nb * 32];
              mma_bf16_16816(c, afrag[0], bq);
              mma_bf16_16816(e, afrag[1], bq);
              const int sh = (nb & 3) * 8 + 2 * tq;
              const unsigned m0 = (unsigned)(c[0] <= 0.f) | ((unsigned)(c[1] <= 0.f) << 1);
              const unsigned m1 = (unsigned)(c[2] <= 0.f) | ((unsigned)(c[3] <= 0.f) << 1);
              const unsigned m2 = (unsigned)(e[0] <= 0.f) | ((unsigned)(e[1] <= 0.f) << 1);
              const unsigned m3 = (unsigned)(e[2] <= 0.f) | ((unsigned)(e[3] <= 0.f) << 1);
              if (m0) atomicOr(&mask[(g) * 4 + gq], m0 << sh);
              if (m1) atomicOr(&mask[(g + 8) * 4 + gq], m1 << sh);
              if (m2) atomicOr(&mask[(g + 16) * 4 + gq], m2 << sh);
              if (m3) atomicOr(&mask[(g + 24) * 4 + gq], m3 << sh);
            }
          }
        }
        __syncwarp();
        const uint4 mk = *reinterpret_cast<const uint4*>(mask + lane * 4);
        const bool mine = (mk.x | mk.y | mk.z | mk.w) != 0;
        if (kFragBufs == 1 && k >= 0) fetch(k, 0);  // (the __syncwarp above: all lanes are done with the buffer)
        if (!dense) {
          if (!__any_sync(0xffffffffu, mine)) continue;
          const int kept = __popc(mk.x) + __popc(mk.y) + __popc(mk.z) + __popc(mk.w);
          if (VO_MATCH_DENSE_GUARD && __any_sync(0xffffffffu, kept >= kDenseMinPerRow)) {
            dense = true;
            // the first tiles of every walk are dense by construction (no second-best yet: everything survives);
            // only a REPEAT means the data defeats the filter, and only then are the next tiles taken unfiltered
            if (++dense_streak >= kDenseStreak) dense_skip = kDenseSkip;
            if (mine) *reinterpret_cast<uint4*>(mask + lane * 4) = make_uint4(0, 0, 0, 0);
          } else {
            dense_streak = 0;
          }
        }
        if (dense) {
          VO_COUNT(3, 1);
          const long long j0 = t * kTileRows;
          const int cnt = (int)((n2 - j0 < kTileRows) ? (n2 - j0) : kTileRows);
          Best3 cur = {best, second, bound, idx};
          cur = dense_tile_scan(rec, orig, j0, cnt, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], cur);
          best = cur.best;
          second = cur.second;
          bound = cur.bound;
          idx = cur.idx;
        } else if (mine) {
          *reinterpret_cast<uint4*>(mask + lane * 4) = make_uint4(0, 0, 0, 0);
          auto survivors = [&](unsigned bits, int w) {
            while (bits) {
              const int bpos = __ffs(bits) - 1;
              bits &= bits - 1;
              const long long j = t * kTileRows + w * 32 + bpos;
              const float* rc = rec + (j >> 1) * kPairFloats + (j & 1);
              const int oj = __ldg(orig + j);
              const float d = exact_sqdist10(rc, a);
              if (d <= bound) update_best_tie(d, oj, best, second, idx);
            }
          };
          survivors(mk.x, 0);
          survivors(mk.y, 1);
          survivors(mk.z, 2);
          survivors(mk.w, 3);
          bound = fminf(bound, second);
        }
        __syncwarp();
        if (MULTI) {
          s_wbest[warp][lane] = best;
          if (second < FLT_MAX) atomicMin(&s_bound[lane], __float_as_int(second));  // second >= 0: int order = float order
          bound = fminf(bound, cta_bound());
        }
        refresh_thr();
      }
    }
  }
  if (MULTI) {
    __syncthreads();
    float* m_best = reinterpret_cast<float*>(smem_raw);
    float* m_second = m_best + 32 * n_warps;
    int* m_idx = reinterpret_cast<int*>(m_second + 32 * n_warps);
    m_best[warp * 32 + lane] = best;
    m_second[warp * 32 + lane] = second;
    m_idx[warp * 32 + lane] = idx;
    __syncthreads();
    if (warp != 0) return;
    best = second = FLT_MAX;
    idx = -1;
    float s_min = FLT_MAX;
    for (int w = 0; w < n_warps; ++w) {
      if (m_idx[w * 32 + lane] >= 0) update_best_tie(m_best[w * 32 + lane], m_idx[w * 32 + lane], best, second, idx);
      s_min = fminf(s_min, m_second[w * 32 + lane]);
    }
    second = fminf(second, s_min);
  }
  if (valid) {
    o_best[r] = best;
    o_second[r] = second;
    o_idx[r] = idx;
  }
}

// merge the per-split triples in ascending split order (exact: comparisons only), apply the
// accept test (my_utilities.h:103-105) and count accepted rows per block of 256.
__global__ void __launch_bounds__(256) match_merge_kernel(
    const float* __restrict__ p_best, const float* __restrict__ p_second, const int* __restrict__ p_idx,
    long long rows, int n_splits, float dist_thr, float ratio_thr, float* __restrict__ o_best,
    float* __restrict__ o_second, int* __restrict__ o_idx, unsigned char* __restrict__ flags,
    int* __restrict__ block_counts) {
  const long long r = (long long)blockIdx.x * 256 + threadIdx.x;
  int acc = 0;
  if (r < rows) {
    float best = FLT_MAX, second = FLT_MAX;
    int idx = -1;
    for (int s = 0; s < n_splits; ++s) {
      const float b = p_best[(long long)s * rows + r], s2 = p_second[(long long)s * rows + r];
      const int i = p_idx[(long long)s * rows + r];
      if (i < 0) continue;  // empty split or no distance below FLT_MAX
      // replay "b then s2" through the sequential rule: b carries the index, s2 only a value
      update_best(b, i, best, second, idx);
      if (s2 < second) second = s2;
    }
    acc = (idx != -1) && (best < dist_thr) && (__fdiv_rn(best, second) < ratio_thr);
    if (o_best) o_best[r] = best;
    if (o_second) o_second[r] = second;
    o_idx[r] = idx;
    flags[r] = (unsigned char)acc;
  }
  const int cnt = __syncthreads_count(acc);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = cnt;
}

// exclusive scan of the per-block counts by one CTA (fixed order -> deterministic offsets)
__global__ void __launch_bounds__(1024) match_scan_counts_kernel(int* __restrict__ counts, long long n_blocks,
                                                                 long long* __restrict__ total) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (long long base = 0; base < n_blocks; base += 1024) {
    const long long i = base + threadIdx.x;
    const int v = (i < n_blocks) ? counts[i] : 0;
    long long x = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      long long w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const long long incl = x + (warp ? s_warp[warp - 1] : 0) + s_carry;
    if (i < n_blocks) counts[i] = (int)(incl - v);  // exclusive offset (fits: <= rows < 2^31 per call)
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

__global__ void __launch_bounds__(256) match_scatter_kernel(
    const unsigned char* __restrict__ flags, const int* __restrict__ idx, const int* __restrict__ block_offsets,
    long long rows, long long row_begin, long long capacity, const int* __restrict__ idA,
    const int* __restrict__ idB, int2* __restrict__ pairs, unsigned long long* __restrict__ correct) {
  __shared__ int s_warp[8];
  const long long r = (long long)blockIdx.x * 256 + threadIdx.x;
  const int f = (r < rows) ? flags[r] : 0;
  const unsigned bal = __ballot_sync(0xffffffffu, f);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  int off = block_offsets[blockIdx.x];
  for (int w = 0; w < warp; ++w) off += s_warp[w];
  off += __popc(bal & ((1u << lane) - 1u));
  int ok = 0;
  if (f) {
    const int j = idx[r];
    if (off < capacity) pairs[off] = make_int2((int)(row_begin + r), j);
    if (idA && idB) ok = (idA[row_begin + r] == idB[j]);
  }
  const int c = __syncthreads_count(ok);
  if (threadIdx.x == 0 && c) atomicAdd(correct, (unsigned long long)c);
}

// ---- id_real join for the printed statistics (my_utilities.h:89-91): the reference counts
// equal-id pairs inside the O(N1*N2) loop; the same integer comes out of a hash join.
constexpr unsigned long long kEmptySlot = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ unsigned hash_u32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__global__ void idjoin_build_kernel(const int* __restrict__ idB, long long n2, unsigned long long* table,
                                    unsigned mask) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n2) return;
  const unsigned key = (unsigned)idB[j];
  unsigned h = hash_u32(key) & mask;
  while (true) {
    unsigned long long cur = table[h];
    if (cur == kEmptySlot) {
      const unsigned long long want = (1ull << 32) | key;
      const unsigned long long old = atomicCAS(&table[h], kEmptySlot, want);
      if (old == kEmptySlot) return;
      cur = old;
    }
    if ((unsigned)(cur & 0xFFFFFFFFull) == key) {
      atomicAdd(&table[h], 1ull << 32);
      return;
    }
    h = (h + 1) & mask;
  }
}

__global__ void idjoin_probe_kernel(const int* __restrict__ idA, long long row_begin, long long row_end,
                                    const unsigned long long* __restrict__ table, unsigned mask,
                                    unsigned long long* possible) {
  const long long i = row_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long c = 0;
  if (i < row_end) {
    const unsigned key = (unsigned)idA[i];
    unsigned h = hash_u32(key) & mask;
    while (true) {
      const unsigned long long cur = table[h];
      if (cur == kEmptySlot) break;
      if ((unsigned)(cur & 0xFFFFFFFFull) == key) {
        c = cur >> 32;
        break;
      }
      h = (h + 1) & mask;
    }
  }
  // integer sum: order-independent, so atomics stay deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(possible, c);
}

template <int DIM>
void launch_scan(dim3 grid, cudaStream_t st, const float* A, long long rb, long long re, const float* B,
                 long long n2, long long split, const unsigned* order, float* ob, float* os, int* oi, bool plain) {
  if (DIM == 10 && !plain) match_scan10_kernel<<<grid, kMatchThreads, 0, st>>>(A, rb, re, B, n2, split, order, ob, os, oi);
  else match_scan_kernel<DIM><<<grid, kMatchThreads, 0, st>>>(A, rb, re, B, n2, split, ob, os, oi);
}

typedef void (*scan_fn)(dim3, cudaStream_t, const float*, long long, long long, const float*, long long, long long,
                        const unsigned*, float*, float*, int*, bool);

constexpr long long kSortMinRows = 8192;  // below this the scan is latency-bound and the ordering does not pay
constexpr long long kIndexMinRows = 8192;        // rows for which building the column index always pays (columns >= kSortMinRows)
constexpr long long kIndexMinPairs = 1ll << 28;  // ... and the work above which it pays for fewer rows

scan_fn scan_for_dim(int dim) {
  switch (dim) {
    case 1: return launch_scan<1>;   case 2: return launch_scan<2>;   case 3: return launch_scan<3>;
    case 4: return launch_scan<4>;   case 5: return launch_scan<5>;   case 6: return launch_scan<6>;
    case 7: return launch_scan<7>;   case 8: return launch_scan<8>;   case 9: return launch_scan<9>;
    case 10: return launch_scan<10>; case 11: return launch_scan<11>; case 12: return launch_scan<12>;
    case 13: return launch_scan<13>; case 14: return launch_scan<14>; case 15: return launch_scan<15>;
    case 16: return launch_scan<16>;
    default: return nullptr;
  }
}

}  // namespace

int vo_scan_block_counts(vo_ctx* ctx, int* d_counts, long long n_blocks, long long* d_total) {
  match_scan_counts_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, n_blocks, d_total);
  VO_CHECK_LAUNCH(ctx, "match_scan_counts_kernel");
  return VO_OK;
}

#ifdef VO_MATCH_COUNTERS
extern "C" int vo_debug_match_counters(unsigned long long out[4], int reset) {
  if (cudaMemcpyFromSymbol(out, g_match_counters, 32) != cudaSuccess) return VO_ERR_CUDA;
  if (reset) {
    unsigned long long z[4] = {0, 0, 0, 0};
    if (cudaMemcpyToSymbol(g_match_counters, z, 32) != cudaSuccess) return VO_ERR_CUDA;
  }
  return VO_OK;
}
#endif

// per-row match index for the exchange of the sharded matcher: best column if the row was accepted, else -1
__global__ void __launch_bounds__(256) match_idx_kernel(const unsigned char* __restrict__ flags, const int* __restrict__ idx,
                                                        long long rows, int* __restrict__ out) {
  const long long r = (long long)blockIdx.x * 256 + threadIdx.x;
  if (r < rows) out[r] = flags[r] ? idx[r] : -1;
}

// pairs (i, m[i]) for m[i] >= 0, ascending i: flags + per-block counts, then the shared scan / scatter
__global__ void __launch_bounds__(256) match_idx_flags_kernel(const int* __restrict__ m, long long rows,
                                                              unsigned char* __restrict__ flags, int* __restrict__ block_counts) {
  const long long r = (long long)blockIdx.x * 256 + threadIdx.x;
  const int acc = (r < rows) && (m[r] >= 0);
  if (r < rows) flags[r] = (unsigned char)acc;
  const int cnt = __syncthreads_count(acc);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = cnt;
}

// curve_shard / n_curve_shards: with n_curve_shards > 1 the call covers ALL rows [0, n1) but scans only this
// shard's share of them: the shard-th contiguous segment of the rows' Morton order on the indexed path (every shard
// then sees the row density of the unsharded problem, so the index prunes as well as on one GPU), a contiguous
// index block otherwise.  Rows of other shards come out as "no match".  d_match_idx (nullable): per-row result
// for the exchange between the shards (accepted ? best column : -1).
int match_dev_impl(vo_ctx* ctx, const float* d_descA, int64_t n1, const float* d_descB, int64_t n2, int dim,
                   float dist_thr, float ratio_thr, const int32_t* d_idA, const int32_t* d_idB, int64_t row_begin,
                   int64_t row_end, int32_t* d_pairs_out, int64_t capacity, int64_t* n_out, int64_t stats[2],
                   float* d_best, float* d_second, int32_t* d_best_idx, int curve_shard, int n_curve_shards,
                   int32_t* d_match_idx) {
  if (!ctx) return VO_ERR_INVALID;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  VO_REQUIRE(ctx, n1 >= 0 && n2 >= 0 && n2 < 0x7fffffffLL, "vo_match: sizes");
  VO_REQUIRE(ctx, dim >= 1 && dim <= kMaxDim, "vo_match: descriptor dimension must be in [1,16]");
  VO_REQUIRE(ctx, row_begin >= 0 && row_begin <= row_end && row_end <= n1, "vo_match: row range");
  VO_REQUIRE(ctx, n_out != nullptr && capacity >= 0, "vo_match: outputs");
  const long long rows = row_end - row_begin;
  *n_out = 0;
  if (stats) stats[0] = stats[1] = 0;
  if (rows == 0) return VO_OK;
  VO_REQUIRE(ctx, rows < 0x7fffffffLL, "vo_match: at most 2^31-1 rows per call");
  VO_REQUIRE(ctx, d_descA && (n2 == 0 || d_descB), "vo_match: null descriptors");
  VO_REQUIRE(ctx, capacity == 0 || d_pairs_out, "vo_match: null pairs_out");
  VO_REQUIRE(ctx, n_curve_shards >= 1 && curve_shard >= 0 && curve_shard < n_curve_shards, "vo_match: shard");
  if (n_curve_shards > 1) {
    const bool can_index = (dim == 10) && n2 >= kSortMinRows && ctx->match_path != VO_MATCH_PATH_BRUTE &&
                           ctx->match_path != VO_MATCH_PATH_ORDERED &&
                           (rows >= kIndexMinRows || (rows >= 32 && rows * n2 >= kIndexMinPairs));
    if (!can_index) {  // no Morton order to shard along: this shard's contiguous block of row indices
      const long long lo = row_begin + rows * curve_shard / n_curve_shards;
      const long long hi = row_begin + rows * (curve_shard + 1) / n_curve_shards;
      if (d_match_idx) VO_CUDA(ctx, cudaMemsetAsync(d_match_idx, 0xFF, (size_t)rows * 4, ctx->stream));
      if (hi == lo) return VO_OK;
      return match_dev_impl(ctx, d_descA, n1, d_descB, n2, dim, dist_thr, ratio_thr, d_idA, d_idB, lo, hi, d_pairs_out,
                            capacity, n_out, stats, d_best ? d_best + (lo - row_begin) : nullptr,
                            d_second ? d_second + (lo - row_begin) : nullptr,
                            d_best_idx ? d_best_idx + (lo - row_begin) : nullptr, 0, 1,
                            d_match_idx ? d_match_idx + (lo - row_begin) : nullptr);
    }
  }

  // column splits: fill whole waves (sm_count * resident CTAs) without starving any CTA of work
  const long long row_blocks = (rows + kMatchThreads - 1) / kMatchThreads;
  const long long slots = (long long)ctx->sm_count * 32;
  long long n_splits = 1;
  if (row_blocks < 2 * slots && n2 > 0) {
    n_splits = (2 * slots + row_blocks - 1) / row_blocks;
    const long long max_splits = (n2 + 4 * kTileRows - 1) / (4 * kTileRows);
    if (n_splits > max_splits) n_splits = max_splits;
    if (n_splits < 1) n_splits = 1;
    if (n_splits > 65535) n_splits = 65535;
  }
  long long split_size = n2 > 0 ? (n2 + n_splits - 1) / n_splits : 1;
  split_size = (split_size + kTileRows - 1) / kTileRows * kTileRows;
  n_splits = n2 > 0 ? (n2 + split_size - 1) / split_size : 1;

  const long long merge_blocks = (rows + 255) / 256;
  const bool want_ids = d_idA && d_idB;
  unsigned table_size = 0;
  if (want_ids && n2 > 0) {
    table_size = 64;
    while ((long long)table_size < 2 * n2) table_size <<= 1;
  }
  // scratch carve-up
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off = vo_align_up(off + bytes, 256); return o; };
  const size_t o_pb = carve((size_t)n_splits * rows * 4), o_ps = carve((size_t)n_splits * rows * 4);
  const size_t o_pi = carve((size_t)n_splits * rows * 4), o_idx = carve((size_t)rows * 4);
  const size_t o_flags = carve((size_t)rows), o_counts = carve((size_t)merge_blocks * 4);
  const size_t o_small = carve(64), o_table = carve((size_t)table_size * 8);
  // optional Morton ordering of the query rows (D = 10 pruned scan) and, for large column sets, of the columns
  // too (tile boxes + tile skipping)
  // the column index costs ~0.4 ms per million columns to build: worth it for large row blocks and for small ones
  // against very many columns (measured, rows x 1M columns: 256 rows 0.96 vs 1.37 ms, 4096 rows 0.86 vs 11.4 ms)
  // ctx->match_path (vo_match_set_path, diagnostics / bench): 1 = plain tiled brute force, 2 = Morton-ordered rows +
  // exact packed scan, 3 = indexed walk with the exact fp32 scan of every visited tile, 0 / 4 = automatic
  const int path = ctx->match_path;
  const bool indexed = (dim == 10) && n2 >= kSortMinRows && path != VO_MATCH_PATH_BRUTE && path != VO_MATCH_PATH_ORDERED &&
                       (rows >= kIndexMinRows || (rows >= 32 && rows * n2 >= kIndexMinPairs));
  const bool ordered = indexed || ((dim == 10) && rows >= kSortMinRows && n2 > 0 && path != VO_MATCH_PATH_BRUTE);
  if (indexed) n_splits = 1;
  size_t sort_tmp_bytes = 0;
  if (ordered) {
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp_bytes, (const unsigned*)nullptr, (unsigned*)nullptr,
                                    (const unsigned*)nullptr, (unsigned*)nullptr, (int)rows, 0, 32, ctx->stream);
    if (indexed) {
      size_t b2 = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, b2, (const unsigned*)nullptr, (unsigned*)nullptr, (const unsigned*)nullptr,
                                      (unsigned*)nullptr, (int)n2, 0, 32, ctx->stream);
      if (b2 > sort_tmp_bytes) sort_tmp_bytes = b2;
    }
  }
  const long long n2p = (n2 + 1) & ~1ll, n_tiles = (n2 + kTileRows - 1) / kTileRows;
  const size_t o_keys = carve(ordered ? (size_t)rows * 4 : 0), o_keys2 = carve(ordered ? (size_t)rows * 4 : 0);
  const size_t o_ids = carve(ordered ? (size_t)rows * 4 : 0), o_order = carve(ordered ? (size_t)rows * 4 : 0);
  const size_t o_mm = carve(128), o_sorttmp = carve(sort_tmp_bytes);
  const size_t o_ckeys = carve(indexed ? (size_t)n2 * 4 : 0), o_ckeys2 = carve(indexed ? (size_t)n2 * 4 : 0);
  const size_t o_cids = carve(indexed ? (size_t)n2 * 4 : 0), o_corder = carve(indexed ? (size_t)n2 * 4 : 0);
  const size_t o_rec = carve(indexed ? (size_t)n2p * 10 * 4 : 0), o_orig = carve(indexed ? (size_t)n2p * 4 : 0);
  const size_t o_box = carve(indexed ? (size_t)n_tiles * 32 : 0);
  const long long n_super = (n_tiles + kSuper - 1) / kSuper;
  const size_t o_sbox = carve(indexed ? (size_t)n_super * 32 : 0);
  const size_t o_frag = carve(indexed ? (size_t)n_tiles * kTileRows * 32 : 0);
  char* base;
  st = vo_scratch(ctx, off, (void**)&base);
  if (st) return st;
  float* pb = (float*)(base + o_pb);
  float* ps = (float*)(base + o_ps);
  int* pi = (int*)(base + o_pi);
  int* idx = d_best_idx ? d_best_idx : (int*)(base + o_idx);
  unsigned char* flags = (unsigned char*)(base + o_flags);
  int* counts = (int*)(base + o_counts);
  long long* d_total = (long long*)(base + o_small);
  unsigned long long* d_correct = (unsigned long long*)(base + o_small + 8);
  unsigned long long* d_possible = (unsigned long long*)(base + o_small + 16);
  unsigned long long* table = (unsigned long long*)(base + o_table);
  VO_CUDA(ctx, cudaMemsetAsync(d_total, 0, 64, ctx->stream));

  const unsigned* order = nullptr;
  if (ordered) {
    unsigned* keys = (unsigned*)(base + o_keys);
    unsigned* keys2 = (unsigned*)(base + o_keys2);
    unsigned* ids = (unsigned*)(base + o_ids);
    unsigned* sorted_ids = (unsigned*)(base + o_order);
    unsigned* mm = (unsigned*)(base + o_mm);
    VO_CUDA(ctx, cudaMemsetAsync(mm, 0xFF, 40, ctx->stream));
    VO_CUDA(ctx, cudaMemsetAsync(mm + 10, 0x00, 40, ctx->stream));
    match_minmax10_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d_descA, row_begin, rows, mm);
    VO_CHECK_LAUNCH(ctx, "match_minmax10_kernel");
    if (indexed) {  // one key range for rows and columns
      match_minmax10_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d_descB, 0, n2, mm);
      VO_CHECK_LAUNCH(ctx, "match_minmax10_kernel");
    }
    int* range_flag = (int*)(base + o_small + 32);  // (zeroed with d_total above)
    match_keys10_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(d_descA, row_begin, rows, mm, keys, ids,
                                                                                 indexed ? range_flag : nullptr);
    VO_CHECK_LAUNCH(ctx, "match_keys10_kernel");
    VO_CUDA(ctx, cub::DeviceRadixSort::SortPairs(base + o_sorttmp, sort_tmp_bytes, keys, keys2, ids, sorted_ids,
                                                 (int)rows, 0, 32, ctx->stream));
    ctx->launches += 6;  // CUB's histogram + exclusive sum + four onesweep passes
    order = sorted_ids;
    if (indexed) {
      unsigned* ckeys = (unsigned*)(base + o_ckeys);
      unsigned* ckeys2 = (unsigned*)(base + o_ckeys2);
      unsigned* cids = (unsigned*)(base + o_cids);
      unsigned* corder = (unsigned*)(base + o_corder);
      float* rec = (float*)(base + o_rec);
      int* orig = (int*)(base + o_orig);
      float* box = (float*)(base + o_box);
      match_keys10_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, ctx->stream>>>(d_descB, 0, n2, mm, ckeys, cids, range_flag);
      VO_CHECK_LAUNCH(ctx, "match_keys10_kernel");
      VO_CUDA(ctx, cub::DeviceRadixSort::SortPairs(base + o_sorttmp, sort_tmp_bytes, ckeys, ckeys2, cids, corder, (int)n2, 0,
                                                   32, ctx->stream));
      ctx->launches += 6;
      match_gather10_kernel<<<(unsigned)((n2p + 255) / 256), 256, 0, ctx->stream>>>(d_descB, corder, n2, rec, orig);
      VO_CHECK_LAUNCH(ctx, "match_gather10_kernel");
      match_tilebox10_kernel<<<(unsigned)n_tiles, kTileRows, 0, ctx->stream>>>(rec, n2, box);
      VO_CHECK_LAUNCH(ctx, "match_tilebox10_kernel");
      float* sbox = (float*)(base + o_sbox);
      match_superbox10_kernel<<<(unsigned)((n_super + 127) / 128), 128, 0, ctx->stream>>>(box, n_tiles, sbox);
      VO_CHECK_LAUNCH(ctx, "match_superbox10_kernel");
      // this call's share of the sorted rows (all of them unless the rows are sharded along the curve)
      const long long seg_lo = rows * curve_shard / n_curve_shards, seg_hi = rows * (curve_shard + 1) / n_curve_shards;
      const long long seg_rows = seg_hi - seg_lo;
      if (n_curve_shards > 1) VO_CUDA(ctx, cudaMemsetAsync(pi, 0xFF, (size_t)rows * 4, ctx->stream));  // others: no match
      // warps per 32-row group: enough to give every SM ~32 warps when the rows alone cannot
      const long long groups = (seg_rows + 31) / 32;
      int n_warps = 1;
      // measured (exp/match_nwarps.py, rows x 1M columns, B200; ms with 1 / 2 / 4 / 8 / 16 warps per group):
      //   2048 rows 2.99 1.85 1.12 0.79 0.66 | 16384 rows 2.71 1.81 1.20 1.05 1.05 | 65536 rows 2.97 2.42 2.04 1.97 2.35
      //   131072 rows (one of 8 curve shards) 3.13 2.93 2.70 2.91 | 262144 (one of 4) 4.75 4.83 4.61 5.22
      //   524288 (one of 2) 8.16 8.80 8.59 9.91 | 1048576 14.90 16.71
      static const long long env_budget = getenv("VO_MATCH_WARP_BUDGET") ? atoll(getenv("VO_MATCH_WARP_BUDGET")) : 0;
      if (groups >= 16384 && env_budget <= 0) {
        n_warps = 1;
      } else if (groups >= 4096 && env_budget <= 0) {
        n_warps = 4;
      } else {
        const long long warp_budget = env_budget > 0 ? env_budget : (groups >= 2048 ? 64 : 32);
        while (n_warps < kMaxScanWarps && groups * n_warps * 2 <= (long long)ctx->sm_count * warp_budget) n_warps *= 2;
      }
      if (const char* e = getenv("VO_MATCH_NWARPS")) n_warps = max(1, min(kMaxScanWarps, atoi(e)));  // experiments
      // magnitudes the filter's error analysis does not cover select the exact scan (flag read on the device)
      static const bool env_exact = getenv("VO_MATCH_FORCE_EXACT") != nullptr;  // diagnostics: skip the filter
      const bool force_exact = env_exact || path == VO_MATCH_PATH_INDEXED_EXACT;
      if (force_exact) VO_CUDA(ctx, cudaMemsetAsync(range_flag, 1, 4, ctx->stream));
      uint2* frag = (uint2*)(base + o_frag);
      match_gatherfrag10_kernel<<<(unsigned)((n_tiles * kTileRows + 255) / 256), 256, 0, ctx->stream>>>(
          d_descB, corder, n2, n_tiles * kTileRows, mm, (uint4*)frag);
      VO_CHECK_LAUNCH(ctx, "match_gatherfrag10_kernel");
      const size_t mma_smem = 1024 + (size_t)n_warps * kMmaWarpSmem;
      static_assert(kMmaWarpSmem >= 3 * 32 * 4, "merge arrays reuse the fragment / mask buffers");
      if (mma_smem > 48 * 1024)
        VO_CUDA(ctx, cudaFuncSetAttribute(match_scan10_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          1024 + kMaxScanWarps * kMmaWarpSmem));
      auto mma_scan = n_warps > 1 ? match_scan10_mma_kernel<true> : match_scan10_mma_kernel<false>;
      if (groups > 0)
      mma_scan<<<(unsigned)groups, 32 * n_warps, mma_smem, ctx->stream>>>(
          d_descA, row_begin, seg_rows, sorted_ids + seg_lo, keys2 + seg_lo, rec, orig, frag, box, sbox, ckeys2, n2, mm,
          range_flag, pb, ps, pi);
      VO_CHECK_LAUNCH(ctx, "match_scan10_mma_kernel");
      const size_t scan_smem = (size_t)n_warps * kTileBytes;
      static_assert(kTileBytes >= 3 * 32 * 4, "merge arrays reuse the tile buffers");
      if (scan_smem > 48 * 1024)
        VO_CUDA(ctx, cudaFuncSetAttribute(match_scan10_indexed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          kMaxScanWarps * kTileBytes));
      if (groups > 0)
      match_scan10_indexed_kernel<<<(unsigned)groups, 32 * n_warps, scan_smem, ctx->stream>>>(
          d_descA, row_begin, seg_rows, sorted_ids + seg_lo, keys2 + seg_lo, rec, orig, box, sbox, ckeys2, n2, range_flag,
          pb, ps, pi);
      VO_CHECK_LAUNCH(ctx, "match_scan10_indexed_kernel");
    }
  }
  if (!indexed) {
    scan_fn scan = scan_for_dim(dim);
    dim3 grid((unsigned)row_blocks, (unsigned)n_splits);
    scan(grid, ctx->stream, d_descA, row_begin, row_end, d_descB, n2, split_size, order, pb, ps, pi,
         path == VO_MATCH_PATH_BRUTE);
    VO_CHECK_LAUNCH(ctx, "match_scan_kernel");
  }
  match_merge_kernel<<<(unsigned)merge_blocks, 256, 0, ctx->stream>>>(pb, ps, pi, rows, (int)n_splits, dist_thr,
                                                                      ratio_thr, d_best, d_second, idx, flags,
                                                                      counts);
  VO_CHECK_LAUNCH(ctx, "match_merge_kernel");
  if (d_match_idx) {
    match_idx_kernel<<<(unsigned)merge_blocks, 256, 0, ctx->stream>>>(flags, idx, rows, d_match_idx);
    VO_CHECK_LAUNCH(ctx, "match_idx_kernel");
    // the sharded path wants the per-row result only (it compacts after the exchange): no pair list, no count, and
    // no host synchronisation between the scan and the all-reduce
    if (!d_pairs_out && capacity == 0 && !stats && !want_ids) {
      if (n_out) *n_out = 0;
      return VO_OK;
    }
  }
  st = vo_scan_block_counts(ctx, counts, merge_blocks, d_total);
  if (st) return st;
  match_scatter_kernel<<<(unsigned)merge_blocks, 256, 0, ctx->stream>>>(
      flags, idx, counts, rows, row_begin, capacity, want_ids ? d_idA : nullptr, want_ids ? d_idB : nullptr,
      reinterpret_cast<int2*>(d_pairs_out), d_correct);
  VO_CHECK_LAUNCH(ctx, "match_scatter_kernel");
  if (want_ids && n2 > 0) {
    VO_CUDA(ctx, cudaMemsetAsync(table, 0xFF, (size_t)table_size * 8, ctx->stream));
    idjoin_build_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, ctx->stream>>>(d_idB, n2, table, table_size - 1);
    VO_CHECK_LAUNCH(ctx, "idjoin_build_kernel");
    idjoin_probe_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(d_idA, row_begin, row_end, table,
                                                                               table_size - 1, d_possible);
    VO_CHECK_LAUNCH(ctx, "idjoin_probe_kernel");
  }
  void* h;
  st = vo_pinned(ctx, 64, &h);
  if (st) return st;
  VO_CUDA(ctx, cudaMemcpyAsync(h, d_total, 24, cudaMemcpyDeviceToHost, ctx->stream));
  VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const long long total = ((long long*)h)[0];
  *n_out = total;
  if (stats) {
    stats[1] = (int64_t)((unsigned long long*)h)[1];
    stats[0] = (int64_t)((unsigned long long*)h)[2];
  }
  if (total > capacity) return vo_set_error(ctx, VO_ERR_CAPACITY, "vo_match", "pairs_out capacity");
  return VO_OK;
}

extern "C" {

int vo_match_dev(vo_ctx* ctx, const float* d_descA, int64_t n1, const float* d_descB, int64_t n2, int dim,
                 float dist_thr, float ratio_thr, const int32_t* d_idA, const int32_t* d_idB, int64_t row_begin,
                 int64_t row_end, int32_t* d_pairs_out, int64_t capacity, int64_t* n_out, int64_t stats[2],
                 float* d_best, float* d_second, int32_t* d_best_idx) {
  return match_dev_impl(ctx, d_descA, n1, d_descB, n2, dim, dist_thr, ratio_thr, d_idA, d_idB, row_begin, row_end,
                        d_pairs_out, capacity, n_out, stats, d_best, d_second, d_best_idx, 0, 1, nullptr);
}

int vo_match_compact_dev(vo_ctx* ctx, const int32_t* d_match_idx, int64_t n1, int32_t* d_pairs_out, int64_t capacity,
                         int64_t* n_out) {
  if (!ctx || !n_out) return VO_ERR_INVALID;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  *n_out = 0;
  VO_REQUIRE(ctx, n1 >= 0 && n1 < 0x7fffffffLL && capacity >= 0, "vo_match_compact: sizes");
  if (n1 == 0) return VO_OK;
  VO_REQUIRE(ctx, d_match_idx && (capacity == 0 || d_pairs_out), "vo_match_compact: null buffers");
  const long long blocks = (n1 + 255) / 256;
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off = vo_align_up(off + bytes, 256); return o; };
  const size_t o_flags = carve((size_t)n1), o_counts = carve((size_t)blocks * 4), o_small = carve(64);
  char* base;
  st = vo_scratch(ctx, off, (void**)&base);
  if (st) return st;
  unsigned char* flags = (unsigned char*)(base + o_flags);
  int* counts = (int*)(base + o_counts);
  long long* d_total = (long long*)(base + o_small);
  match_idx_flags_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_match_idx, n1, flags, counts);
  VO_CHECK_LAUNCH(ctx, "match_idx_flags_kernel");
  st = vo_scan_block_counts(ctx, counts, blocks, d_total);
  if (st) return st;
  match_scatter_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(flags, d_match_idx, counts, n1, 0, capacity, nullptr,
                                                                 nullptr, reinterpret_cast<int2*>(d_pairs_out), nullptr);
  VO_CHECK_LAUNCH(ctx, "match_scatter_kernel");
  void* h;
  st = vo_pinned(ctx, 64, &h);
  if (st) return st;
  VO_CUDA(ctx, cudaMemcpyAsync(h, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
  VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *n_out = *(long long*)h;
  if (*n_out > capacity) return vo_set_error(ctx, VO_ERR_CAPACITY, "vo_match_compact", "pairs_out capacity");
  return VO_OK;
}

int vo_match_sharded_dev(vo_ctx* ctx, const float* d_descA, int64_t n1, const float* d_descB, int64_t n2, int dim,
                         float dist_thr, float ratio_thr, int shard, int n_shards, int32_t* d_match_idx,
                         int32_t* d_pairs_out, int64_t capacity, int64_t* n_out) {
  if (!ctx || !n_out) return VO_ERR_INVALID;
  VO_REQUIRE(ctx, d_match_idx != nullptr, "vo_match_sharded: d_match_idx");
  VO_REQUIRE(ctx, n_shards >= 1 && shard >= 0 && shard < n_shards, "vo_match_sharded: shard");
  VO_REQUIRE(ctx, n_shards == 1 || ctx->nccl_comm == nullptr || ctx->n_ranks == n_shards,
             "vo_match_sharded: n_shards must equal the communicator size");
  *n_out = 0;
  if (n1 == 0) return VO_OK;
  int64_t n_mine = 0;
  // (1) this shard's rows; everybody else's come out as -1
  int st = match_dev_impl(ctx, d_descA, n1, d_descB, n2, dim, dist_thr, ratio_thr, nullptr, nullptr, 0, n1, nullptr, 0,
                          &n_mine, nullptr, nullptr, nullptr, nullptr, shard, n_shards, d_match_idx);
  if (st != VO_OK && st != VO_ERR_CAPACITY) return st;  // (capacity 0: only the per-row result is wanted here)
  // (2) the exchange: every row is owned by exactly one shard, so an element-wise MAX merges the shards' results
  if (n_shards > 1 && ctx->nccl_comm) {
    st = vo_comm_allreduce_max_i32(ctx, d_match_idx, n1);
    if (st) return st;
  }
  // (3) the accepted pairs in ascending row order (identical on every rank after the exchange)
  if (capacity > 0 || d_pairs_out) return vo_match_compact_dev(ctx, d_match_idx, n1, d_pairs_out, capacity, n_out);
  return VO_OK;
}

int vo_match_set_path(vo_ctx* ctx, int path) {
  if (!ctx || path < VO_MATCH_PATH_AUTO || path > VO_MATCH_PATH_INDEXED_FILTERED) return VO_ERR_INVALID;
  ctx->match_path = path;
  return VO_OK;
}

int vo_match(vo_ctx* ctx, const float* descA, int64_t n1, const float* descB, int64_t n2, int dim, float dist_thr,
             float ratio_thr, const int32_t* idA, const int32_t* idB, int64_t row_begin, int64_t row_end,
             int32_t* pairs_out, int64_t capacity, int64_t* n_out, int64_t stats[2]) {
  if (!ctx) return VO_ERR_INVALID;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  VO_REQUIRE(ctx, n1 >= 0 && n2 >= 0 && dim >= 1 && dim <= kMaxDim, "vo_match: sizes");
  VO_REQUIRE(ctx, row_begin >= 0 && row_begin <= row_end && row_end <= n1, "vo_match: row range");
  VO_REQUIRE(ctx, n_out != nullptr, "vo_match: n_out");
  const long long rows = row_end - row_begin;
  *n_out = 0;
  if (stats) stats[0] = stats[1] = 0;
  if (rows == 0) return VO_OK;
  VO_REQUIRE(ctx, descA && (n2 == 0 || descB), "vo_match: null descriptors");
  const bool ids = idA && idB;
  // host staging: only the row range of A travels
  const size_t bA = (size_t)rows * dim * 4, bB = (size_t)n2 * dim * 4;
  const size_t bIA = ids ? (size_t)rows * 4 : 0, bIB = ids ? (size_t)n2 * 4 : 0, bP = (size_t)rows * 8;
  float *dA = nullptr, *dB = nullptr;
  int32_t *dIA = nullptr, *dIB = nullptr, *dP = nullptr;
  // one staging arena per context (grown on demand, reused): no cudaMalloc / cudaFree per call
  char* stage;
  {
    const size_t oA = 0, oB = vo_align_up(oA + bA, 256), oIA = vo_align_up(oB + bB, 256), oIB = vo_align_up(oIA + bIA, 256);
    const size_t oP = vo_align_up(oIB + bIB, 256), total = oP + bP + 256;
    st = vo_stage(ctx, total, (void**)&stage);
    if (st) return st;
    dA = (float*)(stage + oA);
    dB = (float*)(stage + oB);
    dIA = ids ? (int32_t*)(stage + oIA) : nullptr;
    dIB = ids ? (int32_t*)(stage + oIB) : nullptr;
    dP = (int32_t*)(stage + oP);
  }
  auto cleanup = [&]() {};
  cudaError_t e;
  e = cudaMemcpyAsync(dA, descA + row_begin * dim, bA, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && bB) e = cudaMemcpyAsync(dB, descB, bB, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && ids) e = cudaMemcpyAsync(dIA, idA + row_begin, bIA, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && ids && bIB) e = cudaMemcpyAsync(dIB, idB, bIB, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    cleanup();
    return vo_set_error(ctx, VO_ERR_CUDA, "vo_match: H2D", cudaGetErrorString(e));
  }
  int64_t n = 0;
  st = vo_match_dev(ctx, dA, rows, dB, n2, dim, dist_thr, ratio_thr, dIA, dIB, 0, rows, dP, rows, &n, stats, nullptr,
                    nullptr, nullptr);
  if (st == VO_OK) {
    *n_out = n;
    if (n > capacity) {
      st = vo_set_error(ctx, VO_ERR_CAPACITY, "vo_match", "pairs_out capacity");
    } else if (n) {
      e = cudaMemcpyAsync(pairs_out, dP, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) st = vo_set_error(ctx, VO_ERR_CUDA, "vo_match: D2H", cudaGetErrorString(e));
      for (int64_t k = 0; k < n; ++k) pairs_out[2 * k] += (int32_t)row_begin;  // back to global row ids
    }
  }
  cudaStreamSynchronize(ctx->stream);
  cleanup();
  return st;
}

}  // extern "C"
