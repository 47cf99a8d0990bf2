// imc_device.cuh — device-side building blocks shared by the engine kernels:
//   sorter_dev    Utilities.sorter (imc_utilities.jl:23-54) for one cell
//   jl_sum        Julia Base.sum (pairwise, 1024-element sequential leaves) as a parallel reduction
//                 with exactly the reference's association order — the source counts depend on the
//                 last bit of mesh.totalenergy (imc_sourcing.jl:121, :139), so the order matters
//   exclusive scan of per-entry counts (int32 -> int64)
//   warp / block reductions for counters
#pragma once
#include <cuda_runtime.h>
#include "imc_num.h"
#include "imc_math.h"
#include "imc_rng.h"
#include "imc_sortnet.h"

namespace imc {

#define IMC_FULL_MASK 0xffffffffu

// ---- Utilities.sorter -------------------------------------------------------------------------
// vals: Float64 images of the N entries of the literal array; scales: descending.  Pair products are
// formed in the array's element type and converted to T: for T-valued inputs that is one rounding of
// the exact product, for Float64 inputs (Q12/Q31) T(Float64 product) — both equal from_d(a*b).
// The N values are sorted once by a register sorting network (imc_sortnet.h); each candidate scale is then put in its
// place (after every value <= it, where the reference's sort of `vals U {scale}` leaves it) with static indices.
template <class P, int N>
__device__ __forceinline__ void sorter_dev(const double (&vals)[N], const double* scales, int n_scales, Num<P>* prod_out, int* idx_out) {
  constexpr int M = N + 1;
  double v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = vals[i];
  sortnet::sort(v);
  for (int j = 0; j < n_scales; ++j) {
    const double sc = scales[j];
    int pos = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) pos += v[i] <= sc ? 1 : 0;
    double s[M];
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = i < pos ? v[i < N ? i : N - 1] : (i == pos ? sc : v[i > 0 ? i - 1 : 0]);
    Num<P> product = Num<P>::from_d(1.0);
#pragma unroll
    for (int i = 0; i < M / 2; ++i) product *= Num<P>::from_d(s[i] * s[M - 1 - i]);
    if (M & 1) product *= Num<P>::from_d(s[M / 2]);
    if (!is_inf(product) && !is_nan(product)) { *prod_out = product; *idx_out = j; return; }
  }
  *prod_out = Num<P>();
  *idx_out = -1;
}

// ---- reductions ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(IMC_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(IMC_FULL_MASK, v, o);
  return v;
}

// ---- Julia Base.sum ----------------------------------------------------------------------------
// Recursion: a range with (last - first) < 1024 is summed left to right; otherwise it is split at
// mid = first + ((last - first) >> 1) and the halves are added (base/reduce.jl mapreduce_impl).
// Leaves therefore sit on at most two adjacent depths.  Kernel 1: one thread per slot of the deepest
// level walks down from the root; the thread standing on the leftmost slot of a leaf sums it.
// Kernel 2 (one block) folds the levels bottom-up, left + right, as the recursion would.
inline int jl_sum_depth(long long n) {
  int d = 0;
  while (n > 1024) { n = (n + 1) / 2; ++d; }
  return d;
}

template <class P>
__global__ void k_jlsum_leaves(const typename P::store_t* __restrict__ q, long long n, int depth,
                               typename P::comp_t* __restrict__ part, unsigned char* __restrict__ valid) {
  long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (1ll << depth)) return;
  if (n <= 0) { if (tid == 0) { part[0] = 0; valid[0] = 1; } return; }
  long long first = 0, last = n - 1;
  int d = 0;
  while (true) {
    if (last - first < 1024) {
      int rem = depth - d;
      if (rem > 0 && (tid & ((1ll << rem) - 1)) != 0) { valid[tid] = 0; return; }
      Num<P> v = Num<P>::load(q, first);
      for (long long i = first + 1; i <= last; ++i) v = v + Num<P>::load(q, i);
      part[tid] = v.v;
      valid[tid] = 1;
      return;
    }
    long long mid = first + ((last - first) >> 1);
    int bit = (int)((tid >> (depth - 1 - d)) & 1);
    if (bit) first = mid + 1; else last = mid;
    ++d;
  }
}

template <class P>
__global__ void k_jlsum_fold(typename P::comp_t* part, const unsigned char* valid, int depth,
                             typename P::comp_t* out) {
  for (int level = depth; level >= 1; --level) {
    long long nodes = 1ll << (level - 1);
    int sh = depth - level;  // slot stride of this level's nodes is 1 << sh
    for (long long i = threadIdx.x; i < nodes; i += blockDim.x) {
      long long ls = (2 * i) << sh, rs = (2 * i + 1) << sh;
      if (valid[rs]) part[ls] = (Num<P>(part[ls]) + Num<P>(part[rs])).v;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = part[0];
}

// Several Julia sums in one launch pair (a time step needs 4-15 of them: seven per sourcing call, the energydep planes,
// three more in the tally; at deck sizes that fit one SM the launches, not the additions, are their cost).  blockIdx.y
// selects the sum; each sum has its own depth and its own [slots_max] stretch of the part / valid scratch.
constexpr int JLSUM_BATCH = 8;
template <class P>
struct JlSumBatch {
  const typename P::store_t* q[JLSUM_BATCH];
  long long n[JLSUM_BATCH];
  int depth[JLSUM_BATCH];
  typename P::comp_t* out[JLSUM_BATCH];
  int count;
  long long slots_max;
};
template <class P>
__global__ void k_jlsum_leaves_multi(JlSumBatch<P> b, typename P::comp_t* __restrict__ part_all, unsigned char* __restrict__ valid_all) {
  const int s = blockIdx.y;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int depth = b.depth[s];
  if (tid >= (1ll << depth)) return;
  const typename P::store_t* __restrict__ q = b.q[s];
  typename P::comp_t* part = part_all + (size_t)s * b.slots_max;
  unsigned char* valid = valid_all + (size_t)s * b.slots_max;
  const long long n = b.n[s];
  if (n <= 0) { if (tid == 0) { part[0] = 0; valid[0] = 1; } return; }
  long long first = 0, last = n - 1;
  int d = 0;
  while (true) {
    if (last - first < 1024) {
      int rem = depth - d;
      if (rem > 0 && (tid & ((1ll << rem) - 1)) != 0) { valid[tid] = 0; return; }
      Num<P> v = Num<P>::load(q, first);
      for (long long i = first + 1; i <= last; ++i) v = v + Num<P>::load(q, i);
      part[tid] = v.v;
      valid[tid] = 1;
      return;
    }
    long long mid = first + ((last - first) >> 1);
    int bit = (int)((tid >> (depth - 1 - d)) & 1);
    if (bit) first = mid + 1; else last = mid;
    ++d;
  }
}
template <class P>
__global__ void k_jlsum_fold_multi(JlSumBatch<P> b, typename P::comp_t* part_all, const unsigned char* valid_all) {
  const int s = blockIdx.x;
  typename P::comp_t* part = part_all + (size_t)s * b.slots_max;
  const unsigned char* valid = valid_all + (size_t)s * b.slots_max;
  const int depth = b.depth[s];
  for (int level = depth; level >= 1; --level) {
    long long nodes = 1ll << (level - 1);
    int sh = depth - level;
    for (long long i = threadIdx.x; i < nodes; i += blockDim.x) {
      long long ls = (2 * i) << sh, rs = (2 * i + 1) << sh;
      if (valid[rs]) part[ls] = (Num<P>(part[ls]) + Num<P>(part[rs])).v;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *b.out[s] = part[0];
}

// ---- exclusive scan int32 -> int64 (three-phase) ----------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <class TIn>
__global__ void k_scan_tiles(const TIn* __restrict__ in, long long* __restrict__ out, long long n,
                             long long* __restrict__ tile_sums) {
  __shared__ long long warp_tot[SCAN_THREADS / 32];
  long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  long long v[SCAN_ITEMS];
  long long local = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    long long idx = base + k;
    v[k] = idx < n ? (long long)in[idx] : 0;
    local += v[k];
  }
  // inclusive warp scan of thread totals
  long long x = local;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long y = __shfl_up_sync(IMC_FULL_MASK, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[wid] = x;
  __syncthreads();
  long long woff = 0;
  for (int w = 0; w < wid; ++w) woff += warp_tot[w];
  long long excl = woff + x - local;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    long long idx = base + k;
    if (idx < n) out[idx] = excl;
    excl += v[k];
  }
  if (threadIdx.x == SCAN_THREADS - 1) tile_sums[blockIdx.x] = woff + x;
}
static __global__ void k_scan_add(long long* __restrict__ out, long long n, const long long* __restrict__ tile_off) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) out[idx] += tile_off[idx / SCAN_TILE];
}
// single-block exclusive scan of a (small) long long array, in place; total written to *total
static __global__ void k_scan_small(long long* a, long long n, long long* total) {
  __shared__ long long carry;
  __shared__ long long warp_tot[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (long long base = 0; base < n; base += blockDim.x) {
    long long idx = base + threadIdx.x;
    long long v = idx < n ? a[idx] : 0;
    long long x = v;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long y = __shfl_up_sync(IMC_FULL_MASK, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    long long woff = 0;
    for (int w = 0; w < wid; ++w) woff += warp_tot[w];
    long long c = carry;
    if (idx < n) a[idx] = c + woff + x - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = c + woff + x;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

}  // namespace imc
