// imc_warp_reduce.cuh — a warp adds a run of deposit records to a value IN THE REFERENCE'S ORDER (EXACT tallies:
// `mesh.energydep[c, k] += v`, imc_transport.jl:120; the leaves of Julia's pairwise sum, :202).
//
// Kept in a header of its own, depending only on Num<P> and three warp intrinsics, so that tests/warp_emu can compile
// these functions for the host with the intrinsics emulated by 32 threads in lockstep and check them against a plain
// loop (tests/test_warp_emu.py) — the CPU container has no GPU.
#pragma once
#include "imc_num.h"
#ifndef IMC_FULL_MASK
#define IMC_FULL_MASK 0xffffffffu
#endif

namespace imc {

// v (op)= vals[first..last] in order; `started` = v already holds a value (else v = vals[first] first).  wide records (bit 31 of
// the key, MC_RW) are Float64 values added in Float64 and rounded; keys == nullptr: no wide records (census, pairwise leaves)
template <class P>
__device__ __forceinline__ Num<P> warp_seq_add(Num<P> v, bool started, const unsigned* __restrict__ keys, const double* __restrict__ vals,
                                               long long first, long long last, int lane) {
  using N = Num<P>;
  for (long long base = first; base <= last; base += 32) {
    const long long i = base + lane;
    double x = 0.0; unsigned k = 0u;
    if (i <= last) { x = vals[i]; if (keys) k = keys[i]; }
    const int cnt = (int)(last - base + 1 < 32 ? last - base + 1 : 32);
#pragma unroll 8
    for (int j = 0; j < cnt; ++j) {
      const double xj = __shfl_sync(IMC_FULL_MASK, x, j);
      const unsigned kj = keys ? __shfl_sync(IMC_FULL_MASK, k, j) : 0u;
      if (!started) { v = (kj & 0x80000000u) ? N::from_d(N().d() + xj) : N() + N::from_d(xj); started = true; }
      else v = (kj & 0x80000000u) ? N::from_d(v.d() + xj) : v + N::from_d(xj);
    }
  }
  return v;
}
// The same chain with its stagnant stretches skipped (a Float16 sum stops moving once it dwarfs the deposits: 10^6 records
// per source cell of the Su-Olson deck, a few thousand of which change the sum).  Every lane adds ITS OWN record to the
// running value; if no lane's result differs from that value, the sequential chain over the block leaves it unchanged too
// (each addition would see the same value), so the block is skipped; otherwise the chain jumps to the first record that
// does change it and the remaining lanes are tested against the new value.  Same bits as warp_seq_add for any input
// (scratch/f16_chain_shortcut.py checks the idea on the CPU).  EXPERIMENTAL: selected by IMC_EXACT_SKIP=1, off by default
// until it has been run and timed on a GPU.
template <class P>
__device__ __forceinline__ Num<P> warp_seq_add_skip(Num<P> v, const unsigned* __restrict__ keys, const double* __restrict__ vals,
                                                    long long first, long long last, int lane) {
  using N = Num<P>;
  for (long long base = first; base <= last; base += 32) {
    const long long i = base + lane;
    const bool have = i <= last;
    double x = 0.0; bool wide = false;
    if (have) { x = vals[i]; wide = keys && (keys[i] & 0x80000000u); }
    int done = 0;                                        // records base .. base + done - 1 are accounted for
    while (true) {
      const N t = wide ? N::from_d(v.d() + x) : v + N::from_d(x);
      const bool same = (t.v == v.v && signbit(t.v) == signbit(v.v)) || (t.v != t.v && v.v != v.v);
      const unsigned m = __ballot_sync(IMC_FULL_MASK, have && lane >= done && !same);
      if (m == 0u) break;                                // nothing left in this block moves the sum
      const int j0 = __ffs(m) - 1;
      v = N(__shfl_sync(IMC_FULL_MASK, t.v, j0));        // records done .. j0 - 1 leave v as it is, record j0 gives lane j0's t
      done = j0 + 1;
    }
  }
  return v;
}

// Julia Base.sum of vals[first..last] (jl_sum_serial's recursion, leaves summed by warp_seq_add); all lanes return the value
template <class P>
__device__ inline Num<P> warp_jl_sum(const double* __restrict__ vals, long long first, long long last, int lane) {
  using N = Num<P>;
  struct Frame { long long first, last; int state; N v1; };
  Frame st[48];
  int sp = 0;
  st[sp++] = {first, last, 0, N()};
  N ret;
  while (sp > 0) {
    Frame& f = st[sp - 1];
    if (f.state == 0) {
      if (f.last - f.first < 1024) {   // one element, or a sequential leaf: v = A[first] + A[first+1]; v += A[i] ...
        const double x0 = vals[f.first];
        ret = f.first == f.last ? N::from_d(x0) : warp_seq_add<P>(N::from_d(x0), true, nullptr, vals, f.first + 1, f.last, lane);
        --sp;
      } else {
        long long mid = f.first + ((f.last - f.first) >> 1);
        f.state = 1;
        st[sp++] = {f.first, mid, 0, N()};
      }
    } else if (f.state == 1) {
      f.v1 = ret; f.state = 2;
      long long mid = f.first + ((f.last - f.first) >> 1);
      st[sp++] = {mid + 1, f.last, 0, N()};
    } else { ret = f.v1 + ret; --sp; }
  }
  return ret;
}

}  // namespace imc
