// warp_reduce_host.cpp — csrc/imc_warp_reduce.cuh compiled for the host on top of warp_emu.h (TEST INFRASTRUCTURE).
//   g++ -std=c++20 -O1 -fPIC -shared -ffp-contract=off -mf16c -pthread -I<csrc> -o libwarp_reduce_host.so warp_reduce_host.cpp
#include "warp_emu.h"
#include "imc_warp_reduce.cuh"
#include "imc_warp_runs.cuh"
#include <mutex>

using namespace imc;

template <class P>
static double run(int which, double v0, const unsigned* keys, const double* vals, long long n) {
  using N = Num<P>;
  return warp_emu::run_warp<double>([&](int lane) {
    N v = N::from_d(v0);
    if (which == 0) v = warp_seq_add<P>(v, true, keys, vals, 0, n - 1, lane);
    else if (which == 1) v = warp_seq_add_skip<P>(v, keys, vals, 0, n - 1, lane);
    else v = warp_jl_sum<P>(vals, 0, n - 1, lane);          // Julia's pairwise sum(vals); v0 and keys unused
    return v.d();
  });
}

// Base.sum over a vector (mapreduce_impl: sequential below 1024 elements, else split in halves), written as the recursion it is
template <class P>
static Num<P> jl_rec(const double* a, long long first, long long last) {
  using N = Num<P>;
  if (first == last) return N::from_d(a[first]);
  if (last - first < 1024) {
    N v = N::from_d(a[first]) + N::from_d(a[first + 1]);
    for (long long i = first + 2; i <= last; ++i) v = v + N::from_d(a[i]);
    return v;
  }
  const long long mid = first + ((last - first) >> 1);
  return jl_rec<P>(a, first, mid) + jl_rec<P>(a, mid + 1, last);
}

// imc_warp_runs.cuh as the census tally uses it: `per` consecutive particles per lane (cell < 0: a dead particle), every
// deposit added to sums[cell]; returns the number of deposits (one per run of the list is the point of the exercise)
template <class V>
static long long runs_host(const int* cells, const double* vals, int per, double* sums) {
  std::mutex mu;
  long long deposits = 0;
  warp_emu::run_warp<int>([&](int lane) {
    ThreadRuns<V> r;
    auto deposit = [&](int c, V x) { std::lock_guard<std::mutex> g(mu); sums[c] += (double)x; ++deposits; };
    for (int j = 0; j < per; ++j) {
      const int c = cells[lane * per + j];
      if (c >= 0) r.push(c, (V)vals[lane * per + j], deposit);
    }
    bool want, want_f;
    warp_join_runs(lane, r, want, want_f);
    if (want) deposit(r.cell, r.v);
    if (want_f) deposit(r.cell_f, r.v_f);
    return 0;
  });
  return deposits;
}

extern "C" {
// integer = 0: double partial sums (float tallies), 1: 64-bit integers (FIXED tallies); vals are whole numbers, so every order of
// additions gives the same sums
long long warp_runs_host(int integer, const int* cells, const double* vals, int per, double* sums) {
  return integer ? runs_host<long long>(cells, vals, per, sums) : runs_host<double>(cells, vals, per, sums);
}
// which: 0 = warp_seq_add (the chain as the kernels have always run it), 1 = warp_seq_add_skip, 2 = warp_jl_sum;
// prec: 0 F16, 1 F32, 2 F64
double warp_reduce_host(int which, int prec, double v0, const unsigned* keys, const double* vals, long long n) {
  if (n <= 0) return v0;
  if (prec == 0) return run<F16>(which, v0, keys, vals, n);
  if (prec == 1) return run<F32>(which, v0, keys, vals, n);
  return run<F64>(which, v0, keys, vals, n);
}
// the plain loop the two chains must reproduce: v += record, in order (wide records are added in Float64 and rounded)
double plain_chain(int prec, double v0, const unsigned* keys, const double* vals, long long n) {
  auto go = [&](auto tag) {
    using P = decltype(tag); using N = Num<P>;
    N v = N::from_d(v0);
    for (long long i = 0; i < n; ++i) v = (keys && (keys[i] & 0x80000000u)) ? N::from_d(v.d() + vals[i]) : v + N::from_d(vals[i]);
    return v.d();
  };
  if (prec == 0) return go(F16{});
  if (prec == 1) return go(F32{});
  return go(F64{});
}
// the plain recursion warp_jl_sum must reproduce
double plain_jl_sum(int prec, const double* vals, long long n) {
  if (prec == 0) return jl_rec<F16>(vals, 0, n - 1).d();
  if (prec == 1) return jl_rec<F32>(vals, 0, n - 1).d();
  return jl_rec<F64>(vals, 0, n - 1).d();
}
}
