// warp_reduce_host.cpp — csrc/imc_warp_reduce.cuh compiled for the host on top of warp_emu.h (TEST INFRASTRUCTURE).
//   g++ -std=c++20 -O1 -fPIC -shared -ffp-contract=off -mf16c -pthread -I<csrc> -o libwarp_reduce_host.so warp_reduce_host.cpp
#include "warp_emu.h"
#include "imc_warp_reduce.cuh"

using namespace imc;

template <class P>
static double run(int which, double v0, const unsigned* keys, const double* vals, long long n) {
  using N = Num<P>;
  return warp_emu::run_warp<double>([&](int lane) {
    N v = N::from_d(v0);
    if (which == 0) v = warp_seq_add<P>(v, true, keys, vals, 0, n - 1, lane);
    else v = warp_seq_add_skip<P>(v, keys, vals, 0, n - 1, lane);
    return v.d();
  });
}

extern "C" {
// which: 0 = warp_seq_add (the chain as the kernels have always run it), 1 = warp_seq_add_skip; prec: 0 F16, 1 F32, 2 F64
double warp_reduce_host(int which, int prec, double v0, const unsigned* keys, const double* vals, long long n) {
  if (n <= 0) return v0;
  if (prec == 0) return run<F16>(which, v0, keys, vals, n);
  if (prec == 1) return run<F32>(which, v0, keys, vals, n);
  return run<F64>(which, v0, keys, vals, n);
}
// the plain loop both must reproduce: v += record, in order (wide records are added in Float64 and rounded)
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
}
