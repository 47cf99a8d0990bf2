// imc_num.h — deck-precision number types (Float16 / Float32 / Float64) for host and device.
//
// The reference stores every quantity in the deck's PRECISION type and Julia rounds to that
// type after EVERY arithmetic operation (Float16 ops are computed in Float32 and rounded back;
// see SURVEY.md §7 "Float16").  Num<P> reproduces exactly that: one IEEE rounding per
// operator, no contraction (the translation units that include this header are compiled with
// -fmad=false / -ffp-contract=off).  Places where the reference leaks into Float64 through a
// literal are written explicitly with .d() / Num<P>::from_d().
//
// Shared by the CUDA kernels (csrc/) and by the CPU oracle (oracle/), which includes it as
// plain C++17.  Checked independently against numpy.float16 in tests/test_num_math.py.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#include <cuda_fp16.h>
#define IMC_HD __host__ __device__ __forceinline__
#define IMC_D __device__ __forceinline__
#else
#define IMC_HD inline
#endif

namespace imc {

// ---------------------------------------------------------------------------------------
// binary16 <-> float/double conversions (round-to-nearest-even, subnormals kept)
// ---------------------------------------------------------------------------------------
// Under nvcc (host and device passes) the cuda_fp16 conversions are used; under plain g++ (oracle) _Float16.
IMC_HD float half_bits_to_float(uint16_t h) {
#if defined(__CUDACC__)
  return __half2float(__ushort_as_half(h));
#else
  _Float16 x;
  memcpy(&x, &h, 2);
  return (float)x;
#endif
}
IMC_HD uint16_t float_to_half_bits(float f) {
#if defined(__CUDACC__)
  return __half_as_ushort(__float2half_rn(f));
#else
  _Float16 x = (_Float16)f;
  uint16_t h;
  memcpy(&h, &x, 2);
  return h;
#endif
}
IMC_HD uint16_t double_to_half_bits(double d) {
#if defined(__CUDACC__)
  return __half_as_ushort(__double2half(d));
#else
  _Float16 x = (_Float16)d;  // direct (single) rounding, like Julia's Float16(::Float64)
  uint16_t h;
  memcpy(&h, &x, 2);
  return h;
#endif
}
IMC_HD float round_to_half(float f) { return half_bits_to_float(float_to_half_bits(f)); }
IMC_HD float round_d_to_half(double d) { return half_bits_to_float(double_to_half_bits(d)); }

// ---------------------------------------------------------------------------------------
// precision tags
// ---------------------------------------------------------------------------------------
struct F16 {
  using comp_t = float;      // register type (always holds a binary16-representable value)
  using store_t = uint16_t;  // HBM storage type
  static constexpr int id = 0;
  static constexpr int bytes = 2;
  static IMC_HD comp_t rnd(float x) { return round_to_half(x); }
  static IMC_HD comp_t from_d(double x) { return round_d_to_half(x); }
  static IMC_HD store_t pack(comp_t x) { return float_to_half_bits(x); }
  static IMC_HD comp_t unpack(store_t s) { return half_bits_to_float(s); }
};
struct F32 {
  using comp_t = float;
  using store_t = float;
  static constexpr int id = 1;
  static constexpr int bytes = 4;
  static IMC_HD comp_t rnd(float x) { return x; }
  static IMC_HD comp_t from_d(double x) { return (float)x; }
  static IMC_HD store_t pack(comp_t x) { return x; }
  static IMC_HD comp_t unpack(store_t s) { return s; }
};
struct F64 {
  using comp_t = double;
  using store_t = double;
  static constexpr int id = 2;
  static constexpr int bytes = 8;
  static IMC_HD comp_t rnd(double x) { return x; }
  static IMC_HD comp_t from_d(double x) { return x; }
  static IMC_HD store_t pack(comp_t x) { return x; }
  static IMC_HD comp_t unpack(store_t s) { return s; }
};

// ---------------------------------------------------------------------------------------
// Num<P>: value of deck precision P with Julia's per-operation rounding
// ---------------------------------------------------------------------------------------
template <class P>
struct Num {
  using C = typename P::comp_t;
  C v;
  IMC_HD Num() : v(0) {}
  IMC_HD explicit Num(C x) : v(x) {}  // caller guarantees x is P-representable
  static IMC_HD Num from_d(double x) { return Num(P::from_d(x)); }     // T(x::Float64)
  static IMC_HD Num from_i(long long i) { return Num(P::from_d((double)i)); }  // T(i::Int) (|i| < 2^53)
  static IMC_HD Num load(const typename P::store_t* p, size_t i) { return Num(P::unpack(p[i])); }
  IMC_HD void store(typename P::store_t* p, size_t i) const { p[i] = P::pack(v); }
  IMC_HD double d() const { return (double)v; }  // Float64(x) — exact
  IMC_HD Num operator-() const { return Num(-v); }
  friend IMC_HD Num operator+(Num a, Num b) { return Num(P::rnd(a.v + b.v)); }
  friend IMC_HD Num operator-(Num a, Num b) { return Num(P::rnd(a.v - b.v)); }
  friend IMC_HD Num operator*(Num a, Num b) { return Num(P::rnd(a.v * b.v)); }
  friend IMC_HD Num operator/(Num a, Num b) { return Num(P::rnd(a.v / b.v)); }
  IMC_HD Num& operator+=(Num b) { *this = *this + b; return *this; }
  IMC_HD Num& operator-=(Num b) { *this = *this - b; return *this; }
  IMC_HD Num& operator*=(Num b) { *this = *this * b; return *this; }
  friend IMC_HD bool operator<(Num a, Num b) { return a.v < b.v; }
  friend IMC_HD bool operator>(Num a, Num b) { return a.v > b.v; }
  friend IMC_HD bool operator<=(Num a, Num b) { return a.v <= b.v; }
  friend IMC_HD bool operator>=(Num a, Num b) { return a.v >= b.v; }
  friend IMC_HD bool operator==(Num a, Num b) { return a.v == b.v; }
  friend IMC_HD bool operator!=(Num a, Num b) { return a.v != b.v; }
};

template <class P> IMC_HD bool is_nan(Num<P> a) { return a.v != a.v; }
template <class P> IMC_HD bool is_inf(Num<P> a) { return a.v == a.v && a.v - a.v != a.v - a.v; }
// Julia's abs: abs(-0.0) = 0.0, NaN stays NaN — on the device one |x| operand modifier
template <class P> IMC_HD Num<P> nabs(Num<P> a) {
#if defined(__CUDA_ARCH__)
  if constexpr (P::id == 2) return Num<P>(fabs(a.v)); else return Num<P>(fabsf(a.v));
#else
  return Num<P>(a.v < 0 ? -a.v : (a.v == 0 ? (typename P::comp_t)0 : a.v));
#endif
}

IMC_HD bool sign_bit(double x) {
#if defined(__CUDA_ARCH__)
  return (__double_as_longlong(x) < 0);
#else
  uint64_t u; memcpy(&u, &x, 8); return (u >> 63) != 0;
#endif
}
// Julia's min(x, y) for floats: NaN-propagating; min(-0.0, 0.0) = -0.0.
// Device, Float16/Float32: one FMNMX.NAN — PTX min.NaN.f32 returns NaN if either input is NaN and orders
// -0.0 < +0.0 (PTX ISA "min": "if both inputs are 0.0 then +0.0 > -0.0"); checked on B200 by tests/test_gpu_parity.py.
// the generic definition (host, and the reference the device fast paths are self-tested against)
template <class P> IMC_HD Num<P> jl_min_generic(Num<P> a, Num<P> b) {
  if (a.v != a.v) return a;
  if (b.v != b.v) return b;
  if (a.v < b.v) return a;
  if (b.v < a.v) return b;
  return Num<P>(sign_bit((double)a.v) ? a.v : b.v);
}
template <class P> IMC_HD Num<P> jl_min(Num<P> a, Num<P> b) {
#if defined(__CUDA_ARCH__)
  if constexpr (P::id != 2) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a.v), "f"(b.v)); return Num<P>(r); }
  else {
    // Float64: PTX has no min.NaN.f64.  min.f64 (DMNMX) returns the other operand when one is NaN and orders -0 < +0;
    // the NaN is put back by two selects (a's NaN first, like the generic code).  Checked against jl_min_generic on the
    // device over special values and random pairs (imc_cuda_selftest_min, tests/test_gpu_parity.py).
    const double r = fmin(a.v, b.v);
    return Num<P>(a.v != a.v ? a.v : (b.v != b.v ? b.v : r));
  }
#endif
  if (a.v != a.v) return a;
  if (b.v != b.v) return b;
  if (a.v < b.v) return a;
  if (b.v < a.v) return b;
  // equal (or +-0): prefer the one with the sign bit set
  return Num<P>(sign_bit((double)a.v) ? a.v : b.v);
}
// the smaller of two non-negative values, or the one that is not NaN when exactly one is (the NaN guards of
// imc_transport.jl:551-557 around min(dist_bx, dist_by)); both NaN -> NaN
template <class P> IMC_HD Num<P> min_nonnan_generic(Num<P> a, Num<P> b) { return a.v != a.v ? b : (b.v != b.v ? a : jl_min_generic(a, b)); }
template <class P> IMC_HD Num<P> min_nonnan(Num<P> a, Num<P> b) {
#if defined(__CUDA_ARCH__)
  if constexpr (P::id != 2) return Num<P>(fminf(a.v, b.v)); else return Num<P>(fmin(a.v, b.v));
#endif
  return a.v != a.v ? b : (b.v != b.v ? a : jl_min(a, b));
}
template <class P> IMC_HD Num<P> jl_max(Num<P> a, Num<P> b) {
  if (a.v != a.v) return a;
  if (b.v != b.v) return b;
  if (a.v > b.v) return a;
  if (b.v > a.v) return b;
  return Num<P>(sign_bit((double)a.v) ? b.v : a.v);
}

// round(x) — Julia's default RoundNearest = ties to even.
IMC_HD double round_half_even(double x) {
#if defined(__CUDA_ARCH__)
  return rint(x);
#else
  return __builtin_rint(x);  // default FP environment: to nearest even
#endif
}
template <class P> IMC_HD Num<P> jl_round(Num<P> a) { return Num<P>((typename P::comp_t)round_half_even((double)a.v)); }

}  // namespace imc
