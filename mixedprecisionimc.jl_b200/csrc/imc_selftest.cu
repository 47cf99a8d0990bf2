// imc_selftest.cu — device self-tests reachable from the test-suite (not part of include/imc.h).
//
// imc_cuda_selftest_div: FastDivisor::divide (imc_fastdiv.cuh) against the IEEE division `a / b` on the device,
// over random operand pairs and over pairs constructed next to rounding boundaries.
#include <cuda_runtime.h>
#include <stdint.h>
#include "imc_fastdiv.cuh"
#include "imc_num.h"

namespace imc {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float make_float(uint32_t mant, int exp2, bool neg) {   // 1.mant * 2^exp2
  return __uint_as_float((neg ? 0x80000000u : 0u) | ((uint32_t)(exp2 + 127) << 23) | (mant & 0x7fffffu));
}

__global__ void k_selftest_div(uint64_t seed, long long per_thread, int mode, unsigned long long* mismatches, unsigned long long* tested,
                               float* first_bad) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t s = mix64(seed ^ (tid * 0x2545F4914F6CDD1Dull));
  unsigned long long bad = 0, n = 0;
  for (long long it = 0; it < per_thread; ++it) {
    s = mix64(s);
    const uint64_t u0 = s; s = mix64(s); const uint64_t u1 = s;
    float a, b;
    if (mode == 0) {            // random mantissas, exponents over the whole fast range and a little beyond
      b = make_float((uint32_t)u0, (int)((u0 >> 32) % 85) - 42, (u0 >> 60) & 1);
      a = make_float((uint32_t)u1, (int)((u1 >> 32) % 85) - 42, (u1 >> 60) & 1);
    } else if (mode == 1) {     // the tracking shapes: b a direction cosine in (0, 1], a a length in (0, 4)
      b = make_float((uint32_t)u0, -(int)((u0 >> 32) % 30), (u0 >> 60) & 1);
      a = make_float((uint32_t)u1, 1 - (int)((u1 >> 32) % 34), false);
    } else {                    // quotient next to a rounding boundary: a = RN(b * (q + (k/8) ulp(q))), k in -8..8
      b = make_float((uint32_t)u0, (int)((u0 >> 32) % 41) - 20, (u0 >> 60) & 1);
      const float q = make_float((uint32_t)u1, (int)((u1 >> 32) % 41) - 20, false);
      const int k = (int)((u1 >> 40) % 17) - 8;
      const double qd = (double)q + (double)k * 0.125 * (double)(__uint_as_float(__float_as_uint(q) + 1u) - q);
      a = (float)((double)b * qd);
      if (mode == 3) a = __uint_as_float(__float_as_uint(a) + (uint32_t)((u1 >> 50) % 5) - 2u);   // and its neighbours
    }
    FastDivisor d, d2; d.set(b); d2.set_rn(b);
    const float got = d.divide(a), got2 = d2.divide(a), want = a / b;
    ++n;
    if ((__float_as_uint(got) != __float_as_uint(want) || __float_as_uint(got2) != __float_as_uint(want)) && !(got != got && want != want)) {
      if (atomicAdd(mismatches, 1ull) == 0ull) { first_bad[0] = a; first_bad[1] = b; first_bad[2] = got; first_bad[3] = want; }
      ++bad;
    }
  }
  atomicAdd(tested, n);
}

// jl_min / min_nonnan device fast paths (imc_num.h) against the generic definitions, Float64 and Float32: every pair of a
// table of special values (signed zeros, NaNs of both signs, infinities, subnormals, extremes) and random bit patterns
__global__ void k_selftest_min(uint64_t seed, long long per_thread, unsigned long long* mismatches, unsigned long long* tested) {
  const uint64_t special[16] = {0x0000000000000000ull, 0x8000000000000000ull, 0x7ff8000000000000ull, 0xfff8000000000000ull, 0x7ff0000000000000ull,
                                0xfff0000000000000ull, 0x0000000000000001ull, 0x8000000000000001ull, 0x3ff0000000000000ull, 0xbff0000000000000ull,
                                0x7fefffffffffffffull, 0xffefffffffffffffull, 0x7ff0000000000001ull, 0x0010000000000000ull, 0x3fe0000000000000ull, 0x4000000000000000ull};
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t s = mix64(seed ^ (tid * 0x2545F4914F6CDD1Dull));
  unsigned long long bad = 0, n = 0;
  for (long long it = 0; it < per_thread; ++it) {
    s = mix64(s); uint64_t ua = s; s = mix64(s); uint64_t ub = s;
    if (it < 256) { ua = special[(it >> 4) & 15]; ub = special[it & 15]; }      // all 256 special pairs first
    else if ((s & 7) == 0) ub = ua ^ ((s >> 8) & 1 ? 0x8000000000000000ull : 1ull);   // equal magnitudes / neighbours
    const double a = __longlong_as_double((long long)ua), b = __longlong_as_double((long long)ub);
    const float af = __uint_as_float((uint32_t)(ua >> 32)), bf = __uint_as_float((uint32_t)(ub >> 32));
    auto same64 = [](double x, double y) { return __double_as_longlong(x) == __double_as_longlong(y) || (x != x && y != y); };
    auto same32 = [](float x, float y) { return __float_as_uint(x) == __float_as_uint(y) || (x != x && y != y); };
    if (!same64(jl_min(Num<F64>(a), Num<F64>(b)).v, jl_min_generic(Num<F64>(a), Num<F64>(b)).v)) ++bad;
    if (!same64(min_nonnan(Num<F64>(a), Num<F64>(b)).v, min_nonnan_generic(Num<F64>(a), Num<F64>(b)).v)) ++bad;
    if (!same32(jl_min(Num<F32>(af), Num<F32>(bf)).v, jl_min_generic(Num<F32>(af), Num<F32>(bf)).v)) ++bad;
    if (!same32(min_nonnan(Num<F32>(af), Num<F32>(bf)).v, min_nonnan_generic(Num<F32>(af), Num<F32>(bf)).v)) ++bad;
    n += 4;
  }
  if (bad) atomicAdd(mismatches, bad);
  atomicAdd(tested, n);
}

}  // namespace imc

extern "C" int imc_cuda_selftest_min(int device, uint64_t seed, long long per_thread, unsigned long long* mismatches, unsigned long long* tested) {
  using namespace imc;
  if (cudaSetDevice(device) != cudaSuccess) return -3;
  unsigned long long* d = nullptr;
  if (cudaMalloc(&d, 2 * sizeof(unsigned long long)) != cudaSuccess) return -4;
  cudaMemset(d, 0, 2 * sizeof(unsigned long long));
  k_selftest_min<<<148 * 4, 256>>>(seed, per_thread, d, d + 1);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[2] = {0, 0};
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (mismatches) *mismatches = h[0];
  if (tested) *tested = h[1];
  return e == cudaSuccess ? 0 : -3;
}

extern "C" int imc_cuda_selftest_div(int device, uint64_t seed, long long per_thread, unsigned long long* mismatches,
                                     unsigned long long* tested, float first_bad[4]) {
  using namespace imc;
  if (cudaSetDevice(device) != cudaSuccess) return -3;
  unsigned long long* d = nullptr; float* fb = nullptr;
  if (cudaMalloc(&d, 2 * sizeof(unsigned long long)) != cudaSuccess || cudaMalloc(&fb, 4 * sizeof(float)) != cudaSuccess) return -4;
  cudaMemset(d, 0, 2 * sizeof(unsigned long long)); cudaMemset(fb, 0, 4 * sizeof(float));
  for (int mode = 0; mode < 4; ++mode) k_selftest_div<<<148 * 8, 256>>>(seed + mode, per_thread, mode, d, d + 1, fb);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[2] = {0, 0};
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  if (first_bad) cudaMemcpy(first_bad, fb, 4 * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d); cudaFree(fb);
  if (mismatches) *mismatches = h[0];
  if (tested) *tested = h[1];
  return e == cudaSuccess ? 0 : -3;
}
