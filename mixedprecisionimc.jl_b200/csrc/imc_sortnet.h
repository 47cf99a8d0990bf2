// imc_sortnet.h — ascending sort of a small fixed-size array held in registers (host and device).
//
// Utilities.sorter (imc_utilities.jl:23-54) sorts the <= 13 factors of a product before multiplying them pairwise
// (smallest x largest ...), once per cell and quantity.  A run-time-indexed insertion sort keeps the array in local
// memory and is bound by its dependent loads (k_src_energies: 1250 instructions per cell at 27 % issue utilisation);
// the Bose-Nelson network below has only compile-time indices, so the array lives in registers and the
// compare-exchanges of one layer are independent.  Any correct sort produces the same array for finite values, so the
// result is the reference's; tests/test_sortnet.py proves each size with the 0-1 principle (all 2^N inputs).
#pragma once

#if defined(__CUDACC__)
#define IMC_SN_HD __host__ __device__ __forceinline__
#else
#define IMC_SN_HD inline
#endif

namespace imc {
namespace sortnet {

template <class T, int N>
IMC_SN_HD void cswap(T (&v)[N], int i, int j) {   // i < j: afterwards v[i] <= v[j]
  const T a = v[i], b = v[j];
  const bool sw = a > b;
  v[i] = sw ? b : a;
  v[j] = sw ? a : b;
}

// Bose-Nelson: merge the sorted runs [I, I + X) and [J, J + Y)
template <class T, int N, int I, int X, int J, int Y>
IMC_SN_HD void merge(T (&v)[N]) {
  if constexpr (X == 1 && Y == 1) cswap(v, I, J);
  else if constexpr (X == 1 && Y == 2) { cswap(v, I, J + 1); cswap(v, I, J); }
  else if constexpr (X == 2 && Y == 1) { cswap(v, I, J); cswap(v, I + 1, J); }
  else {
    constexpr int A = X / 2;
    constexpr int B = (X & 1) ? Y / 2 : (Y + 1) / 2;
    merge<T, N, I, A, J, B>(v);
    merge<T, N, I + A, X - A, J + B, Y - B>(v);
    merge<T, N, I + A, X - A, J, B>(v);
  }
}
template <class T, int N, int I, int M>
IMC_SN_HD void sort_range(T (&v)[N]) {
  if constexpr (M > 1) {
    constexpr int A = M / 2;
    sort_range<T, N, I, A>(v);
    sort_range<T, N, I + A, M - A>(v);
    merge<T, N, I, A, I + A, M - A>(v);
  }
}
template <class T, int N>
IMC_SN_HD void sort(T (&v)[N]) { sort_range<T, N, 0, N>(v); }

}  // namespace sortnet
}  // namespace imc
