// Host check of csrc/imc_sortnet.h: 0-1 principle (a comparator network that sorts every 0/1 input sorts every input).
#include <cstdint>
#include "imc_sortnet.h"

template <int N>
static int check() {
  for (uint32_t bits = 0; bits < (1u << N); ++bits) {
    double v[N];
    for (int i = 0; i < N; ++i) v[i] = (bits >> i) & 1u ? 1.0 : 0.0;
    imc::sortnet::sort(v);
    for (int i = 1; i < N; ++i) if (v[i - 1] > v[i]) return 0;
  }
  return 1;
}
extern "C" int sortnet_ok(int n) {
  switch (n) {
    case 2: return check<2>(); case 3: return check<3>(); case 4: return check<4>(); case 5: return check<5>(); case 6: return check<6>();
    case 7: return check<7>(); case 8: return check<8>(); case 9: return check<9>(); case 10: return check<10>(); case 11: return check<11>();
    case 12: return check<12>(); case 13: return check<13>(); case 14: return check<14>(); case 16: return check<16>();
    default: return -1;
  }
}
// also against a plain sort on arbitrary doubles (duplicates, negatives, zeros, infinities)
extern "C" int sortnet_sort13(double* v) { double a[13]; for (int i = 0; i < 13; ++i) a[i] = v[i]; imc::sortnet::sort(a); for (int i = 0; i < 13; ++i) v[i] = a[i]; return 0; }
