// imc_exact.cu — stable key sort of the deposit records for the EXACT tally mode (CUB radix sort; the
// rest of the mode — record generation in the tracking kernels, per-cell reduction in the reference's
// summation order — is in imc_kernels.cuh).
#include <cub/device/device_radix_sort.cuh>
#include "imc_exact.h"

namespace imc {

cudaError_t exact_sort_pairs(void* temp, size_t& temp_bytes, const unsigned* keys_in, unsigned* keys_out,
                             const double* vals_in, double* vals_out, long long n, int end_bit, cudaStream_t stream) {
  // radix sort is stable: records of one cell keep their (particle, segment) order; bit 31 (wide flag) is not sorted on
  return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, n, 0, end_bit, stream);
}

}  // namespace imc
