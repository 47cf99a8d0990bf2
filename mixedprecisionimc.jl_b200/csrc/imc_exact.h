#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
namespace imc {
cudaError_t exact_sort_pairs(void* temp, size_t& temp_bytes, const unsigned* keys_in, unsigned* keys_out,
                             const double* vals_in, double* vals_out, long long n, int end_bit, cudaStream_t stream);
}
