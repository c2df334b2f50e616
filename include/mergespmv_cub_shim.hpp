// mergespmv_cub_shim.hpp -- header-only stand-in for the one CUB entry point the reference's
// gpu_spmv.cu uses: cub::DeviceSpmv::CsrMV<ValueT> (cub/device/device_spmv.cuh:129-164), same
// signature and defaults, forwarding to the C ABI of libmergespmv.so.  With this header on the
// include path instead of <cub/device/device_spmv.cuh>, TestGpuMergeCsrmv (gpu_spmv.cu:376-435)
// compiles and runs unchanged.  Needs <cuda_runtime.h> only for cudaError_t / cudaStream_t.
#pragma once

#include <cuda_runtime.h>

#include "mergespmv.h"

namespace cub {

struct DeviceSpmv {
    static cudaError_t CsrMV(void* d_temp_storage, size_t& temp_storage_bytes, float* d_values,
                             int* d_row_offsets, int* d_column_indices, float* d_vector_x, float* d_vector_y,
                             int num_rows, int num_cols, int num_nonzeros, cudaStream_t stream = 0,
                             bool debug_synchronous = false)
    {
        return (cudaError_t)mspmv_csrmv_f32(d_temp_storage, &temp_storage_bytes, d_values, d_row_offsets,
                                            d_column_indices, d_vector_x, d_vector_y, num_rows, num_cols,
                                            num_nonzeros, (mspmv_stream_t)stream, debug_synchronous ? 1 : 0);
    }
    static cudaError_t CsrMV(void* d_temp_storage, size_t& temp_storage_bytes, double* d_values,
                             int* d_row_offsets, int* d_column_indices, double* d_vector_x, double* d_vector_y,
                             int num_rows, int num_cols, int num_nonzeros, cudaStream_t stream = 0,
                             bool debug_synchronous = false)
    {
        return (cudaError_t)mspmv_csrmv_f64(d_temp_storage, &temp_storage_bytes, d_values, d_row_offsets,
                                            d_column_indices, d_vector_x, d_vector_y, num_rows, num_cols,
                                            num_nonzeros, (mspmv_stream_t)stream, debug_synchronous ? 1 : 0);
    }
    // y = alpha*A*x + beta*y (the --alpha/--beta surface the reference's kernels leave disabled)
    static cudaError_t CsrMV(void* d_temp_storage, size_t& temp_storage_bytes, float* d_values,
                             int* d_row_offsets, int* d_column_indices, float* d_vector_x, float* d_vector_y,
                             int num_rows, int num_cols, int num_nonzeros, float alpha, float beta,
                             cudaStream_t stream, bool debug_synchronous)
    {
        return (cudaError_t)mspmv_csrmv_axpby_f32(d_temp_storage, &temp_storage_bytes, d_values, d_row_offsets,
                                                  d_column_indices, d_vector_x, d_vector_y, num_rows, num_cols,
                                                  num_nonzeros, alpha, beta, (mspmv_stream_t)stream,
                                                  debug_synchronous ? 1 : 0);
    }
    static cudaError_t CsrMV(void* d_temp_storage, size_t& temp_storage_bytes, double* d_values,
                             int* d_row_offsets, int* d_column_indices, double* d_vector_x, double* d_vector_y,
                             int num_rows, int num_cols, int num_nonzeros, double alpha, double beta,
                             cudaStream_t stream, bool debug_synchronous)
    {
        return (cudaError_t)mspmv_csrmv_axpby_f64(d_temp_storage, &temp_storage_bytes, d_values, d_row_offsets,
                                                  d_column_indices, d_vector_x, d_vector_y, num_rows, num_cols,
                                                  num_nonzeros, alpha, beta, (mspmv_stream_t)stream,
                                                  debug_synchronous ? 1 : 0);
    }
};

}  // namespace cub
