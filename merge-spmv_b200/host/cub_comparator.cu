// cub_comparator.cu -- same-box yardstick, NOT part of the product: the toolkit's own (deprecated)
// cub::DeviceSpmv::CsrMV -- the maintained descendant of the reference's kernels
// (cub/device/device_spmv.cuh:129-164 in the reference; <cub/device/device_spmv.cuh> of CUDA 12.9
// here) -- compiled for sm_100a.  The reference's bundled 2016 CUB no longer compiles (texture
// references, non-sync shuffles), so this is the closest thing to "the reference's GPU code on a
// B200".  It lives in its own translation unit because include/mergespmv_cub_shim.hpp defines the
// same cub::DeviceSpmv name for the drop-in; gpu_spmv --cub times it next to the merge CsrMV.
#include <cub/device/device_spmv.cuh>

extern "C" int toolkit_cub_csrmv_f64(void* temp, size_t* bytes, const double* values, const int* row_offsets,
                                     const int* col, const double* x, double* y, int rows, int cols, int nnz,
                                     cudaStream_t stream)
{
    return (int)cub::DeviceSpmv::CsrMV(temp, *bytes, values, row_offsets, col, x, y, rows, cols, nnz, stream);
}

extern "C" int toolkit_cub_csrmv_f32(void* temp, size_t* bytes, const float* values, const int* row_offsets,
                                     const int* col, const float* x, float* y, int rows, int cols, int nnz,
                                     cudaStream_t stream)
{
    return (int)cub::DeviceSpmv::CsrMV(temp, *bytes, values, row_offsets, col, x, y, rows, cols, nnz, stream);
}
