// merge_search.cuh -- stand-alone MergePathSearch kernels: the coordinates of arbitrary diagonals
// (debug / parity export) and of every tile boundary (DeviceSpmvSearchKernel,
// dispatch_spmv_orig.cuh:104-143).  The default CsrMV path does not launch them: the pipe kernel's
// producer warp finds its coordinates itself (spmv_pipe.cuh); they serve mspmv_merge_path_search,
// mspmv_csrmv_swath_coords and the two-launch mode (option pipe_search = 0).
#pragma once

#include "merge_common.cuh"

namespace mspmv {

// One thread per boundary (a 32-ary warp-cooperative variant was measured slower here: 18.9 vs
// 10.7 us for 59k boundaries -- it turns a latency-bound kernel into a load-throughput-bound one).
// Also clears the ticket counter of the carry fold for this call (the temp blob arrives uninitialised).
__global__ void tile_search_kernel(const int* __restrict__ row_end_offsets, int num_rows,
                                   int num_nonzeros, int tile_items, int num_tiles,
                                   int2* __restrict__ coords, unsigned int* __restrict__ ticket)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (ticket != nullptr && i == 0) *ticket = 0u;
    if (i <= num_tiles)
        coords[i] = merge_path_search_global((int64_t)i * tile_items, row_end_offsets, num_rows,
                                             num_nonzeros);
}

// Arbitrary diagonals (debug / parity export).
__global__ void diagonal_search_kernel(const int* __restrict__ row_end_offsets, int num_rows,
                                       int num_nonzeros, const int* __restrict__ diagonals, int n,
                                       int2* __restrict__ coords)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        coords[i] = merge_path_search_global(diagonals[i], row_end_offsets, num_rows, num_nonzeros);
}

}  // namespace mspmv
