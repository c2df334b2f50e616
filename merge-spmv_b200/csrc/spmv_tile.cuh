// spmv_tile.cuh -- "tile" engine: one threadblock per fixed-size merge tile.
//
// This is the simple engine (selected with MSPMV_ENGINE=tile, and used for problems too small
// for the persistent streaming engine in spmv_stream.cuh).  It keeps the reference's three-step
// shape -- search kernel, tile kernel, carry fix-up kernel (dispatch_spmv_orig.cuh:665-745) --
// but every step is re-written for sm_100a: plain LDG staging into shared memory, a per-thread
// merge walk over smem products, a warp-shuffle segmented scan instead of BlockScan over
// KeyValuePairs, and a deterministic (atomic-free) fix-up for both fp32 and fp64.
#pragma once

#include <limits.h>

#include "merge_common.cuh"

namespace mspmv {

// ---- step 1: tile boundary coordinates (DeviceSpmvSearchKernel, dispatch_spmv_orig.cuh:104-143)
__global__ void tile_search_kernel(const int* __restrict__ row_end_offsets, int num_rows,
                                   int num_nonzeros, int tile_items, int num_tiles,
                                   int2* __restrict__ coords)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
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

// ---- step 2: one tile per threadblock (DeviceSpmvKernel / AgentSpmv::ConsumeTile,
// dispatch_spmv_orig.cuh:158-186, agent_spmv_orig.cuh:413-639,856-914)
template <typename T, int THREADS, int IPT, bool AXPBY>
__global__ __launch_bounds__(THREADS) void spmv_tile_kernel(
    const T* __restrict__ values, const int* __restrict__ row_end_offsets,
    const int* __restrict__ column_indices, const T* __restrict__ x, T* __restrict__ y,
    int num_rows, const int2* __restrict__ coords, int* __restrict__ carry_rows,
    T* __restrict__ carry_vals, T alpha, T beta)
{
    constexpr int TILE = THREADS * IPT;
    constexpr int NWARPS = THREADS / 32;
    __shared__ int s_row_end[TILE + 1];
    __shared__ T s_prod[TILE];
    __shared__ T s_y[TILE];
    __shared__ Seg<T> s_warp[NWARPS];

    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int2 c0 = coords[tile];
    const int2 c1 = coords[tile + 1];
    const int x0 = c0.x, y0 = c0.y;
    const int nrows = c1.x - c0.x;
    const int nnzs = c1.y - c0.y;
    const int items = nrows + nnzs;

    // gather: products of the tile's nonzeros, strip-mined so loads coalesce
    // (agent_spmv_orig.cuh:472-494)
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        int j = tid + i * THREADS;
        if (j < nnzs) {
            int c = __ldg(column_indices + y0 + j);
            T v = __ldg(values + y0 + j);
            s_prod[j] = v * __ldg(x + c);
        }
    }
    // row-end offsets of the rows that end in this tile, plus a sentinel instead of the
    // reference's tile_num_rows+1'th load (agent_spmv_orig.cuh:527-531)
    for (int j = tid; j < nrows; j += THREADS) s_row_end[j] = __ldg(row_end_offsets + x0 + j);
    if (tid == 0) s_row_end[nrows] = INT_MAX;
    __syncthreads();

    Seg<T> elem;
    int head_row;
    T head_val;
    thread_merge_walk<T, IPT>(
        tid * IPT, items, nrows, nnzs, y0, [&](int i) { return s_row_end[i]; },
        [&](int j) { return s_prod[j]; }, [&](int r, T v) { s_y[r] = v; }, elem, head_row, head_val);

    Seg<T> zero;
    zero.val = T(0);
    zero.ended = 0;
    Seg<T> excl, total;
    block_seg_scan_exclusive<T, NWARPS>(elem, zero, s_warp, tid, 1, excl, total);
    if (elem.ended) s_y[head_row] = head_val + excl.val;
    __syncthreads();

    for (int j = tid; j < nrows; j += THREADS)
        y[x0 + j] = epilogue<T, AXPBY>(s_y[j], alpha, beta, y + x0 + j);

    // carry-out: partial sum of the row that continues into the next tile
    // (agent_spmv_orig.cuh:906-913).  Row index c1.x may equal num_rows; the fix-up drops it.
    if (tid == 0) {
        carry_rows[tile] = c1.x;
        carry_vals[tile] = total.val;
    }
}

// ---- step 3: deterministic carry fix-up (replaces DeviceSegmentFixupKernel,
// dispatch_spmv_orig.cuh:199-224 / agent_segment_fixup.cuh:226-341).  The n carries are sorted
// by row.  The thread owning the first carry of each row sums that row's run left to right
// and adds it to y once: same order as the CPU loop cpu_spmv.cpp:348-352, no atomics, and the
// guard row < num_rows that the reference's GPU path lacks.
template <typename T, bool AXPBY>
__global__ void carry_fixup_kernel(const int* __restrict__ carry_rows, const T* __restrict__ carry_vals,
                                   int n, int num_rows, T* __restrict__ y, T alpha)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int row = carry_rows[i];
    if (row >= num_rows) return;
    if (i > 0 && carry_rows[i - 1] == row) return;  // not the head of this row's run
    T sum = carry_vals[i];
    for (int j = i + 1; j < n && carry_rows[j] == row; ++j) sum += carry_vals[j];
    if (AXPBY) sum *= alpha;
    y[row] += sum;
}

}  // namespace mspmv
