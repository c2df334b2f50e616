// merge_common.cuh -- device building blocks shared by the sm_100a merge-based CsrMV kernels.
//
// Reference behaviour being re-implemented (file:line in /root/reference):
//   MergePathSearch            cub/thread/thread_search.cuh:53-84 == cpu_spmv.cpp:223-245
//   per-thread merge walk      cub/agent/agent_spmv_orig.cuh:557-578 (staged), :327-358 (direct) --
//                              re-expressed as a bitmap-driven unrolled loop inside the kernels
//   reduce-by-key scan op      cub/thread/thread_operators.cuh:278-302 (ReduceByKeyOp)
//   block reduce-by-key        agent_spmv_orig.cuh:583-626 (BlockScan of KeyValuePair)
// The block-wide scan of (row, partial) pairs is replaced by a warp-shuffle segmented scan of
// (row-ended flag, tail partial) -- one element per thread instead of one per merge item.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx_sm100.cuh"

namespace mspmv {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// Merge-path search along one diagonal of (row_end_offsets  (+)  natural numbers).
// `row_end(i)` returns list A element i (0 <= i < a_len); list B element j is b_base + j.
// Returns the smallest x in [max(diag - b_len, 0), min(diag, a_len)] with
// row_end(x) > b_base + diag - x - 1 (ties consume A first), and y = diag - x: the coordinate
// is unique, so any search order gives the reference's result bit for bit.
// ---------------------------------------------------------------------------------------------
template <typename RowEndFn>
__device__ __forceinline__ int2 merge_path_search(int diag, RowEndFn row_end, int a_len, int b_len,
                                                  int b_base)
{
    int lo = max(diag - b_len, 0);
    int hi = min(diag, a_len);
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (row_end(mid) <= b_base + diag - mid - 1)
            lo = mid + 1;
        else
            hi = mid;
    }
    return make_int2(min(lo, a_len), diag - lo);
}

// Search in global memory over the whole matrix (tile / swath / shard boundaries).
__device__ __forceinline__ int2 merge_path_search_global(int64_t diag64, const int* __restrict__ row_end_offsets,
                                                         int num_rows, int num_nonzeros)
{
    // diagonals past the end clamp to (num_rows, num_nonzeros)
    int64_t total = (int64_t)num_rows + num_nonzeros;
    int diag = (int)(diag64 < total ? diag64 : total);
    return merge_path_search(
        diag, [&](int i) { return __ldg(row_end_offsets + i); }, num_rows, num_nonzeros, 0);
}

// ---------------------------------------------------------------------------------------------
// Warp-cooperative 32-ary merge-path search in global memory: ~log33(range) dependent loads
// instead of log2(range).  Same unique answer as merge_path_search().
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int2 warp_merge_path_search_global(int64_t diag64,
                                                              const int* __restrict__ row_end_offsets,
                                                              int num_rows, int num_nonzeros, int lane)
{
    int64_t total = (int64_t)num_rows + num_nonzeros;
    int diag = (int)(diag64 < total ? diag64 : total);
    int lo = max(diag - num_nonzeros, 0);
    int hi = min(diag, num_rows);
    while (lo < hi) {
        int span = hi - lo;
        // 32 pivots strictly inside [lo, hi); duplicates are fine for tiny spans
        int pivot = lo + (int)(((int64_t)span * (lane + 1)) / 33);
        pivot = min(pivot, hi - 1);
        bool go_up = __ldg(row_end_offsets + pivot) <= diag - pivot - 1;  // predicate is monotone in pivot
        unsigned up = __ballot_sync(kFull, go_up);
        int n_up = __popc(up);
        int new_lo = n_up > 0 ? __shfl_sync(kFull, pivot, n_up - 1) + 1 : lo;
        int new_hi = n_up < 32 ? __shfl_sync(kFull, pivot, n_up) : hi;
        lo = new_lo;
        hi = new_hi;
    }
    return make_int2(min(lo, num_rows), diag - lo);
}

// ---------------------------------------------------------------------------------------------
// Segmented (reduce-by-key) scan element: `ended` = a row end occurred at or after the segment
// start, `val` = partial sum since the last row end.  combine(a, b) with a earlier than b is
// ReduceByKeyOp (thread_operators.cuh:290-300) specialised to "key changed" flags.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Seg {
    T val;
    int ended;
};

template <typename T>
__device__ __forceinline__ Seg<T> seg_combine(const Seg<T>& a, const Seg<T>& b)
{
    Seg<T> r;
    r.ended = a.ended | b.ended;
    r.val = b.ended ? b.val : a.val + b.val;
    return r;
}

template <typename T>
__device__ __forceinline__ T shfl_up(T v, int d)
{
    return __shfl_up_sync(kFull, v, d);
}

// Inclusive segmented scan across the 32 lanes of a warp.
template <typename T>
__device__ __forceinline__ Seg<T> warp_seg_scan_inclusive(Seg<T> s, int lane)
{
    // short-row matrices: every lane closes a row, so every lane starts a new segment and the
    // inclusive result is the lane's own element -- one vote instead of five shuffle rounds
    if (__all_sync(kFull, s.ended)) return s;
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
        T v = shfl_up(s.val, d);
        int e = __shfl_up_sync(kFull, s.ended, d);
        if (lane >= d) {
            if (!s.ended) s.val += v;
            s.ended |= e;
        }
    }
    return s;
}

// Same result, fewer instructions: one ballot tells every lane where its segment starts (the last
// lane below it that closed a row), so the shuffle rounds carry only the value and the flag of the
// result is a mask test.  After round d a lane holds the sum of lanes [max(start, lane-2d+1), lane];
// a lane adds lane-d's partial only if that lane is inside its segment, and a lane that closed a
// row itself never adds -- exactly seg_combine applied left to right.
template <typename T>
__device__ __forceinline__ Seg<T> warp_seg_scan_inclusive_ballot(Seg<T> s, int lane)
{
    const unsigned m = __ballot_sync(kFull, s.ended != 0);
    if (m == kFull) return s;
    const unsigned below = m & ((1u << lane) - 1u);
    // first lane whose partial I may add; 32 (nobody) if I closed a row myself
    const int start = s.ended ? 32 : (below ? 31 - __clz((int)below) : 0);
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
        T v = shfl_up(s.val, d);
        if (lane - d >= start) s.val += v;
    }
    s.ended = (m & ((2u << lane) - 1u)) != 0u;
    return s;
}

// Block-wide exclusive segmented scan over NWARPS*32 threads (all must call).
//   in        this thread's element
//   carry_in  element logically preceding thread 0 (the previous tile's carry-out, or {0,0})
//   excl      out: combine(carry_in, elements of threads 0..t-1)
//   total     out: combine(carry_in, all elements) -- identical in every thread
// s_warp must hold NWARPS Seg<T>.  Contains one named barrier among the participating threads.
template <typename T, int NWARPS, bool BALLOT = false>
__device__ __forceinline__ void block_seg_scan_exclusive(Seg<T> in, Seg<T> carry_in, Seg<T>* s_warp,
                                                         int tid, int barrier_id, Seg<T>& excl,
                                                         Seg<T>& total)
{
    const int lane = tid & 31, warp = tid >> 5;
    Seg<T> inc = BALLOT ? warp_seg_scan_inclusive_ballot(in, lane) : warp_seg_scan_inclusive(in, lane);
    if (lane == 31) s_warp[warp] = inc;

    Seg<T> prev;  // warp-exclusive
    prev.val = shfl_up(inc.val, 1);
    prev.ended = __shfl_up_sync(kFull, inc.ended, 1);
    if (lane == 0) {
        prev.val = T(0);
        prev.ended = 0;
    }
    named_bar_sync(barrier_id, NWARPS * 32);

    Seg<T> run = carry_in;
    Seg<T> before = carry_in;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) {
        if (w == warp) before = run;
        run = seg_combine(run, s_warp[w]);
    }
    excl = seg_combine(before, prev);
    total = run;
}

// y = alpha*sum + beta*y_old epilogue (SpmvGold, gpu_spmv.cu:72-92); AXPBY=false is y = sum.
template <typename T, bool AXPBY>
__device__ __forceinline__ T epilogue(T sum, T alpha, T beta, const T* y_ptr)
{
    if (AXPBY) {
        T r = alpha * sum;
        if (beta != T(0)) r += beta * (*y_ptr);
        return r;
    }
    return sum;
}

}  // namespace mspmv
