// spmv_tile.cuh -- "tile" engine (default): one small threadblock per fixed-length diagonal swath
// ("tile") of the (row_end_offsets (+) N_nnz) merge path, at full SM occupancy.
//
// Shape of one CsrMV (the reference's launch sequence, dispatch_spmv_orig.cuh:665-745, re-built
// for sm_100a):
//   1. tile_search_kernel      MergePathSearch at every tile boundary (DeviceSpmvSearchKernel,
//                              dispatch_spmv_orig.cuh:104-143) -> (row, nonzero) coordinates
//   2. spmv_tile_kernel        128 threads x IPT merge items per block.  One warp stages the tile's
//                              values / column indices / row offsets into shared memory with
//                              cp.async.bulk (TMA, SASS UBLKCP) completing on an mbarrier -- copies
//                              are 16-byte aligned for any tile start because the slice is taken
//                              from the aligned-down address (ragged ends patched by scalar
//                              copies).  Then: strip-mined LDG gather of x[col] and in-place
//                              multiply (agent_spmv_orig.cuh:472-494); row ends scattered into a
//                              bitmap (bit p <=> merge item p is a row end); every thread's start
//                              coordinate is a popcount prefix of that bitmap -- the unique
//                              coordinate the reference finds with a per-thread MergePathSearch
//                              (agent_spmv_orig.cuh:539-545); branch-free unrolled walk over its
//                              IPT flag bits (:557-578); warp-shuffle segmented scan closes rows
//                              that span threads; finished rows go to y from registers; the
//                              partial last row is the tile's carry-out (:906-913).
//   3. carry_fixup_kernel(s)   deterministic two-level segmented reduction of the per-tile carries
//                              into y (replaces DeviceSegmentFixupKernel,
//                              dispatch_spmv_orig.cuh:199-224; no atomics, same result every run,
//                              guard row < num_rows that the reference's GPU path lacks).
// Latency (HBM stream, L2 gathers, shuffles) is hidden by thread-level parallelism: ~19 KB of
// shared memory per block lets 10-16 blocks share an SM.
#pragma once

#include <limits.h>

#include "merge_common.cuh"
#include "tma_stage.cuh"  // stage_superset(): TMA bulk copies of aligned supersets

namespace mspmv {

template <typename T>
struct TileCfg {
#ifndef MSPMV_TILE_THREADS
#define MSPMV_TILE_THREADS 128
#endif
    // Merge items per thread, tuned on B200 (profiles/tuning_r01.txt): fp64 9, fp32 13 -- both give
    // ~15 KB of shared memory per block.  Odd, so the strided shared-memory walk is conflict-free
    // for dense spans.
#ifndef MSPMV_TILE_IPT
#define MSPMV_TILE_IPT (sizeof(T) == 8 ? 9 : 13)
#endif
    static constexpr int THREADS = MSPMV_TILE_THREADS;
    static constexpr int IPT = MSPMV_TILE_IPT;
    static constexpr int TILE = THREADS * IPT;    // merge items per block
    static constexpr int BW = TILE / 32 + 2;      // bitmap words
    static constexpr int FIX = 256;               // carries per fix-up block
    static constexpr int LOCAL_SPAN = 32768;      // a warp whose columns span fewer elements than this lets its gathers allocate in L1
    static constexpr int ROWCAP = 384;            // row offsets staged in shared memory; tiles with more read them from L2
    // Odd IPT keeps the thread-blocked shared-memory reads conflict-free where threads hold no row end
    // (long rows); where every thread holds exactly one (rows of ~IPT items, e.g. the banded config)
    // an even IPT does (tools/lsu_model.py).  Both are legal; the bitmap layout only needs IPT < 32.
    static_assert(TILE % 32 == 0 && IPT >= 2 && IPT < 32, "bitmap layout");
};

// ---- step 1: tile boundary coordinates (DeviceSpmvSearchKernel, dispatch_spmv_orig.cuh:104-143)
// One thread per boundary (a 32-ary warp-cooperative variant was measured slower here: 18.9 vs
// 10.7 us for 59k boundaries -- it turns a latency-bound kernel into a load-throughput-bound one).
// Also clears the fix-up ticket counter for this call (the temp blob arrives uninitialised).
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

// ---- step 2: one tile per threadblock
// tile_body: everything a block does with its tile once the tile's start / end coordinates (c0, c1)
// are known.  Shared by the three-launch path (coordinates from tile_search_kernel) and the
// single-launch path for small matrices (coordinates searched by the block itself).
template <typename T, bool AXPBY>
__device__ __forceinline__ void tile_body(
    const T* __restrict__ values, const int* __restrict__ row_offsets,
    const int* __restrict__ column_indices, const T* __restrict__ x, T* __restrict__ y,
    const int2* __restrict__ coords, const int tid, const int tile, const int2 c0, const int2 c1,
    int* __restrict__ carry_rows, T* __restrict__ carry_vals, T alpha,
    T beta, int num_rows, int num_nonzeros, int shift_v, int shift_c, int shift_r, int prefetch_ahead)
{
    using C = TileCfg<T>;
    constexpr int IPT = C::IPT;
    constexpr int NW = C::THREADS / 32;
    constexpr int GV = 16 / (int)sizeof(T);  // elements per 16 bytes
    alignas(128) __shared__ T s_val[C::TILE + 2 * GV];
    alignas(128) __shared__ int s_col[C::TILE + 8];
    alignas(128) __shared__ int s_row[C::ROWCAP + 8];
    alignas(16) __shared__ uint32_t s_bits[C::BW];
    alignas(16) __shared__ Seg<T> s_warp[NW];
    alignas(8) __shared__ uint64_t s_bar;

    const int warp = tid >> 5, lane = tid & 31;
    const int x0 = c0.x, y0 = c0.y;
    const int nrows = c1.x - c0.x;           // rows that end in this tile
    const int nnzs = c1.y - c0.y;
    const int items = nrows + nnzs;

    // element i of an array lives at buffer position (i + shift) - base, base = aligned-down start
    const int base_v = (y0 + shift_v) & ~(GV - 1);
    const int base_c = (y0 + shift_c) & ~3;
    const int jr0 = x0 + 1;                  // row_end_offsets[x0 + r] == row_offsets[jr0 + r]
    const int base_r = (jr0 + shift_r) & ~3;
    const int off_v = y0 + shift_v - base_v, off_c = y0 + shift_c - base_c, off_r = jr0 + shift_r - base_r;

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    if (tid < C::BW) s_bits[tid] = 0u;
    __syncthreads();

    // ---- TMA staging of the tile (warp 0) -----------------------------------------------------
    if (warp == 0) {
        const uint64_t policy = l2_policy_evict_first();
        uint32_t b = stage_superset<T>(values, shift_v, y0, y0 + nnzs, num_nonzeros, s_val, base_v, &s_bar,
                                       policy, lane);
        b += stage_superset<int>(column_indices, shift_c, y0, y0 + nnzs, num_nonzeros, s_col, base_c, &s_bar,
                                 policy, lane);
        if (nrows <= C::ROWCAP)
            b += stage_superset<int>(row_offsets, shift_r, jr0, jr0 + nrows, num_rows + 1, s_row, base_r,
                                     &s_bar, policy, lane);
        __syncwarp();
        if (lane == 0) {
            if (b) mbar_arrive_expect_tx(&s_bar, b);
            else mbar_arrive(&s_bar);
        }
    }
    // ---- L2 prefetch for the tile `prefetch_ahead` tiles later (warp 1, off the critical path) ---
    if (prefetch_ahead > 0 && warp == 1 && lane == 0) {
        const long long ft = (long long)tile + prefetch_ahead;
        if (ft < (long long)gridDim.x) {
            const int2 f0 = __ldg(coords + ft), f1 = __ldg(coords + ft + 1);
            l2_prefetch_range<T>(values, shift_v, f0.y, f1.y, num_nonzeros);
            l2_prefetch_range<int>(column_indices, shift_c, f0.y, f1.y, num_nonzeros);
            l2_prefetch_range<int>(row_offsets, shift_r, f0.x + 1, f1.x + 1, num_rows + 1);
        }
    }
    mbar_wait(&s_bar, 0);

    // ---- gather x[col] and multiply in place, strip-mined so the index stream is coalesced -----
    {
        int cidx[IPT];
        T xv[IPT];
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const int j = tid + i * C::THREADS;
            cidx[i] = j < nnzs ? s_col[off_c + j] : -1;
        }
#if MSPMV_GATHER_FLAVOUR == 5
        // x is the only reused data: keep it in L2 (evict_last).  L1: a warp whose columns span a
        // narrow range re-uses lines (banded / FEM-like locality) and lets them allocate; a warp with
        // scattered columns has no L1 reuse, and there L1 capacity only limits the misses in flight,
        // so its gathers do not allocate.  Warp-uniform decision from two sampled columns per thread.
        const uint64_t keep = l2_policy_evict_last();
        int cmin = cidx[0] >= 0 ? cidx[0] : INT_MAX, cmax = cidx[0];
        if (cidx[IPT - 1] >= 0) {
            cmin = min(cmin, cidx[IPT - 1]);
            cmax = max(cmax, cidx[IPT - 1]);
        }
        cmin = __reduce_min_sync(kFull, cmin);
        cmax = __reduce_max_sync(kFull, cmax);
        if (cmax - cmin < C::LOCAL_SPAN) {
#pragma unroll
            for (int i = 0; i < IPT; ++i) xv[i] = cidx[i] >= 0 ? ld_gather_l1(x + cidx[i], keep) : T(0);
        } else {
#pragma unroll
            for (int i = 0; i < IPT; ++i) xv[i] = cidx[i] >= 0 ? ld_gather(x + cidx[i], keep) : T(0);
        }
#elif MSPMV_GATHER_FLAVOUR >= 3
        const uint64_t keep = l2_policy_evict_last();  // x is the only reused data: keep it in L2
#pragma unroll
        for (int i = 0; i < IPT; ++i) xv[i] = cidx[i] >= 0 ? ld_gather(x + cidx[i], keep) : T(0);
#else
#pragma unroll
        for (int i = 0; i < IPT; ++i) xv[i] = cidx[i] >= 0 ? ld_gather(x + cidx[i]) : T(0);
#endif
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const int j = tid + i * C::THREADS;
            if (j < nnzs) s_val[off_v + j] *= xv[i];
        }
    }
    // ---- row-end flags: merge item p = row_end[r] - y0 + r is the end of local row r ------------
    for (int r = tid; r < nrows; r += C::THREADS) {
        const int e = nrows <= C::ROWCAP ? s_row[off_r + r] : __ldg(row_offsets + jr0 + r);
        const int pos = e - y0 + r;
        atomicOr(&s_bits[pos >> 5], 1u << (pos & 31));
    }
    __syncthreads();

    // ---- my IPT flag bits; row ends before my first item == my start row (popcount prefix) ------
    const int diag = tid * IPT;
    const uint32_t w0 = s_bits[diag >> 5], w1 = s_bits[(diag >> 5) + 1];
    const uint32_t bits = __funnelshift_r(w0, w1, diag & 31) & ((1u << IPT) - 1u);
    int before_warp = 0;
#pragma unroll
    for (int k = lane; k < NW * IPT; k += 32)
        if (k < warp * IPT) before_warp += __popc(s_bits[k]);  // warp w owns words [w*IPT, (w+1)*IPT)
    before_warp = __reduce_add_sync(kFull, before_warp);
#if defined(MSPMV_PREFIX_SHFL)
    // whole words of my warp before mine: lane k < IPT counts word k of the warp, 4-step scan, pick mine
    int wcnt = lane < IPT ? __popc(s_bits[warp * IPT + lane]) : 0;
    int wpre = wcnt;
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
        int v = __shfl_up_sync(kFull, wpre, d);
        if (lane >= d) wpre += v;
    }
    const int rel = (diag >> 5) - warp * IPT;  // my first word, relative to the warp's first word
    int in_warp = __shfl_sync(kFull, wpre - wcnt, rel) + __popc(w0 & ((1u << (diag & 31)) - 1u));
#else
    // whole words of my warp before mine (independent broadcast loads), plus the low part of my word
    int in_warp = __popc(w0 & ((1u << (diag & 31)) - 1u));
#pragma unroll
    for (int k = 0; k < IPT - 1; ++k)
        if (warp * IPT + k < (diag >> 5)) in_warp += __popc(s_bits[warp * IPT + k]);
#endif
    const int xs = before_warp + in_warp;

    // ---- serial walk over my IPT merge items (cpu_spmv.cpp:324-340; agent_spmv_orig.cuh:557-578)
    int ny = off_v + diag - xs;              // buffer position of my first product
    T sums[IPT];
    T running = T(0);
    if (items == C::TILE) {                  // every tile but the last: all IPT items exist
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const bool is_end = (bits >> i) & 1u;
            if (!is_end) running += s_val[ny];
            ny += is_end ? 0 : 1;
            sums[i] = running;
            if (is_end) running = T(0);
        }
    } else {
        const int my_items = min(max(items - diag, 0), IPT);
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const bool is_end = (bits >> i) & 1u;
            if (!is_end && i < my_items) running += s_val[ny];
            ny += is_end ? 0 : 1;
            sums[i] = running;
            if (is_end) running = T(0);
        }
    }
    Seg<T> elem, zero, excl, total;
    elem.val = running;
    elem.ended = bits != 0u;
    zero.val = T(0);
    zero.ended = 0;
    block_seg_scan_exclusive<T, NW>(elem, zero, s_warp, tid, 1, excl, total);

    // ---- finished rows to y from registers; my first row end also takes the carry-in -------------
    {
        int row = x0 + xs;
        T add = excl.val;
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            if ((bits >> i) & 1u) {
                y[row] = epilogue<T, AXPBY>(sums[i] + add, alpha, beta, y + row);
                add = T(0);
                ++row;
            }
        }
    }
    // carry-out: the row that continues into the next tile (agent_spmv_orig.cuh:906-913).
    // c1.x may equal num_rows; the fix-up drops such carries.
    if (tid == 0) {
        carry_rows[tile] = c1.x;
        carry_vals[tile] = total.val;
    }
}

template <typename T, bool AXPBY>
__global__ __launch_bounds__(TileCfg<T>::THREADS) void spmv_tile_kernel(
    const T* __restrict__ values, const int* __restrict__ row_offsets,
    const int* __restrict__ column_indices, const T* __restrict__ x, T* __restrict__ y,
    const int2* __restrict__ coords, int* __restrict__ carry_rows, T* __restrict__ carry_vals, T alpha,
    T beta, int num_rows, int num_nonzeros, int shift_v, int shift_c, int shift_r, int prefetch_ahead)
{
    const int tid = threadIdx.x;  // read here: the launch bound gives the compiler its range
    const int tile = blockIdx.x;
    const int2 c0 = __ldg(coords + tile);
    const int2 c1 = __ldg(coords + tile + 1);
    tile_body<T, AXPBY>(values, row_offsets, column_indices, x, y, coords, tid, tile, c0, c1, carry_rows, carry_vals,
                        alpha, beta, num_rows, num_nonzeros, shift_v, shift_c, shift_r, prefetch_ahead);
}

// ---- single-launch path for small matrices -------------------------------------------------------
// The paper (section IV.B) and the reference (dispatch_spmv_orig.cuh:674-679: the search kernel is
// skipped when there are fewer tiles than SMs can hide) name the extra launches as the small-matrix
// overhead.  Here a matrix of at most a few thousand tiles runs in ONE launch:
//   * warps 0 and 1 find the block's start / end coordinate themselves with the warp-cooperative
//     32-ary MergePathSearch (4 dependent L2 round trips instead of ~20);
//   * tile_body as usual;
//   * the last block to finish (ticket counter, zeroed by a 4-byte memset node in front of the
//     launch) folds ALL carries into y with a block-wide segmented scan, 128 at a time, in carry
//     order -- runs of equal rows are summed left to right and added to y once, the order of the
//     CPU loop cpu_spmv.cpp:348-352; carries with row >= num_rows are dropped.
// Other blocks' y stores and carries are visible to the last block: every block fences before
// taking its ticket, and the last block's second fence invalidates its L1 (SASS CCTL.IVALL).
template <typename T, bool AXPBY>
__global__ __launch_bounds__(TileCfg<T>::THREADS) void spmv_tile_fused_kernel(
    const T* __restrict__ values, const int* __restrict__ row_offsets,
    const int* __restrict__ column_indices, const T* __restrict__ x, T* __restrict__ y,
    int* __restrict__ carry_rows, T* __restrict__ carry_vals, T alpha, T beta, int num_rows,
    int num_nonzeros, int shift_v, int shift_c, int shift_r, unsigned int* __restrict__ ticket)
{
    using C = TileCfg<T>;
    constexpr int NW = C::THREADS / 32;
    __shared__ int2 s_coord[2];
    __shared__ bool s_last;
    alignas(16) __shared__ Seg<T> s_fix[NW];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x;
    if (warp < 2) {
        const int2 c = warp_merge_path_search_global((int64_t)(tile + warp) * C::TILE, row_offsets + 1, num_rows,
                                                     num_nonzeros, lane);
        if (lane == 0) s_coord[warp] = c;
    }
    __syncthreads();
    const int2 c0 = s_coord[0], c1 = s_coord[1];
    tile_body<T, AXPBY>(values, row_offsets, column_indices, x, y, nullptr, tid, tile, c0, c1, carry_rows, carry_vals,
                        alpha, beta, num_rows, num_nonzeros, shift_v, shift_c, shift_r, 0);
    const int n = gridDim.x;
    if (n == 1) return;  // a single tile has no carry to fold (dispatch_spmv_orig.cuh:721)

    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(ticket, 1u) == (unsigned)n - 1u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    Seg<T> carry;  // the run that touches the end of the previous chunk
    carry.val = T(0);
    carry.ended = 0;
    for (int base = 0; base < n; base += C::THREADS) {
        const int i = base + tid;
        const int row = i < n ? carry_rows[i] : INT_MAX;
        const int prev_row = (i > 0 && i < n) ? carry_rows[i - 1] : -1;
        const int next_row = i + 1 < n ? carry_rows[i + 1] : INT_MAX;
        Seg<T> e, excl, total;
        e.val = i < n ? carry_vals[i] : T(0);
        e.ended = row != prev_row;  // a new run starts here
        block_seg_scan_exclusive<T, NW>(e, carry, s_fix, tid, 1, excl, total);
        const T run_sum = e.ended ? e.val : excl.val + e.val;  // my run up to and including me
        if (i < n && row != next_row && row < num_rows) y[row] += AXPBY ? alpha * run_sum : run_sum;
        carry.val = total.val;
        carry.ended = 0;
        __syncthreads();  // s_fix is reused by the next chunk
    }
}

// ---- step 3: deterministic carry fix-up -------------------------------------------------------
// Level 1: blocks of FIX consecutive carries (sorted by row).  Within a block, runs of equal rows
// are summed left to right; a run that ends inside the block is added to y once; the run that
// touches the end of the block is handed to level 2 as (row, partial).
template <typename T, bool AXPBY>
__global__ __launch_bounds__(256) void carry_fixup_block_kernel(const int* __restrict__ carry_rows,
                                                                const T* __restrict__ carry_vals, int n,
                                                          int num_rows, T* __restrict__ y, T alpha,
                                                          int* __restrict__ carry2_rows,
                                                          T* __restrict__ carry2_vals,
                                                          unsigned int* __restrict__ ticket)
{
    constexpr int FIX = 256;
    __shared__ Seg<T> s_warp[FIX / 32];
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    const int i = blockIdx.x * FIX + tid;
    const int row = i < n ? carry_rows[i] : INT_MAX;
    const T val = i < n ? carry_vals[i] : T(0);
    const int prev_row = (tid > 0 && i < n) ? carry_rows[i - 1] : -1;  // block-local run detection; threads past n read nothing
    const int next_row = (i + 1 < n) ? carry_rows[i + 1] : INT_MAX;

    // inclusive segmented scan keyed by "row differs from the previous carry in this block"
    Seg<T> e;
    e.val = val;
    e.ended = (tid == 0) || (row != prev_row);  // a new run starts here
    // scan with "run start" flags: value accumulates since the most recent start
    const int lane = tid & 31, warp = tid >> 5;
    Seg<T> inc = e;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T v = shfl_up(inc.val, d);
        int f = __shfl_up_sync(kFull, inc.ended, d);
        if (lane >= d) {
            if (!inc.ended) inc.val += v;
            inc.ended |= f;
        }
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    Seg<T> run;
    run.val = T(0);
    run.ended = 1;
    for (int w = 0; w < warp; ++w) {
        Seg<T> t = s_warp[w];
        run.val = t.ended ? t.val : run.val + t.val;
    }
    const T run_sum = inc.ended ? inc.val : run.val + inc.val;  // sum of my run up to and including me

    const bool last_in_block = (tid == FIX - 1) || (i == n - 1);
    if (i < n) {
        if (last_in_block) {
            carry2_rows[blockIdx.x] = row;
            carry2_vals[blockIdx.x] = run_sum;
        } else if (row != next_row && row < num_rows) {
            y[row] += AXPBY ? alpha * run_sum : run_sum;
        }
    }

    // Level 2 in the same launch: the last block to finish (ticket counter cleared by the search
    // kernel) folds the per-block carries.  Runs are short here (a row must span > FIX tiles to
    // put two entries in one run); the thread owning the first entry of a run sums it left to
    // right and adds it to y once -- the order of the CPU loop cpu_spmv.cpp:348-352.
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int nb = gridDim.x;
    for (int b = tid; b < nb; b += FIX) {
        const int r2 = carry2_rows[b];
        if (r2 >= num_rows) continue;
        if (b > 0 && carry2_rows[b - 1] == r2) continue;  // not the head of this row's run
        T sum = carry2_vals[b];
        for (int j = b + 1; j < nb && carry2_rows[j] == r2; ++j) sum += carry2_vals[j];
        y[r2] += AXPBY ? alpha * sum : sum;
    }
}

// Level 2 (and the stream engine's only level): the n carries are sorted by row; the thread that
// owns the first carry of a run sums the run left to right and adds it to y once -- the order of
// the CPU loop cpu_spmv.cpp:348-352.  Runs here are short (a row must span > FIX tiles to put two
// entries in one run).
template <typename T, bool AXPBY>
__global__ void carry_fixup_runs_kernel(const int* __restrict__ carry_rows, const T* __restrict__ carry_vals,
                                        int n, int num_rows, T* __restrict__ y, T alpha)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int row = carry_rows[i];
    if (row >= num_rows) return;
    if (i > 0 && carry_rows[i - 1] == row) return;  // not the head of this row's run
    T sum = carry_vals[i];
    for (int j = i + 1; j < n && carry_rows[j] == row; ++j) sum += carry_vals[j];
    y[row] += AXPBY ? alpha * sum : sum;
}

}  // namespace mspmv
