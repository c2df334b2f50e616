// spmv_tile3.cuh -- tile engine, variant 3 (opt-in: mspmv_set_option("tile_variant", 3)).
//
// Same tiles, staging, bitmap, prefix, scan, stores and carries as tile_body (spmv_tile.cuh); what
// changes is the middle.  tile_body gathers x[col] in strip-mined order, writes the products back to
// shared memory and re-reads them in thread-blocked order for the walk.  Here every thread first
// learns its start coordinate (bitmap + popcount prefix need only the row offsets), then reads ITS
// OWN contiguous run of column indices and values from the staged tile, gathers x for them and
// walks with the products in registers: no product store / reload, one shared-memory read per
// operand; the popcount prefix is replaced by the row owners scattering start rows, the warp scan
// is ballot-driven.  ~1/6 fewer executed instructions per thread (profiles/sass_static_r01.txt,
// tools/sass_lines.py) and ~9 % fewer LSU wavefronts per tile (tools/lsu_model.py).  The order of the
// floating-point operations is unchanged, so the results are bit-identical to tile_body's.
// Trade-off to be measured: the gathers issue after the prefix instead of right after the TMA wait,
// and a warp's gathers cover 32*IPT consecutive nonzeros instead of 32.
#pragma once

#include "spmv_tile.cuh"

#ifndef MSPMV_V3_MIN_BLOCKS
#define MSPMV_V3_MIN_BLOCKS 12  // register budget: 65536 / (128 * 12) = 42 -> 40 registers per thread, like tile_body
#endif
#ifndef MSPMV_V3_BALLOT_SCAN
#define MSPMV_V3_BALLOT_SCAN 1  // 1: ballot-driven warp segmented scan (value-only shuffles); 0: the scan of tile_body
#endif
#ifndef MSPMV_V3_XS_SCATTER
#define MSPMV_V3_XS_SCATTER 1  // 1: start rows scattered by the row owners; 0: popcount prefix as in tile_body
#endif

namespace mspmv {

template <typename T, bool AXPBY>
__device__ __forceinline__ void tile_body_v3(
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
#if MSPMV_V3_XS_SCATTER
    alignas(16) __shared__ int s_xs[C::THREADS];  // start row of every thread, written by the row owners
#endif

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
#if MSPMV_V3_XS_SCATTER
    s_xs[tid] = nrows;  // threads that start after the tile's last row end
#endif
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

    // ---- row-end flags: merge item p = row_end[r] - y0 + r is the end of local row r ------------
#if MSPMV_V3_XS_SCATTER
    // The thread that owns row r also tells every thread t whose first merge item lies in
    // (end of row r-1, end of row r] that it starts at row r: xs(t) = #{rows ending before t*IPT} --
    // the coordinate the reference finds with a per-thread MergePathSearch
    // (agent_spmv_orig.cuh:539-545) -- without any per-thread popcount prefix.  A range covers
    // ceil(row length / IPT) threads, so the loop is short except for rows that span most of a tile.
    for (int r = tid; r < nrows; r += C::THREADS) {
        int e, ep;
        if (nrows <= C::ROWCAP) {
            e = s_row[off_r + r];
            ep = r > 0 ? s_row[off_r + r - 1] : y0 - r;  // r == 0: previous end position -1
        } else {
            e = __ldg(row_offsets + jr0 + r);
            ep = r > 0 ? __ldg(row_offsets + jr0 + r - 1) : y0 - r;
        }
        const int pos = e - y0 + r;
        const int prev = ep - y0 + r - 1;
        atomicOr(&s_bits[pos >> 5], 1u << (pos & 31));
        const int t_hi = pos / IPT;
        for (int t = (prev + IPT) / IPT; t <= t_hi; ++t) s_xs[t] = r;
    }
    __syncthreads();
    const int diag = tid * IPT;
    const uint32_t w0 = s_bits[diag >> 5], w1 = s_bits[(diag >> 5) + 1];
    const uint32_t bits = __funnelshift_r(w0, w1, diag & 31) & ((1u << IPT) - 1u);
    const int xs = s_xs[tid];
#else
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
    // whole words of my warp before mine (independent broadcast loads), plus the low part of my word
    int in_warp = __popc(w0 & ((1u << (diag & 31)) - 1u));
#pragma unroll
    for (int k = 0; k < IPT - 1; ++k)
        if (warp * IPT + k < (diag >> 5)) in_warp += __popc(s_bits[warp * IPT + k]);
    const int xs = before_warp + in_warp;

#endif

    // ---- thread-blocked loads: my nonzeros are the contiguous run k0 .. of the tile's nonzeros ----
    // `skip` bit i: merge item i of mine is not a nonzero (a row end, or past the end of the last
    // tile).  Loads depend only on these bits, so all of them issue before any is consumed.
    const int my_items = min(max(items - diag, 0), IPT);
    const uint32_t skip = bits | ~((1u << my_items) - 1u);
    int cidx[IPT];
    T xv[IPT];
    {
        const int* pc = s_col + (off_c + diag - xs);
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            cidx[i] = -1;
            if (!(skip & (1u << i))) cidx[i] = *pc++;
        }
    }
    {
        // x is the only reused data: keep it in L2 (evict_last); L1 policy per warp by column span,
        // as in tile_body (a warp covers 32*IPT consecutive merge items here)
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
    }

    // ---- serial walk (cpu_spmv.cpp:324-340; agent_spmv_orig.cuh:557-578): multiply and accumulate
    // in one pass.  mul_rn keeps the product a separate rounding (no FMA contraction), so the bits
    // equal tile_body's, which stores the rounded products first.
    T sums[IPT];
    T running = T(0);
    {
        const T* pv = s_val + (off_v + diag - xs);  // values are read only now: fewer live registers
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            if (!(skip & (1u << i))) running += mul_rn(*pv++, xv[i]);
            sums[i] = running;
            if (bits & (1u << i)) running = T(0);
        }
    }
    Seg<T> elem, zero, excl, total;
    elem.val = running;
    elem.ended = bits != 0u;
    zero.val = T(0);
    zero.ended = 0;
    block_seg_scan_exclusive<T, NW, MSPMV_V3_BALLOT_SCAN != 0>(elem, zero, s_warp, tid, 1, excl, total);

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
__global__ __launch_bounds__(TileCfg<T>::THREADS, MSPMV_V3_MIN_BLOCKS) void spmv_tile3_kernel(
    const T* __restrict__ values, const int* __restrict__ row_offsets,
    const int* __restrict__ column_indices, const T* __restrict__ x, T* __restrict__ y,
    const int2* __restrict__ coords, int* __restrict__ carry_rows, T* __restrict__ carry_vals, T alpha,
    T beta, int num_rows, int num_nonzeros, int shift_v, int shift_c, int shift_r, int prefetch_ahead)
{
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int2 c0 = __ldg(coords + tile);
    const int2 c1 = __ldg(coords + tile + 1);
    tile_body_v3<T, AXPBY>(values, row_offsets, column_indices, x, y, coords, tid, tile, c0, c1, carry_rows,
                           carry_vals, alpha, beta, num_rows, num_nonzeros, shift_v, shift_c, shift_r,
                           prefetch_ahead);
}

}  // namespace mspmv
