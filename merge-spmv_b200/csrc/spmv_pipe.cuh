// spmv_pipe.cuh -- "pipe" engine: ONE launch per CsrMV.  Persistent threadblocks, each owning a
// contiguous run of equal-length diagonal swaths ("tiles") of the (row_end_offsets (+) N_nnz) merge
// path, fed by a TMA pipeline.
//
// What one call does (the reference needs three launches: search, spmv, fix-up --
// dispatch_spmv_orig.cuh:665-745):
//   * grid = min(tiles, SMs x resident blocks); block b owns tiles [b*T/G, (b+1)*T/G).
//   * the last warp of every block is the PRODUCER.  It walks the block's piece of the merge path
//     itself: the run's start coordinate by a warp-cooperative 32-ary MergePathSearch over the whole
//     matrix (thread_search.cuh:53-84, same unique coordinate), every further tile boundary by a
//     bounded search whose first probe is a 32-row window placed where the previous tile's row count
//     predicts the boundary (one L2 trip for regular matrices, off the consumers' critical path) --
//     this replaces DeviceSpmvSearchKernel (dispatch_spmv_orig.cuh:104-143).  For each tile it issues
//     cp.async.bulk copies (TMA, SASS UBLKCP) into two shared-memory rings, each slot with a
//     full / empty mbarrier pair: the column indices (needed only until the tile's gathers are
//     issued) and the values + row-offset slice (needed until the tile's rows are stored).
//   * the other warps are the CONSUMERS (all of them compute; no gather/reduce specialisation).
//     Per tile, with nnzs nonzeros and nrows row ends:
//       G   every thread owns IPT consecutive nonzero SLOTS: reads its flag words and values, then its
//           column indices (shared memory), and issues the x[col] gathers (LDG, cache-hinted) -- with
//           AHEAD one tile ahead of the walk (a template switch; measured slower, not shipped);
//       P1  row owners (thread r <-> row r of the tile) set bit (row_end[r] - y0) of a bitmap: a
//           row boundary lies in front of that slot (slot nnzs = boundary at the tile's end);
//       W   walk my slots: running = fma(value, x, running); at a flagged slot the finished
//           segment sum is parked IN PLACE (the slot's own value is already in a register) and the
//           sum restarts (the merge walk of agent_spmv_orig.cuh:557-578 without per-item row
//           bookkeeping);
//       S   one warp-shuffle segmented scan of (had a boundary, tail sum) per thread
//           (ReduceByKeyOp, thread_operators.cuh:278-302) gives every thread the partial that
//           precedes it -- added to its first finished segment (patched in shared memory, or kept
//           in a register until now: FIRST_IN_REG) -- and the tile's carry-out, which stays in
//           registers for the block's next tile;
//       Y   row owners read their row's parked sum and store y, coalesced (empty rows get 0).
//     P1 of tile i+1 runs before the barrier that ends S of tile i, so a tile costs two named
//     barriers among the consumers.
//   * a block's last carry-out (row, partial) goes to global memory; the last block to finish
//     (ticket) folds the G carries into y in carry order -- the serial loop of
//     cpu_spmv.cpp:348-352, guard row < num_rows -- replacing DeviceSegmentFixupKernel
//     (dispatch_spmv_orig.cuh:199-224).  No floating-point atomics: same bits every run.
#pragma once

#include <limits.h>

#include "merge_common.cuh"
#include "tma_stage.cuh"

namespace mspmv {

#ifndef MSPMV_PIPE_VALS_FIRST
#define MSPMV_PIPE_VALS_FIRST 1
#endif

// Tuning aid, never in the shipped library (make variants/libmergespmv_prof.so): thread 0 of every block
// accumulates the SM cycles it spends in each phase of a tile; mspmv_debug_profile() reads the totals.
#ifndef MSPMV_PIPE_PROFILE
#define MSPMV_PIPE_PROFILE 0
#endif
#if MSPMV_PIPE_PROFILE
__device__ unsigned long long g_pipe_prof[8];
__device__ __forceinline__ long long prof_clock()
{
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}
#define PIPE_PROF_MARK(k)                           \
    do {                                            \
        if (tid == 0) {                             \
            const long long _t = prof_clock();      \
            prof_acc[k] += _t - prof_last;          \
            prof_last = _t;                         \
        }                                           \
    } while (0)
#else
#define PIPE_PROF_MARK(k) \
    do {                  \
    } while (0)
#endif

// Compile-time shape of one kernel instantiation.
//   IPT     nonzero slots per consumer thread (odd: conflict-free strided shared-memory reads)
//   VST     slots of the value / row-offset ring      CST   slots of the column-index ring
//   AHEAD   1: gathers of tile i+1 are issued before tile i is walked (software pipelining)
//   NW      consumer warps per block
//   FIR     1: a thread's first finished segment waits in a register for the scan instead of being parked and
//           patched in shared memory -- one scattered shared-memory read less per thread and tile, a few more
//           instructions: +3.5 % on the gather-bound power-law config, -10 % on the instruction-bound banded
//           one (profiles/sweep_r02_switches.txt), hence a property of the shape
//   VF      1: a thread reads its flag words and values BEFORE it issues its gathers (the LSU serves a warp's
//           shared-memory reads and its 32-sector gathers from one queue, in order)
template <typename T, int IPT_, int VST_, int CST_, int AHEAD_, int NW_ = 4, int FIR_ = 0, int VF_ = MSPMV_PIPE_VALS_FIRST>
struct PipeCfg {
    using value_type = T;
    static constexpr int NW = NW_;
    static constexpr int CONSUMERS = NW * 32;
    static constexpr int THREADS = CONSUMERS + 32;  // + the producer warp
    static constexpr int IPT = IPT_;
    static constexpr int TILE = CONSUMERS * IPT;    // merge items per tile
    static constexpr int STAGES = VST_;
    static constexpr int CSTAGES = CST_;
    static constexpr bool AHEAD = AHEAD_ != 0;
    static constexpr bool FIRST_IN_REG = FIR_ != 0;
    static constexpr bool VALS_FIRST = VF_ != 0;
    static constexpr int BW = TILE / 32 + 2;        // bitmap words (slot TILE is never flagged; +1 for the funnel shift)
    // row offsets staged per tile (tiles with more rows read them through L2): a third of the tile, at most 384
    static constexpr int ROWCAP = (TILE / 3 < 384 ? ((TILE / 3 + 7) & ~7) : 384);
    static constexpr int GV = 16 / (int)sizeof(T);
    static constexpr int LOCAL_SPAN = 32768;        // columns; a warp whose indices span fewer lets its gathers allocate in L1
    static_assert(IPT >= 2 && IPT < 32 && BW <= CONSUMERS, "bitmap layout");
    // the row offsets of tile i+1 are read (P1) while tile i still owns its slot: two value slots at least
    static_assert(VST_ >= 2 && CST_ >= 1, "rings");
};

template <class C>
struct alignas(128) PipeStage {
    typename C::value_type val[C::TILE + 2 * C::GV];  // staged values; a flagged slot later holds its parked segment sum
    int row[C::ROWCAP + 8];
};
template <class C>
struct alignas(128) PipeColStage {
    int col[C::TILE + 8];
};

template <class C>
struct alignas(128) PipeCtl {
    uint64_t full[C::STAGES];     // values + row offsets of a tile have landed
    uint64_t empty[C::STAGES];
    uint64_t full_c[C::CSTAGES];  // column indices of a tile have landed
    uint64_t empty_c[C::CSTAGES];
    int4 coord[C::CSTAGES];       // (x0, y0, x1, y1) of the tile in the column slot (read once, kept in registers)
    uint32_t bits[2][C::BW];
    Seg<typename C::value_type> warp[C::NW];
    int last;
};

template <class C>
constexpr size_t pipe_smem_bytes()
{
    return sizeof(PipeCtl<C>) + sizeof(PipeStage<C>) * C::STAGES + sizeof(PipeColStage<C>) * C::CSTAGES;
}

// 32-ary warp search for the coordinate on diagonal `diag`, x known to lie in [lo, hi]: the smallest
// x with x == hi or row_end[x] > diag - x - 1 (thread_search.cuh:53-84; ties consume the row end first).
__device__ __forceinline__ int2 warp_merge_path_search_bounded(int diag, const int* __restrict__ row_end_offsets,
                                                               int num_rows, int lo, int hi, int lane)
{
    while (lo < hi) {
        const int span = hi - lo;
        int pivot = lo + (int)(((int64_t)span * (lane + 1)) / 33);
        pivot = min(pivot, hi - 1);
        const bool go_up = __ldg(row_end_offsets + pivot) <= diag - pivot - 1;  // monotone in pivot
        const unsigned up = __ballot_sync(kFull, go_up);
        const int n_up = __popc(up);
        const int new_lo = n_up > 0 ? __shfl_sync(kFull, pivot, n_up - 1) + 1 : lo;
        const int new_hi = n_up < 32 ? __shfl_sync(kFull, pivot, n_up) : hi;
        lo = new_lo;
        hi = new_hi;
    }
    return make_int2(min(lo, num_rows), diag - lo);
}

// The same search, first probing the 32 consecutive rows around lo + guess (the previous tile's row
// count): regular matrices resolve in that one round; the rest falls through to the 32-ary search
// over what the window excluded.
__device__ __forceinline__ int2 warp_merge_path_search_window(int diag, const int* __restrict__ row_end_offsets,
                                                              int num_rows, int lo, int hi, int guess, int lane)
{
    if (hi - lo > 32) {
        int w0 = lo + guess - 15;  // window [w0, w0 + 32) inside [lo, hi)
        w0 = max(lo, min(w0, hi - 32));
        const int pivot = w0 + lane;
        const bool go_up = __ldg(row_end_offsets + pivot) <= diag - pivot - 1;
        const unsigned up = __ballot_sync(kFull, go_up);  // monotone: a prefix of the lanes
        const int n_up = __popc(up);
        if (n_up == 0) hi = w0;
        else if (n_up == 32) lo = w0 + 32;
        else lo = hi = w0 + n_up;
    }
    return warp_merge_path_search_bounded(diag, row_end_offsets, num_rows, lo, hi, lane);
}

// Two coordinates over the whole matrix in the SAME rounds (each lane probes one pivot per diagonal per
// round): the start and the end of a block's first tile cost the dependent L2 trips of one search.
__device__ __forceinline__ void warp_merge_path_search_global2(int64_t diag_a64, int64_t diag_b64,
                                                               const int* __restrict__ row_end_offsets, int num_rows,
                                                               int num_nonzeros, int lane, int2& ca, int2& cb)
{
    const int64_t total = (int64_t)num_rows + num_nonzeros;
    const int da = (int)(diag_a64 < total ? diag_a64 : total), db = (int)(diag_b64 < total ? diag_b64 : total);
    int lo_a = max(da - num_nonzeros, 0), hi_a = min(da, num_rows);
    int lo_b = max(db - num_nonzeros, 0), hi_b = min(db, num_rows);
    while (lo_a < hi_a || lo_b < hi_b) {
        int piv_a = lo_a + (int)(((int64_t)(hi_a - lo_a) * (lane + 1)) / 33);
        int piv_b = lo_b + (int)(((int64_t)(hi_b - lo_b) * (lane + 1)) / 33);
        piv_a = min(piv_a, max(hi_a - 1, lo_a));
        piv_b = min(piv_b, max(hi_b - 1, lo_b));
        // a finished search (lo == hi) keeps probing a harmless valid row (or none when the matrix is empty)
        const bool act_a = lo_a < hi_a, act_b = lo_b < hi_b;
        const int va = act_a ? __ldg(row_end_offsets + piv_a) : 0;
        const int vb = act_b ? __ldg(row_end_offsets + piv_b) : 0;
        const unsigned up_a = __ballot_sync(kFull, act_a && va <= da - piv_a - 1);
        const unsigned up_b = __ballot_sync(kFull, act_b && vb <= db - piv_b - 1);
        const int na = __popc(up_a), nb = __popc(up_b);
        const int a_lo = na > 0 ? __shfl_sync(kFull, piv_a, na - 1) + 1 : lo_a;
        const int a_hi = na < 32 ? __shfl_sync(kFull, piv_a, na) : hi_a;
        const int b_lo = nb > 0 ? __shfl_sync(kFull, piv_b, nb - 1) + 1 : lo_b;
        const int b_hi = nb < 32 ? __shfl_sync(kFull, piv_b, nb) : hi_b;
        if (act_a) lo_a = a_lo, hi_a = a_hi;
        if (act_b) lo_b = b_lo, hi_b = b_hi;
    }
    ca = make_int2(min(lo_a, num_rows), da - lo_a);
    cb = make_int2(min(lo_b, num_rows), db - lo_b);
}

template <class C, bool AXPBY, bool SEARCH>
__global__ __launch_bounds__(C::THREADS) void spmv_pipe_kernel(
    const typename C::value_type* __restrict__ values, const int* __restrict__ row_offsets,
    const int* __restrict__ column_indices, const typename C::value_type* __restrict__ x,
    typename C::value_type* __restrict__ y,
    const int2* __restrict__ coords_in,  // !SEARCH: tile coordinates from tile_search_kernel
    int2* __restrict__ coords_out,       // optional export of the coordinates the producer found
    int* __restrict__ carry_rows, typename C::value_type* __restrict__ carry_vals, unsigned int* __restrict__ ticket,
    typename C::value_type alpha, typename C::value_type beta, int num_rows, int num_nonzeros, int num_tiles,
    int shift_v, int shift_c, int shift_r)
{
    using T = typename C::value_type;
    constexpr int IPT = C::IPT, NW = C::NW, STAGES = C::STAGES, CSTAGES = C::CSTAGES, GV = C::GV;
    MSPMV_DYNAMIC_SHARED(smem_raw);
    PipeCtl<C>& ctl = *reinterpret_cast<PipeCtl<C>*>(smem_raw);
    PipeStage<C>* stages = reinterpret_cast<PipeStage<C>*>(smem_raw + sizeof(PipeCtl<C>));
    PipeColStage<C>* cstages =
        reinterpret_cast<PipeColStage<C>*>(smem_raw + sizeof(PipeCtl<C>) + sizeof(PipeStage<C>) * STAGES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x;
    const int t0 = (int)(((int64_t)blockIdx.x * num_tiles) / G);
    const int t1 = (int)(((int64_t)(blockIdx.x + 1) * num_tiles) / G);
    const int n = t1 - t0;  // >= 1: the host launches at most num_tiles blocks

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&ctl.full[s], 1);
            mbar_init(&ctl.empty[s], NW);
        }
#pragma unroll
        for (int s = 0; s < CSTAGES; ++s) {
            mbar_init(&ctl.full_c[s], 1);
            mbar_init(&ctl.empty_c[s], NW);
        }
        fence_mbar_init();
    }
    for (int i = tid; i < 2 * C::BW; i += C::THREADS) (&ctl.bits[0][0])[i] = 0u;
    __syncthreads();

    // =============================== producer warp ================================================
    if (warp == NW) {
        const uint64_t policy = l2_policy_evict_first();
        const int* row_end = row_offsets + 1;  // device_spmv.cuh:148
        const int64_t total = (int64_t)num_rows + num_nonzeros;
        // end coordinate of tile k of the run, given its start coordinate
        auto next_coord = [&](int k, const int2 c, int guess) -> int2 {
            if (!SEARCH) return __ldg(coords_in + t0 + k + 1);
            const int64_t d64 = (int64_t)(t0 + k + 1) * C::TILE;
            const int d1 = (int)(d64 < total ? d64 : total);
            // the path is monotone: x1 in [x0, x0 + (d1 - d0)] = [x0, d1 - y0]
            return warp_merge_path_search_window(d1, row_end, num_rows, max(c.x, d1 - num_nonzeros),
                                                 min(d1 - c.y, num_rows), guess, lane);
        };
        int2 c0, c1;
        if (SEARCH) {
            warp_merge_path_search_global2((int64_t)t0 * C::TILE, (int64_t)(t0 + 1) * C::TILE, row_end, num_rows,
                                           num_nonzeros, lane, c0, c1);
        } else {
            c0 = __ldg(coords_in + t0);
            c1 = next_coord(0, c0, 0);
        }
        for (int i = 0; i < n; ++i) {
            const int s = i % STAGES, sc = i % CSTAGES;
            if (coords_out != nullptr && lane == 0) {
                coords_out[t0 + i] = c0;
                if (t0 + i + 1 == num_tiles) coords_out[num_tiles] = c1;
            }
            const int x0 = c0.x, y0 = c0.y, nrows = c1.x - c0.x, nnzs = c1.y - c0.y;
            // column indices first: the consumers need them earliest
            if (i >= CSTAGES) mbar_wait(&ctl.empty_c[sc], (uint32_t)((i / CSTAGES) - 1) & 1u);
            const int base_c = (y0 + shift_c) & ~3;
            uint32_t b = stage_superset<int>(column_indices, shift_c, y0, y0 + nnzs, num_nonzeros, cstages[sc].col,
                                             base_c, &ctl.full_c[sc], policy, lane);
            if (lane == 0) ctl.coord[sc] = make_int4(c0.x, c0.y, c1.x, c1.y);
            __syncwarp();
            if (lane == 0) {
                if (b) mbar_arrive_expect_tx(&ctl.full_c[sc], b);
                else mbar_arrive(&ctl.full_c[sc]);
            }
            // the next tile's end coordinate, searched while the value slot is still busy
            int2 c2 = c1;
            if (i + 1 < n) c2 = next_coord(i + 1, c1, nrows);
            if (i >= STAGES) mbar_wait(&ctl.empty[s], (uint32_t)((i / STAGES) - 1) & 1u);
            PipeStage<C>& st = stages[s];
            const int base_v = (y0 + shift_v) & ~(GV - 1);
            const int jr0 = x0 + 1;  // row_end_offsets[x0 + r] == row_offsets[jr0 + r]
            const int base_r = (jr0 + shift_r) & ~3;
            b = stage_superset<T>(values, shift_v, y0, y0 + nnzs, num_nonzeros, st.val, base_v, &ctl.full[s], policy,
                                  lane);
            if (nrows <= C::ROWCAP)
                b += stage_superset<int>(row_offsets, shift_r, jr0, jr0 + nrows, num_rows + 1, st.row, base_r,
                                         &ctl.full[s], policy, lane);
            __syncwarp();
            if (lane == 0) {
                if (b) mbar_arrive_expect_tx(&ctl.full[s], b);
                else mbar_arrive(&ctl.full[s]);
            }
            c0 = c1;
            c1 = c2;
        }
        return;
    }

    // =============================== consumer warps ===============================================
    // P1: row owners flag the slot in front of which their row ends
    auto mark_rows = [&](const int4 c, const PipeStage<C>& st, uint32_t* bits) {
        const int nrows = c.z - c.x, jr0 = c.x + 1;
        const int off_r = (jr0 + shift_r) & 3;
        for (int r = tid; r < nrows; r += C::CONSUMERS) {
            const int e = nrows <= C::ROWCAP ? st.row[off_r + r] : __ldg(row_offsets + jr0 + r);
            const int pos = e - c.y;
            atomicOr(&bits[pos >> 5], 1u << (pos & 31));
        }
    };

    const uint64_t keep = l2_policy_evict_last();  // x is the only reused data: keep it in L2
    const int base = tid * IPT;  // my first slot

    // G: read my column indices of tile k and issue the x gathers (consumed by the walk: with
    // GATHER_AHEAD one step later, the loads fly while the previous tile is walked and stored).
    auto gather = [&](int k, const int4 c, T(&xo)[IPT]) {
        const int sc = k % CSTAGES;
        const int nnzs = c.w - c.y;
        const int n_mine = min(max(nnzs - base, 0), IPT);  // my slots that hold a nonzero
        const int* pc = cstages[sc].col + (((c.y + shift_c) & 3) + base);
        int cidx[IPT];
#pragma unroll
        for (int j = 0; j < IPT; ++j) cidx[j] = j < n_mine ? pc[j] : -1;
        // L1 policy per warp by column span (profiles/tuning_r01.txt): narrow span = lines are re-used
        int cmin = cidx[0] >= 0 ? cidx[0] : INT_MAX, cmax = cidx[0];
        if (cidx[IPT - 1] >= 0) {
            cmin = min(cmin, cidx[IPT - 1]);
            cmax = max(cmax, cidx[IPT - 1]);
        }
        cmin = __reduce_min_sync(kFull, cmin);
        cmax = __reduce_max_sync(kFull, cmax);
        if (cmax - cmin < C::LOCAL_SPAN) {
#pragma unroll
            for (int j = 0; j < IPT; ++j) xo[j] = cidx[j] >= 0 ? ld_gather_l1(x + cidx[j], keep) : T(0);
        } else {
#pragma unroll
            for (int j = 0; j < IPT; ++j) xo[j] = cidx[j] >= 0 ? ld_gather(x + cidx[j], keep) : T(0);
        }
        __syncwarp();  // every lane's indices have been consumed by its gathers: hand the slot back
        if (lane == 0) mbar_arrive(&ctl.empty_c[sc]);
    };

    Seg<T> carry;  // the row that continues from this block's previous tile
    carry.val = T(0);
    carry.ended = 0;
    int4 cur, nxt;
    T xa[IPT], xb[IPT];
    mbar_wait(&ctl.full_c[0], 0);
    cur = ctl.coord[0];
    if (C::AHEAD) gather(0, cur, xa);
    mbar_wait(&ctl.full[0], 0);
    mark_rows(cur, stages[0], ctl.bits[0]);
    nxt = cur;
    named_bar_sync(2, C::CONSUMERS);

    // one tile: xc = its x values (AHEAD: gathered one step ago), xn = where the next tile's go
#if MSPMV_PIPE_PROFILE
    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_last = prof_clock();
#endif
    auto step = [&](int i, T(&xc)[IPT], T(&xn)[IPT]) {
        const int s = i % STAGES, bsel = i & 1;
        PIPE_PROF_MARK(7);  // loop overhead / between tiles
        PipeStage<C>& st = stages[s];
        const int x0 = cur.x, y0 = cur.y, nrows = cur.z - cur.x, nnzs = cur.w - cur.y;
        const int off_v = (y0 + shift_v) & (GV - 1);
        const int n_mine = min(max(nnzs - base, 0), IPT);
        T* pv = st.val + (off_v + base);
        // VALS_FIRST: my flag words and values are read BEFORE my gathers are issued -- the LSU serves a warp's
        // shared-memory reads and its 32-sector gathers from one queue, in order, so reads issued after the
        // gathers would wait behind them (profiles/gather_ceiling_r02.txt).
        uint32_t w0 = 0u, w1 = 0u;
        T vv[IPT];
        if (C::VALS_FIRST) {
            w0 = ctl.bits[bsel][base >> 5];
            w1 = ctl.bits[bsel][(base >> 5) + 1];
#pragma unroll
            for (int j = 0; j < IPT; ++j) vv[j] = j < n_mine ? pv[j] : T(0);
        }
        if (C::AHEAD) {
            if (i + 1 < n) {
                const int sc1 = (i + 1) % CSTAGES;
                mbar_wait(&ctl.full_c[sc1], (uint32_t)((i + 1) / CSTAGES) & 1u);
                nxt = ctl.coord[sc1];
                gather(i + 1, nxt, xn);
            }
        } else {
            gather(i, cur, xc);
        }

        PIPE_PROF_MARK(0);  // shared-memory reads issued, column indices read, gathers issued
        // ---- W: walk my slots (cpu_spmv.cpp:324-340) ------------------------------------------------
        if (!C::VALS_FIRST) {
            w0 = ctl.bits[bsel][base >> 5];
            w1 = ctl.bits[bsel][(base >> 5) + 1];
        }
        const uint32_t bits = __funnelshift_r(w0, w1, base & 31) & ((1u << IPT) - 1u);
        // FIRST_IN_REG: the sum of my first finished segment stays in a register -- it still lacks the partial
        // that precedes me, which only the scan knows; later segments are complete and are parked right away.
        const uint32_t first_bit = C::FIRST_IN_REG ? (bits & (0u - bits)) : 0u;  // lowest set flag
        T first_sum = T(0);
        T running = T(0);
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            T v;
            if (C::VALS_FIRST) v = vv[j];
            else v = j < n_mine ? pv[j] : T(0);
            if ((bits >> j) & 1u) {  // a row ends in front of slot j: park its sum, start over
                if (C::FIRST_IN_REG && ((first_bit >> j) & 1u)) first_sum = running;
                else pv[j] = running;
                running = T(0);
            }
            running = fma(v, xc[j], running);
        }
        PIPE_PROF_MARK(1);  // walk (waits for the gathered x values)
        // ---- S: block-wide segmented scan of (had a boundary, tail sum) ----------------------------
        Seg<T> elem, excl, total;
        elem.val = running;
        elem.ended = bits != 0u;
        block_seg_scan_exclusive<T, NW, true>(elem, carry, ctl.warp, tid, 1, excl, total);
        if (bits != 0u) {  // my first segment, now complete
            if (C::FIRST_IN_REG) pv[__ffs((int)bits) - 1] = first_sum + excl.val;
            else pv[__ffs((int)bits) - 1] += excl.val;
        }
        if (tid < C::BW) ctl.bits[bsel][tid] = 0u;  // everybody read its bits before the scan's barrier
        carry.val = total.val;

        PIPE_PROF_MARK(2);  // scan (one named barrier inside)
        // ---- P1 of the next tile, then the barrier that publishes the parked sums -----------------
        if (i + 1 < n) {
            if (!C::AHEAD) {
                const int sc1 = (i + 1) % CSTAGES;
                mbar_wait(&ctl.full_c[sc1], (uint32_t)((i + 1) / CSTAGES) & 1u);
                nxt = ctl.coord[sc1];
            }
            const int s1 = (i + 1) % STAGES;
            mbar_wait(&ctl.full[s1], (uint32_t)((i + 1) / STAGES) & 1u);
            mark_rows(nxt, stages[s1], ctl.bits[bsel ^ 1]);
        }
        PIPE_PROF_MARK(3);  // waits for the next tile's data + P1
        named_bar_sync(2, C::CONSUMERS);
        PIPE_PROF_MARK(4);  // barrier

        // ---- Y: row owners store y ----------------------------------------------------------------
        {
            const int jr0 = x0 + 1;
            const int off_r = (jr0 + shift_r) & 3;
            for (int r = tid; r < nrows; r += C::CONSUMERS) {
                int e, ep;
                if (nrows <= C::ROWCAP) {
                    e = st.row[off_r + r];
                    ep = r > 0 ? st.row[off_r + r - 1] : e - 1;
                } else {
                    e = __ldg(row_offsets + jr0 + r);
                    ep = r > 0 ? __ldg(row_offsets + jr0 + r - 1) : e - 1;
                }
                const T sum = ep != e ? st.val[off_v + (e - y0)] : T(0);  // ep == e: empty row
                y[x0 + r] = epilogue<T, AXPBY>(sum, alpha, beta, y + x0 + r);
            }
        }
        // hand the stage back: my generic-proxy writes into the slot are ordered before the producer's next bulk copy
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl.empty[s]);
        cur = nxt;
        PIPE_PROF_MARK(5);  // Y + release
    };
    if (C::AHEAD) {
        for (int i = 0; i < n; i += 2) {  // ping-pong register sets: no copy waits for the gathers in flight
            step(i, xa, xb);
            if (i + 1 < n) step(i + 1, xb, xa);
        }
    } else {
        for (int i = 0; i < n; ++i) step(i, xa, xa);
    }

#if MSPMV_PIPE_PROFILE
    if (tid == 0) {
        for (int k = 0; k < 8; ++k) atomicAdd(&g_pipe_prof[k], (unsigned long long)prof_acc[k]);
        atomicAdd(&g_pipe_prof[6], (unsigned long long)n);  // tiles (slot 6 is not a phase)
    }
#endif
    // ---- the run's carry-out; the last block to finish folds all of them (cpu_spmv.cpp:348-352) ---
    if (tid == 0) {
        carry_rows[blockIdx.x] = cur.z;  // may equal num_rows: dropped by the fold
        carry_vals[blockIdx.x] = carry.val;
    }
    if (G == 1) return;  // dispatch_spmv_orig.cuh:721
    __threadfence();
    named_bar_sync(2, C::CONSUMERS);
    if (tid == 0) ctl.last = atomicAdd(ticket, 1u) == (unsigned)G - 1u;
    named_bar_sync(2, C::CONSUMERS);
    if (!ctl.last) return;
    __threadfence();
    Seg<T> run;  // the run of equal rows that touches the end of the previous chunk
    run.val = T(0);
    run.ended = 0;
    for (int cb = 0; cb < G; cb += C::CONSUMERS) {
        const int i = cb + tid;
        const int row = i < G ? carry_rows[i] : INT_MAX;
        const int prev_row = (i > 0 && i < G) ? carry_rows[i - 1] : -1;
        const int next_row = i + 1 < G ? carry_rows[i + 1] : INT_MAX;
        Seg<T> e, excl, total;
        e.val = i < G ? carry_vals[i] : T(0);
        e.ended = row != prev_row;  // a new run starts here
        block_seg_scan_exclusive<T, NW, false>(e, run, ctl.warp, tid, 1, excl, total);
        const T run_sum = e.ended ? e.val : excl.val + e.val;  // my run up to and including me
        if (i < G && row != next_row && row < num_rows) y[row] += AXPBY ? alpha * run_sum : run_sum;
        run.val = total.val;
        run.ended = 0;
        named_bar_sync(2, C::CONSUMERS);  // ctl.warp is reused by the next chunk
    }
}

}  // namespace mspmv
