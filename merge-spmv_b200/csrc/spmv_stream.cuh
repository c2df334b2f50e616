// spmv_stream.cuh -- "stream" engine: persistent, TMA-fed merge-based CsrMV for sm_100a.
//
// Decomposition (the reference's, applied recursively -- README.md:5, paper section III.A):
//   GPU  -> G threadblocks: block c owns the contiguous diagonal swath [c*S, (c+1)*S) of the
//           (row_end_offsets (+) N_nnz) merge path, S = ceil((rows+nnz)/G) -- exactly thread c
//           of OmpMergeCsrmv with p = G (cpu_spmv.cpp:311-321).  G = CTAs resident on the GPU.
//   block -> tiles of TILE = 256*IPT merge items walked in order; the partial row that crosses a
//           tile boundary is carried in registers, so only ONE carry per block reaches global
//           memory (the reference emits one per tile, agent_spmv_orig.cuh:906-913).
//   tile -> 256 threads x IPT consecutive merge items each.  Instead of one MergePathSearch per
//           thread (agent_spmv_orig.cuh:539-545) the tile's row ends are scattered once into a
//           shared-memory bitmap (bit p set <=> merge item p is a row end, p = row_end[r] - y0 +
//           r - x0); a thread's start coordinate is then a popcount prefix -- the same unique
//           coordinate MergePathSearch returns -- and its serial walk (:557-578) is a branch-free
//           unrolled loop driven by its IPT flag bits.  Partial rows are closed by a warp-shuffle
//           segmented scan; finished rows are written to y straight from registers.
//
// Three warp-specialised pipeline stages per block, decoupled by mbarriers over a deep shared-
// memory ring (chunks of the *absolute* nonzero index space, so every bulk copy is 16-byte
// aligned no matter where the swath starts; ragged ends of the first/last chunk and arrays whose
// base is not 16-byte aligned are patched with a handful of scalar copies):
//   1. producer warp   cp.async.bulk (TMA, SASS UBLKCP) of values / column indices / row offsets
//                      -> full_raw[chunk]
//   2. gather warps    wait full_raw; LDS.128 column indices, then one 4/8-byte cp.async
//                      (LDGSTS) per nonzero copies x[col] straight into an x ring -- no
//                      registers are held while the gather is in flight, so thousands of L2
//                      requests per SM stay outstanding -> full_x[chunk] when they have landed.
//                      (fp32: x[col] overwrites col in place; fp64 has a separate ring.)
//   3. reduce warps    wait full_x; bitmap / walk with fma(value, x, sum) / scan / y stores
//                      -> empty[chunk]
#pragma once

#include <limits.h>

#include "merge_common.cuh"
#include "tma_stage.cuh"

namespace mspmv {

// ------------------------------------------------------------------------------------------------
// Geometry
// ------------------------------------------------------------------------------------------------
template <typename T>
struct StreamCfg {
    static constexpr int REDUCERS = 128;                       // 4 reduce warps
    static constexpr int GATHERERS = 128;                      // 4 gather warps
    static constexpr int THREADS = REDUCERS + GATHERERS + 64;  // + 2 producer warps (nonzeros, row offsets)
    static constexpr int IPT = 9;                              // odd: conflict-free strided smem walk
    static constexpr int TILE = REDUCERS * IPT;
    static constexpr int CH = GATHERERS * 4;                   // nonzeros per ring chunk: one aligned group of 4 per gather thread
    static constexpr int NSLOT = sizeof(T) == 8 ? 8 : 16;      // chunks in the nonzero ring
    static constexpr int RN = CH * NSLOT;                      // ring capacity (power of two)
    static constexpr int RCH = 256;                            // row offsets per ring chunk
    static constexpr int RSLOT = 8;
    static constexpr int RR = RCH * RSLOT;
    static constexpr int BW = TILE / 32 + 2;                   // words per row-end bitmap
    static constexpr int CTAS_PER_SM = sizeof(T) == 8 ? 2 : 3;
    static constexpr int MIN_SWATH = 1024;                     // merge items; small inputs use fewer blocks
    static_assert((RN & (RN - 1)) == 0 && (RR & (RR - 1)) == 0, "rings are power-of-two sized");
    static_assert((NSLOT & (NSLOT - 1)) == 0 && (RSLOT & (RSLOT - 1)) == 0, "slot counts are powers of two");
    static_assert(RN >= TILE + 3 * CH && RR >= TILE + 1 + 2 * RCH, "ring must hold one tile plus slack");
    static_assert(TILE % 32 == 0 && BW <= 96 && (IPT & 1) == 1 && IPT < 32, "bitmap layout");
};

template <typename T>
struct StreamSmem {
    using C = StreamCfg<T>;
    alignas(128) T val[C::RN];        // values
    alignas(128) int col[C::RN];      // column indices; fp32: overwritten in place by x[col]
    alignas(128) T xg[sizeof(T) == 8 ? C::RN : 4];  // fp64: gathered x[col]
    alignas(128) int row[C::RR];      // row_offsets entries (index j = row + 1)
    alignas(16) uint32_t bits[2][C::BW];  // row-end flags per merge item, double-buffered
    alignas(16) Seg<T> warp[C::REDUCERS / 32];
    alignas(8) uint64_t full_raw[C::NSLOT];   // TMA landed
    uint64_t full_x[C::NSLOT];                // x[col] gathers of the chunk have landed
    uint64_t empty_n[C::NSLOT];               // reduce warps are past this chunk
    uint64_t full_r[C::RSLOT];
    uint64_t empty_r[C::RSLOT];
    int2 swath[2];                    // start / end coordinate of this block's swath
};

struct StreamGeom {
    int num_swaths = 0;
    int swath_items = 0;
    int threads = 0;
    int tile_items = 0;
    size_t smem_bytes = 0;
};

template <typename T>
static StreamGeom stream_geometry(int64_t merge_items, int sm_count)
{
    using C = StreamCfg<T>;
    StreamGeom g;
    int64_t want = (merge_items + C::MIN_SWATH - 1) / C::MIN_SWATH;
    int64_t cap = (int64_t)sm_count * C::CTAS_PER_SM;
    int64_t n = want < 1 ? 1 : (want > cap ? cap : want);
    g.num_swaths = (int)n;
    g.swath_items = (int)((merge_items + n - 1) / n);
    g.threads = C::THREADS;
    g.tile_items = C::TILE;
    g.smem_bytes = sizeof(StreamSmem<T>) + 128;
    return g;
}

// ------------------------------------------------------------------------------------------------
// The kernel.  VEC: values/column_indices bases are 16-byte aligned, so the gather stage works on
// aligned groups of four nonzeros with 128-bit shared-memory accesses.
// ------------------------------------------------------------------------------------------------
template <typename T, bool AXPBY, bool VEC>
__global__ __launch_bounds__(StreamCfg<T>::THREADS, StreamCfg<T>::CTAS_PER_SM) void spmv_stream_kernel(
    const T* __restrict__ values, const int* __restrict__ row_offsets,
    const int* __restrict__ column_indices, const T* __restrict__ x, T* __restrict__ y, int num_rows,
    int num_nonzeros, int swath_items, int2* __restrict__ swath_coords, int* __restrict__ carry_rows,
    T* __restrict__ carry_vals, T alpha, T beta, int shift_v, int shift_c, int shift_r)
{
    using C = StreamCfg<T>;
    constexpr int NRW = C::REDUCERS / 32;   // reduce warps: 0 .. NRW-1
    constexpr int NGW = C::GATHERERS / 32;  // gather warps: NRW .. NRW+NGW-1; producer: NRW+NGW
    constexpr int IPT = C::IPT;
    MSPMV_DYNAMIC_SHARED(smem_raw);
    StreamSmem<T>& sm =
        *reinterpret_cast<StreamSmem<T>*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int64_t total_items = (int64_t)num_rows + num_nonzeros;
    const int64_t d_begin64 = (int64_t)blockIdx.x * swath_items;
    const int d_begin = (int)(d_begin64 < total_items ? d_begin64 : total_items);
    const int64_t d_end64 = d_begin64 + swath_items;
    const int d_end = (int)(d_end64 < total_items ? d_end64 : total_items);
    const int* __restrict__ row_end_offsets = row_offsets + 1;

    // ---- prologue: barriers, bitmaps, the two swath boundary searches (warps 0 and 1) ----------
    if (tid == C::THREADS - 32) {
        for (int i = 0; i < C::NSLOT; ++i) {
            mbar_init(&sm.full_raw[i], 1);
            mbar_init(&sm.full_x[i], C::GATHERERS);
            mbar_init(&sm.empty_n[i], 1);
        }
        for (int i = 0; i < C::RSLOT; ++i) {
            mbar_init(&sm.full_r[i], 1);
            mbar_init(&sm.empty_r[i], 1);
        }
        fence_mbar_init();
    }
    for (int i = tid; i < 2 * C::BW; i += C::THREADS) (&sm.bits[0][0])[i] = 0u;
    if (warp < 2) {
        int2 c = warp_merge_path_search_global(warp == 0 ? d_begin : d_end, row_end_offsets, num_rows,
                                               num_nonzeros, lane);
        if (lane == 0) sm.swath[warp] = c;
    }
    __syncthreads();
    const int X0 = sm.swath[0].x, Y0 = sm.swath[0].y;
    const int X1 = sm.swath[1].x, Y1 = sm.swath[1].y;
    if (tid == 0) {
        swath_coords[blockIdx.x] = sm.swath[0];
        if (blockIdx.x == gridDim.x - 1) swath_coords[gridDim.x] = sm.swath[1];
    }

    // Chunk index ranges.  Nonzeros: absolute index i lives in chunk i / CH.  Row offsets are
    // addressed by j = row + 1 (row_end_offsets[row] == row_offsets[j]); the swath needs
    // j in [X0 + 1, X1 + 1).
    const int kn_lo = Y0 / C::CH;
    const int kn_hi = Y1 > Y0 ? (Y1 - 1) / C::CH + 1 : kn_lo;
    const int J0 = X0 + 1, J1 = X1 + 1;
    const int kr_lo = J0 / C::RCH;
    const int kr_hi = J1 > J0 ? (J1 - 1) / C::RCH + 1 : kr_lo;

    if (warp == NRW + NGW) {
        // ======================= stage 1a: producer warp, values + column indices ==============
        // Interior chunks of 16-byte-aligned arrays are two bulk copies issued by lane 0 with a
        // handful of instructions; the (possibly ragged) first and last chunk of the swath, and
        // arrays with misaligned bases, take the general whole-warp path.
        const uint64_t policy = l2_policy_evict_first();
        const bool aligned = (shift_v | shift_c) == 0;
        constexpr uint32_t kChunkBytes = C::CH * (uint32_t)(sizeof(T) + sizeof(int));
        for (int kn = kn_lo; kn < kn_hi; ++kn) {
            const int kk = kn - kn_lo, slot = kk & (C::NSLOT - 1), use = kk / C::NSLOT;
            if (aligned && kn > kn_lo && kn + 1 < kn_hi) {
                if (lane == 0) {
                    if (use > 0) mbar_wait(&sm.empty_n[slot], (use - 1) & 1);
                    const int pos = (kn & (C::NSLOT - 1)) * C::CH;
                    mbar_arrive_expect_tx(&sm.full_raw[slot], kChunkBytes);
                    bulk_g2s(&sm.val[pos], values + (size_t)kn * C::CH, C::CH * (uint32_t)sizeof(T),
                             &sm.full_raw[slot], policy);
                    bulk_g2s(&sm.col[pos], column_indices + (size_t)kn * C::CH, C::CH * (uint32_t)sizeof(int),
                             &sm.full_raw[slot], policy);
                }
            } else {
                __syncwarp();
                if (use > 0) mbar_wait(&sm.empty_n[slot], (use - 1) & 1);
                int lo = max(kn * C::CH, Y0), hi = (int)min((int64_t)(kn + 1) * C::CH, (int64_t)Y1);
                uint32_t b = stage_range<T>(values, shift_v, lo, hi, sm.val, C::RN - 1, &sm.full_raw[slot],
                                            policy, lane);
                b += stage_range<int>(column_indices, shift_c, lo, hi, sm.col, C::RN - 1, &sm.full_raw[slot],
                                      policy, lane);
                __syncwarp();
                if (lane == 0) {
                    if (b) mbar_arrive_expect_tx(&sm.full_raw[slot], b);
                    else mbar_arrive(&sm.full_raw[slot]);
                }
            }
        }
        return;
    }
    if (warp == NRW + NGW + 1) {
        // ======================= stage 1b: producer warp, row offsets ===========================
        const uint64_t policy = l2_policy_evict_first();
        for (int kr = kr_lo; kr < kr_hi; ++kr) {
            const int kk = kr - kr_lo, slot = kk & (C::RSLOT - 1), use = kk / C::RSLOT;
            if (use > 0) mbar_wait(&sm.empty_r[slot], (use - 1) & 1);
            int lo = max(kr * C::RCH, J0), hi = (int)min((int64_t)(kr + 1) * C::RCH, (int64_t)J1);
            uint32_t b = stage_range<int>(row_offsets, shift_r, lo, hi, sm.row, C::RR - 1, &sm.full_r[slot],
                                          policy, lane);
            __syncwarp();
            if (lane == 0) {
                if (b) mbar_arrive_expect_tx(&sm.full_r[slot], b);
                else mbar_arrive(&sm.full_r[slot]);
            }
        }
        return;
    }

    // where the gathered x of nonzero j lives: fp64 -> xg ring, fp32 -> in place of its column index
    auto x_slot = [&](int j) -> T* {
        if (sizeof(T) == 8) return &sm.xg[j & (C::RN - 1)];
        return reinterpret_cast<T*>(&sm.col[(j + shift_c) & (C::RN - 1)]);
    };

    if (warp >= NRW) {
        // =================================== stage 2: gather warps (NRW .. NRW+NGW-1) ==========
        // x[col] of every nonzero of the chunk by asynchronous 4/8-byte copies
        // (the gather of agent_spmv_orig.cuh:472-494, without the register round trip)
        const int gt = tid - C::REDUCERS;  // 0 .. GATHERERS-1
        for (int k = kn_lo; k < kn_hi; ++k) {
            const int kk = k - kn_lo, slot = kk & (C::NSLOT - 1);
            mbar_wait(&sm.full_raw[slot], (kk / C::NSLOT) & 1);
            if (VEC) {
                const int j = k * C::CH + 4 * gt;
                const int4 c = *reinterpret_cast<const int4*>(&sm.col[j & (C::RN - 1)]);
                if (j >= Y0 && j + 4 <= Y1) {
                    cp_async_gather<sizeof(T)>(x_slot(j + 0), x + c.x);
                    cp_async_gather<sizeof(T)>(x_slot(j + 1), x + c.y);
                    cp_async_gather<sizeof(T)>(x_slot(j + 2), x + c.z);
                    cp_async_gather<sizeof(T)>(x_slot(j + 3), x + c.w);
                } else {  // swath edge: slots outside [Y0, Y1) hold no data
                    if (j + 0 >= Y0 && j + 0 < Y1) cp_async_gather<sizeof(T)>(x_slot(j + 0), x + c.x);
                    if (j + 1 >= Y0 && j + 1 < Y1) cp_async_gather<sizeof(T)>(x_slot(j + 1), x + c.y);
                    if (j + 2 >= Y0 && j + 2 < Y1) cp_async_gather<sizeof(T)>(x_slot(j + 2), x + c.z);
                    if (j + 3 >= Y0 && j + 3 < Y1) cp_async_gather<sizeof(T)>(x_slot(j + 3), x + c.w);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = k * C::CH + gt + e * C::GATHERERS;
                    if (j >= Y0 && j < Y1) {
                        const int c = sm.col[(j + shift_c) & (C::RN - 1)];
                        cp_async_gather<sizeof(T)>(x_slot(j), x + c);
                    }
                }
            }
            cp_async_arrive_noinc(&sm.full_x[slot]);
        }
        return;
    }

    // ===================================== stage 3: reduce warps ===============================
    int x0 = X0, y0 = Y0;                            // current tile start coordinate
    int d = d_begin;                                 // current diagonal
    int n_waited = 0, r_waited = 0;                  // chunks (relative index) already acquired
    int n_released = 0, r_released = 0;
    int buf = 0;
    Seg<T> carry;                                    // partial row carried across tiles, in registers
    carry.val = T(0);
    carry.ended = 0;

    while (d < d_end) {
        const int items = min(C::TILE, d_end - d);
        const int y_need = min(y0 + items, Y1);      // exclusive bound on nonzeros this tile can touch
        const int nrows_max = min(items, X1 - x0);   // row ends this tile can contain

        // ---- acquire the ring chunks this tile can touch ---------------------------------------
        if (nrows_max > 0) {
            int kk_need = (x0 + nrows_max) / C::RCH - kr_lo;  // j = x0 + nrows_max is the last needed
            while (r_waited <= kk_need) {
                mbar_wait(&sm.full_r[r_waited % C::RSLOT], (r_waited / C::RSLOT) & 1);
                ++r_waited;
            }
        }

        // ---- row-end flags: merge item p = row_end[r] - y0 + (r - x0) is the end of row r --------
        uint32_t* bits_w = sm.bits[buf];
        for (int r = tid; r < nrows_max; r += C::REDUCERS) {
            int e = sm.row[(x0 + r + 1 + shift_r) & (C::RR - 1)];
            int pos = e - y0 + r;
            if (pos >= items) break;                 // positions increase with r
            atomicOr(&bits_w[pos >> 5], 1u << (pos & 31));
        }
        if (y_need > Y0) {
            int kk_need = (y_need - 1) / C::CH - kn_lo;
            while (n_waited <= kk_need) {
                mbar_wait(&sm.full_x[n_waited % C::NSLOT], (n_waited / C::NSLOT) & 1);
                ++n_waited;
            }
        }
        named_bar_sync(1, C::REDUCERS);

        // clear the other bitmap for the next tile (its last readers finished before the barrier)
        if (tid < C::BW) sm.bits[buf ^ 1][tid] = 0u;

        // my IPT flag bits and the number of row ends before my first item (= my start row offset)
        const int diag = tid * IPT;
        const uint32_t w0 = bits_w[diag >> 5], w1 = bits_w[(diag >> 5) + 1];
        const uint32_t bits = __funnelshift_r(w0, w1, diag & 31) & ((1u << IPT) - 1u);
        const int cnt = __popc(bits);
        // row ends before my first item: whole words of earlier warps (one REDUX), whole words of my
        // warp before mine (independent loads), and the low part of my first word
        int before_warp = 0, all = 0;
#pragma unroll
        for (int k = lane; k < C::BW; k += 32) {
            int pc = __popc(bits_w[k]);
            all += pc;
            if (k < warp * IPT) before_warp += pc;   // warp w owns bitmap words [w*IPT, (w+1)*IPT)
        }
        before_warp = __reduce_add_sync(kFull, before_warp);
        const int nrows = __reduce_add_sync(kFull, all);
        int in_warp = __popc(w0 & ((1u << (diag & 31)) - 1u));
#pragma unroll
        for (int k = 0; k < IPT - 1; ++k)
            if (warp * IPT + k < (diag >> 5)) in_warp += __popc(bits_w[warp * IPT + k]);
        const int xs = before_warp + in_warp;        // row ends before my span == my start row - x0

        // serial walk over my IPT merge items (cpu_spmv.cpp:324-340; agent_spmv_orig.cuh:557-578):
        // flag bit -> the row ends here, else consume the next product.  Loads depend only on the
        // flag bits, so they all issue up front.
        const int my_items = min(max(items - diag, 0), IPT);
        int ny = y0 + diag - xs;                     // my first nonzero (absolute index)
        T sums[IPT];
        T running = T(0);
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const bool is_end = (bits >> i) & 1u;
            if (!is_end && i < my_items) running = fma(sm.val[(ny + shift_v) & (C::RN - 1)], *x_slot(ny), running);
            ny += is_end ? 0 : 1;
            sums[i] = running;
            if (is_end) running = T(0);
        }
        Seg<T> elem;
        elem.val = running;
        elem.ended = cnt > 0;
        Seg<T> excl, total;
        block_seg_scan_exclusive<T, NRW>(elem, carry, sm.warp, tid, 1, excl, total);

        // finished rows straight to y: my k-th row end is row x0 + xs + k; the first one also
        // collects what earlier threads / tiles accumulated for that row
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

        // ---- advance; release ring chunks that are now entirely behind the tile end -----------------
        x0 += nrows;
        y0 += items - nrows;
        d += items;
        carry.val = total.val;
        carry.ended = 0;
        buf ^= 1;
        if (tid == 0) {
            fence_proxy_async();  // ring slots read/written through the generic proxy are re-filled by bulk copies
            while (n_released < n_waited && (int64_t)(kn_lo + n_released + 1) * C::CH <= y0) {
                mbar_arrive(&sm.empty_n[n_released % C::NSLOT]);
                ++n_released;
            }
            // row chunk k holds j in [k*RCH, (k+1)*RCH); rows < x0 are dead, i.e. j <= x0
            while (r_released < r_waited && (int64_t)(kr_lo + r_released + 1) * C::RCH <= x0 + 1) {
                mbar_arrive(&sm.empty_r[r_released % C::RSLOT]);
                ++r_released;
            }
        }
    }

    // swath carry-out (cpu_spmv.cpp:343-344): the row that continues into the next swath
    if (tid == 0) {
        carry_rows[blockIdx.x] = X1;
        carry_vals[blockIdx.x] = carry.val;
    }
}

#ifndef MSPMV_PTX_HEADER  // host-side launch: not part of what the CPU interpreter of tests/emu compiles
template <typename T, bool AXPBY>
static int stream_launch(const StreamGeom& g, const T* values, const int* row_offsets, const int* col,
                         const T* x, T* y, int num_rows, int num_nonzeros, int2* swath_coords,
                         int* carry_rows, T* carry_vals, T alpha, T beta, cudaStream_t stream)
{
    int shift_v = (int)((reinterpret_cast<uintptr_t>(values) & 15) / sizeof(T));
    int shift_c = (int)((reinterpret_cast<uintptr_t>(col) & 15) / sizeof(int));
    int shift_r = (int)((reinterpret_cast<uintptr_t>(row_offsets) & 15) / sizeof(int));
    const bool vec = shift_v == 0 && shift_c == 0;
    auto kernel = vec ? spmv_stream_kernel<T, AXPBY, true> : spmv_stream_kernel<T, AXPBY, false>;
    static bool configured[2][64] = {};  // per template instantiation, variant and device
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (!configured[vec][dev & 63]) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
        if (e != cudaSuccess) return (int)e;
        configured[vec][dev & 63] = true;
    }
    kernel<<<g.num_swaths, g.threads, g.smem_bytes, stream>>>(values, row_offsets, col, x, y, num_rows,
                                                              num_nonzeros, g.swath_items, swath_coords,
                                                              carry_rows, carry_vals, alpha, beta,
                                                              shift_v, shift_c, shift_r);
    return 0;
}
#endif  // MSPMV_PTX_HEADER

}  // namespace mspmv
