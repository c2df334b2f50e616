// mergespmv.cu -- C ABI (include/mergespmv.h) and host-side dispatch of the sm_100a kernels.
//
// Replaces, for the one path this library covers:
//   cub::DeviceSpmv::CsrMV            cub/device/device_spmv.cuh:129-164
//   DispatchSpmv::Dispatch            cub/device/dispatch/dispatch_spmv_orig.cuh:544-752
//   AliasTemporaries                  cub/util_device.cuh:62-103 (256-byte aligned carve-up)
// Unlike the reference's Dispatch there is no per-call cudaGetDevice / attribute / occupancy
// query (dispatch_spmv_orig.cuh:597-629): device properties are cached per device on first use.
#include <cuda_runtime.h>

#include <atomic>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <vector>

#include "../../include/mergespmv.h"
#include "merge_common.cuh"
#include "merge_search.cuh"
#include "spmv_pipe.cuh"
#include "carry_exchange.cuh"

namespace mspmv {

static std::atomic<uint64_t> g_launches{0};

#define MSPMV_TRY(expr)                         \
    do {                                        \
        cudaError_t _e = (expr);                \
        if (_e != cudaSuccess) return (int)_e;  \
    } while (0)

struct DeviceInfo {
    int sm_count = 0;
    int max_smem_optin = 0;
    bool valid = false;
};

static int device_info(DeviceInfo& out)
{
    static std::mutex mu;
    static DeviceInfo cache[64];
    int dev = 0;
    MSPMV_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(mu);
    if (!cache[dev].valid) {
        MSPMV_TRY(cudaDeviceGetAttribute(&cache[dev].sm_count, cudaDevAttrMultiProcessorCount, dev));
        MSPMV_TRY(cudaDeviceGetAttribute(&cache[dev].max_smem_optin,
                                         cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        cache[dev].valid = true;
    }
    out = cache[dev];
    return 0;
}

// ---- pipe engine knobs (mspmv_set_option / environment; -1 = environment / default) ----------------
//   pipe_search        1: the producer warp finds the tile coordinates itself (one launch per CsrMV, default)
//                      0: tile_search_kernel first (two launches)
//   pipe_blocks_per_sm resident blocks per SM the grid is sized for (0 = as many as the shared-memory budget allows)
//   pipe_smem_kb       shared-memory budget per SM in KB (the rest of the 228 KB stays L1 for the gather misses)
static std::atomic<int> g_pipe_search{-1}, g_pipe_blocks_per_sm{-1}, g_pipe_smem_kb{-1};
static int env_int(const char* name, int dflt)
{
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}
static int pipe_search()
{
    int v = g_pipe_search.load(std::memory_order_relaxed);
    if (v >= 0) return v;
    static int env = env_int("MSPMV_PIPE_SEARCH", 1);
    return env;
}
static int pipe_blocks_per_sm_opt()
{
    int v = g_pipe_blocks_per_sm.load(std::memory_order_relaxed);
    if (v >= 0) return v;
    static int env = env_int("MSPMV_PIPE_BLOCKS_PER_SM", 0);
    return env;
}
// A (gather-bound): 160 KB of shared memory, 68 KB of L1 for the gather misses in flight; B (short rows,
// streaming-bound): 192 KB, one more resident block (profiles/sweep_r02_*.txt)
static int pipe_smem_kb(int cfg)
{
    int v = g_pipe_smem_kb.load(std::memory_order_relaxed);
    if (v > 0) return v;
    static int env = env_int("MSPMV_PIPE_SMEM_KB", 0);
    if (env > 0) return env;
    return cfg ? 192 : 160;
}

// Two instantiations of the pipe kernel per value type: A for matrices with long rows (the x gathers
// dominate), B for short rows (row bookkeeping dominates).  <IPT fp64, IPT fp32, value-ring slots,
// column-ring slots, gather-ahead, consumer warps, first-segment-in-register, values-before-gathers>; overridable at build time for tuning sweeps.
// (nvcc splits -D values at commas, hence one macro per field; PB_* default to PA_*)
#ifndef PA_I64
#define PA_I64 11
#endif
#ifndef PA_I32
#define PA_I32 13
#endif
#ifndef PA_VST
#define PA_VST 2
#endif
#ifndef PA_CST
#define PA_CST 1
#endif
#ifndef PA_AHEAD
#define PA_AHEAD 0
#endif
#ifndef PA_NW
#define PA_NW 4
#endif
#ifndef PA_FIR
#define PA_FIR 1
#endif
#ifndef PB_FIR
#define PB_FIR 0
#endif
#ifndef PA_VF
#define PA_VF 1
#endif
#ifndef PB_VF
#define PB_VF 1
#endif
#ifndef PB_I64
#define PB_I64 9
#endif
#ifndef PB_I32
#define PB_I32 PA_I32
#endif
#ifndef PB_VST
#define PB_VST PA_VST
#endif
#ifndef PB_CST
#define PB_CST 2
#endif
#ifndef PB_AHEAD
#define PB_AHEAD PA_AHEAD
#endif
#ifndef PB_NW
#define PB_NW PA_NW
#endif
#define MSPMV_PIPE_A PA_I64, PA_I32, PA_VST, PA_CST, PA_AHEAD, PA_NW, PA_FIR, PA_VF
#define MSPMV_PIPE_B PB_I64, PB_I32, PB_VST, PB_CST, PB_AHEAD, PB_NW, PB_FIR, PB_VF
#ifndef MSPMV_PIPE_B_MAX_ROW_ITEMS
#define MSPMV_PIPE_B_MAX_ROW_ITEMS 16  // B when (rows + nnz) / rows <= this
#endif
template <typename T, int I64, int I32, int VST, int CST, int AHEAD, int NW, int FIR, int VF>
using PipeCfgSel = PipeCfg<T, (sizeof(T) == 8 ? I64 : I32), VST, CST, AHEAD, NW, FIR, VF>;
template <typename T>
using PipeCfgA = PipeCfgSel<T, MSPMV_PIPE_A>;
template <typename T>
using PipeCfgB = PipeCfgSel<T, MSPMV_PIPE_B>;

static std::atomic<int> g_pipe_config{-1};  // 0 = by shape, 1 = A, 2 = B
static int pipe_config_for(int num_rows, int num_nonzeros)
{
    int v = g_pipe_config.load(std::memory_order_relaxed);
    if (v < 0) {
        static int env = env_int("MSPMV_PIPE_CONFIG", 0);
        v = env;
    }
    if (v == 1 || v == 2) return v - 1;
    const int64_t items = (int64_t)num_rows + num_nonzeros;
    return (num_rows > 0 && items <= (int64_t)MSPMV_PIPE_B_MAX_ROW_ITEMS * num_rows) ? 1 : 0;
}

template <class C>
static int pipe_blocks_per_sm(int cfg)
{
    const int per_block = (int)pipe_smem_bytes<C>() + 1024;  // + the per-block reservation of the driver
    int b = pipe_smem_kb(cfg) * 1024 / per_block;
    const int by_threads = 2048 / C::THREADS;
    if (b > by_threads) b = by_threads;
    const int want = pipe_blocks_per_sm_opt();
    if (want > 0 && want < b) b = want;
    return b < 1 ? 1 : b;
}

static inline size_t align256(size_t n) { return (n + 255) & ~size_t(255); }

template <typename T>
struct Plan {
    int64_t merge_items;
    int num_tiles;   // equal-length diagonal swaths of the merge path
    int num_blocks;  // threadblocks (each owns a contiguous run of tiles)
    int pipe_cfg;    // 0 = PipeCfgA, 1 = PipeCfgB
    int tile_items;  // merge items per tile of the chosen kernel
    size_t off_coords, off_carry_rows, off_carry_vals, off_ticket, bytes;
};

template <typename T>
static int make_plan(int num_rows, int num_nonzeros, Plan<T>& p)
{
    p.merge_items = (int64_t)num_rows + num_nonzeros;
    DeviceInfo di;
    int rc = device_info(di);
    if (rc) return rc;
    p.pipe_cfg = pipe_config_for(num_rows, num_nonzeros);
    p.tile_items = p.pipe_cfg ? PipeCfgB<T>::TILE : PipeCfgA<T>::TILE;
    p.num_tiles = (int)((p.merge_items + p.tile_items - 1) / p.tile_items);
    const int resident =
        di.sm_count * (p.pipe_cfg ? pipe_blocks_per_sm<PipeCfgB<T>>(1) : pipe_blocks_per_sm<PipeCfgA<T>>(0));
    p.num_blocks = p.num_tiles < resident ? p.num_tiles : resident;
    // 256-byte aligned regions, like AliasTemporaries (util_device.cuh:62-103).  The tile coordinates are
    // only written in the two-launch mode / by the debug export; the carries are one per block, but the
    // grid depends on run-time options, so both are sized by what cannot change: sm_count * 32 blocks at most.
    const size_t max_blocks = (size_t)(p.num_tiles < di.sm_count * 32 ? p.num_tiles : di.sm_count * 32);
    size_t off = 0;
    p.off_coords = off;
    off += align256(sizeof(int2) * (size_t)(p.num_tiles + 1));
    p.off_carry_rows = off;
    off += align256(sizeof(int) * max_blocks);
    p.off_carry_vals = off;
    off += align256(sizeof(T) * max_blocks);
    p.off_ticket = off;
    off += 256;
    p.bytes = off + 256;  // slack so an unaligned blob can be aligned up (util_device.cuh:68-80)
    return 0;
}

static int post_launch(const char* name, dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       int debug_sync)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (debug_sync)
        std::printf("Invoking %s<<<%u, %u, %zu, %p>>>()\n", name, grid.x, block.x, smem, (void*)stream);
    MSPMV_TRY(cudaPeekAtLastError());
    if (debug_sync) MSPMV_TRY(cudaStreamSynchronize(stream));
    return 0;
}

template <class C, bool AXPBY, bool SEARCH>
static int pipe_launch_impl(const Plan<typename C::value_type>& p, int2* coords, int2* coords_out, int* carry_rows,
                            typename C::value_type* carry_vals,
                            unsigned int* ticket, const typename C::value_type* values, const int* row_offsets,
                            const int* col, const typename C::value_type* x, typename C::value_type* y,
                            int num_rows, int num_nonzeros, typename C::value_type alpha,
                            typename C::value_type beta, cudaStream_t stream, int debug_sync)
{
    using T = typename C::value_type;
    constexpr size_t smem = pipe_smem_bytes<C>();
    static std::atomic<bool> configured[64];  // per device; a benign race only repeats the calls
    int dev = 0;
    MSPMV_TRY(cudaGetDevice(&dev));
    if (!configured[dev & 63].load(std::memory_order_relaxed)) {
        MSPMV_TRY(cudaFuncSetAttribute(spmv_pipe_kernel<C, AXPBY, SEARCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
        int pct = (pipe_smem_kb(p.pipe_cfg) * 100 + 227) / 228;
        if (pct > 100) pct = 100;
        MSPMV_TRY(cudaFuncSetAttribute(spmv_pipe_kernel<C, AXPBY, SEARCH>,
                                       cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        configured[dev & 63].store(true, std::memory_order_relaxed);
    }
    const int shift_v = (int)((reinterpret_cast<uintptr_t>(values) & 15) / sizeof(T));
    const int shift_c = (int)((reinterpret_cast<uintptr_t>(col) & 15) / sizeof(int));
    const int shift_r = (int)((reinterpret_cast<uintptr_t>(row_offsets) & 15) / sizeof(int));
    if (!SEARCH) {
        dim3 sgrid((p.num_tiles + 1 + 127) / 128), sblock(128);
        tile_search_kernel<<<sgrid, sblock, 0, stream>>>(row_offsets + 1, num_rows, num_nonzeros, C::TILE,
                                                         p.num_tiles, coords, ticket);
        int rc = post_launch("tile_search_kernel", sgrid, sblock, 0, stream, debug_sync);
        if (rc) return rc;
    } else if (p.num_blocks > 1) {
        MSPMV_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), stream));  // the temp blob arrives uninitialised
    }
    dim3 grid(p.num_blocks), block(C::THREADS);
    spmv_pipe_kernel<C, AXPBY, SEARCH><<<grid, block, smem, stream>>>(
        values, row_offsets, col, x, y, coords, coords_out, carry_rows, carry_vals, ticket, alpha, beta, num_rows,
        num_nonzeros, p.num_tiles, shift_v, shift_c, shift_r);
    return post_launch("spmv_pipe_kernel", grid, block, smem, stream, debug_sync);
}

static std::atomic<int> g_pipe_export_coords{0};  // debug: the producer warps write the coordinates they found

template <typename T, bool AXPBY>
static int csrmv_launch(const Plan<T>& p, char* temp, const T* values, const int* row_offsets,
                        const int* col, const T* x, T* y, int num_rows, int num_nonzeros, T alpha,
                        T beta, cudaStream_t stream, int debug_sync)
{
    int2* coords = reinterpret_cast<int2*>(temp + p.off_coords);
    int* carry_rows = reinterpret_cast<int*>(temp + p.off_carry_rows);
    T* carry_vals = reinterpret_cast<T*>(temp + p.off_carry_vals);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(temp + p.off_ticket);
    auto go = [&](auto cfg, auto search) {
        using C = decltype(cfg);
        constexpr bool SEARCH = decltype(search)::value;
        int2* coords_out = (SEARCH && g_pipe_export_coords.load(std::memory_order_relaxed)) ? coords : nullptr;
        return pipe_launch_impl<C, AXPBY, SEARCH>(p, coords, coords_out, carry_rows, carry_vals, ticket, values,
                                                  row_offsets, col, x, y, num_rows, num_nonzeros, alpha, beta, stream,
                                                  debug_sync);
    };
    if (p.pipe_cfg) return pipe_search() ? go(PipeCfgB<T>(), std::true_type()) : go(PipeCfgB<T>(), std::false_type());
    return pipe_search() ? go(PipeCfgA<T>(), std::true_type()) : go(PipeCfgA<T>(), std::false_type());
}

template <typename T, bool AXPBY>
static int csrmv(void* d_temp, size_t* temp_bytes, const T* values, const int* row_offsets,
                 const int* col, const T* x, T* y, int num_rows, int num_cols, int num_nonzeros,
                 T alpha, T beta, void* stream_v, int debug_sync)
{
    (void)num_cols;
    if (!temp_bytes || num_rows < 0 || num_nonzeros < 0) return (int)cudaErrorInvalidValue;
    if ((int64_t)num_rows + num_nonzeros > (int64_t)INT_MAX - 65536) return (int)cudaErrorInvalidValue;
    Plan<T> p;
    int rc = make_plan<T>(num_rows, num_nonzeros, p);
    if (rc) return rc;
    if (d_temp == nullptr) {  // size query (dispatch_spmv_orig.cuh:651-655)
        *temp_bytes = p.bytes;
        return 0;
    }
    if (*temp_bytes < p.bytes) return (int)cudaErrorInvalidValue;  // util_device.cuh:90-93
    if (num_rows == 0) return 0;
    if (!row_offsets || !y || (num_nonzeros > 0 && (!values || !col || !x))) return (int)cudaErrorInvalidValue;
    char* temp = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(d_temp) + 255) & ~uintptr_t(255));
    return csrmv_launch<T, AXPBY>(p, temp, values, row_offsets, col, x, y, num_rows, num_nonzeros,
                                  alpha, beta, (cudaStream_t)stream_v, debug_sync);
}

// ---- multi-GPU carry application -------------------------------------------------------------
// One thread; p <= a few dozen records.  Shard order, guard row < num_rows_global
// (cpu_spmv.cpp:348-352), only rows this rank stores.
template <typename T>
__global__ void apply_carries_kernel(T* __restrict__ y_local, int y_row_begin, int y_rows,
                                     int num_rows_global, const int* __restrict__ carry_rows,
                                     const T* __restrict__ carry_vals, int num_shards)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int g = 0; g < num_shards - 1; ++g) {
        int row = carry_rows[g];
        if (row < num_rows_global && row >= y_row_begin && row < y_row_begin + y_rows)
            y_local[row - y_row_begin] += carry_vals[g];
    }
}

template <typename T>
static int apply_carries(T* y_local, int y_row_begin, int y_rows, int num_rows_global,
                         const int* carry_rows, const T* carry_vals, int num_shards, void* stream_v)
{
    cudaStream_t stream = (cudaStream_t)stream_v;
    apply_carries_kernel<T><<<1, 32, 0, stream>>>(y_local, y_row_begin, y_rows, num_rows_global,
                                                  carry_rows, carry_vals, num_shards);
    return post_launch("apply_carries_kernel", dim3(1), dim3(32), 0, stream, 0);
}

template <typename T>
static int exchange_carries(T* y_local, int local_rows, int y_row_begin, int y_rows, int num_rows_global,
                            const int* carry_rows, void* const* peer_bufs, int rank, int num_shards,
                            unsigned long long* epoch, void* stream_v)
{
    if (!y_local || !carry_rows || !peer_bufs || !epoch || local_rows < 1 || num_shards < 1 || rank < 0 ||
        rank >= num_shards)
        return (int)cudaErrorInvalidValue;
    cudaStream_t stream = (cudaStream_t)stream_v;
    carry_exchange_kernel<T><<<1, 32, 0, stream>>>(y_local, local_rows - 1, y_row_begin, y_rows, num_rows_global,
                                                   carry_rows, reinterpret_cast<uint64_t* const*>(peer_bufs), rank,
                                                   num_shards, epoch, kExchangePhasePush | kExchangePhaseFold);
    return post_launch("carry_exchange_kernel", dim3(1), dim3(32), 0, stream, 0);
}

static void host_search(const int* row_offsets, int num_rows, int num_nonzeros, int64_t diagonal,
                        int* out_x, int* out_y)
{
    const int* row_end = row_offsets + 1;
    int64_t total = (int64_t)num_rows + num_nonzeros;
    int diag = (int)(diagonal < total ? diagonal : total);
    int lo = diag - num_nonzeros > 0 ? diag - num_nonzeros : 0;
    int hi = diag < num_rows ? diag : num_rows;
    while (lo < hi) {
        int mid = (int)(((int64_t)lo + hi) >> 1);
        if (row_end[mid] <= diag - mid - 1)
            lo = mid + 1;
        else
            hi = mid;
    }
    *out_x = lo < num_rows ? lo : num_rows;
    *out_y = diag - lo;
}

}  // namespace mspmv

// ================================================================================================
// Session: host-buffer operator (upload once, apply many) -- gpu_spmv.cu:542-556 + :421-432.
// ================================================================================================
struct mspmv_session {
    int device = 0, value_bytes = 0, rows = 0, cols = 0, nnz = 0;
    int* d_row_offsets = nullptr;
    int* d_col = nullptr;
    void* d_val = nullptr;
    void* d_temp = nullptr;
    size_t temp_bytes = 0;
    static constexpr int kSlots = 3;
    void* d_x[kSlots] = {nullptr, nullptr, nullptr};
    void* d_y[kSlots] = {nullptr, nullptr, nullptr};
    cudaStream_t s_h2d = nullptr, s_compute = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_x[kSlots], ev_k[kSlots], ev_y[kSlots];
    bool events = false;
};

static int session_csrmv(mspmv_session* s, int slot, cudaStream_t stream)
{
    if (s->value_bytes == 8)
        return mspmv_csrmv_f64(s->d_temp, &s->temp_bytes, (const double*)s->d_val, s->d_row_offsets,
                               s->d_col, (const double*)s->d_x[slot], (double*)s->d_y[slot], s->rows,
                               s->cols, s->nnz, stream, 0);
    return mspmv_csrmv_f32(s->d_temp, &s->temp_bytes, (const float*)s->d_val, s->d_row_offsets,
                           s->d_col, (const float*)s->d_x[slot], (float*)s->d_y[slot], s->rows, s->cols,
                           s->nnz, stream, 0);
}

extern "C" {

using namespace mspmv;

int mspmv_csrmv_f32(void* t, size_t* tb, const float* v, const int* ro, const int* ci, const float* x,
                    float* y, int r, int c, int nnz, mspmv_stream_t s, int dbg)
{
    return csrmv<float, false>(t, tb, v, ro, ci, x, y, r, c, nnz, 1.f, 0.f, s, dbg);
}
int mspmv_csrmv_f64(void* t, size_t* tb, const double* v, const int* ro, const int* ci,
                    const double* x, double* y, int r, int c, int nnz, mspmv_stream_t s, int dbg)
{
    return csrmv<double, false>(t, tb, v, ro, ci, x, y, r, c, nnz, 1.0, 0.0, s, dbg);
}
int mspmv_csrmv_axpby_f32(void* t, size_t* tb, const float* v, const int* ro, const int* ci,
                          const float* x, float* y, int r, int c, int nnz, float alpha, float beta,
                          mspmv_stream_t s, int dbg)
{
    return csrmv<float, true>(t, tb, v, ro, ci, x, y, r, c, nnz, alpha, beta, s, dbg);
}
int mspmv_csrmv_axpby_f64(void* t, size_t* tb, const double* v, const int* ro, const int* ci,
                          const double* x, double* y, int r, int c, int nnz, double alpha, double beta,
                          mspmv_stream_t s, int dbg)
{
    return csrmv<double, true>(t, tb, v, ro, ci, x, y, r, c, nnz, alpha, beta, s, dbg);
}

int mspmv_merge_path_search(const int* d_row_offsets, int num_rows, int num_nonzeros,
                            const int* d_diagonals, int n, int* d_coords, mspmv_stream_t stream)
{
    if (n <= 0) return 0;
    dim3 grid((n + 127) / 128), block(128);
    diagonal_search_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
        d_row_offsets + 1, num_rows, num_nonzeros, d_diagonals, n, reinterpret_cast<int2*>(d_coords));
    return post_launch("diagonal_search_kernel", grid, block, 0, (cudaStream_t)stream, 0);
}

int mspmv_csrmv_swath_coords(const int* d_row_offsets, int num_rows, int num_nonzeros, int value_bytes,
                             int* num_swaths, int* d_coords, mspmv_stream_t stream)
{
    if (!num_swaths || (value_bytes != 4 && value_bytes != 8)) return (int)cudaErrorInvalidValue;
    int64_t per;
    int n;
    if (value_bytes == 8) {
        Plan<double> p;
        int rc = make_plan<double>(num_rows, num_nonzeros, p);
        if (rc) return rc;
        n = p.num_tiles;
        per = p.tile_items;
    } else {
        Plan<float> p;
        int rc = make_plan<float>(num_rows, num_nonzeros, p);
        if (rc) return rc;
        n = p.num_tiles;
        per = p.tile_items;
    }
    *num_swaths = n;
    if (!d_coords) return 0;
    dim3 grid((n + 1 + 127) / 128), block(128);
    tile_search_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
        d_row_offsets + 1, num_rows, num_nonzeros, (int)per, n, reinterpret_cast<int2*>(d_coords), nullptr);
    return post_launch("tile_search_kernel", grid, block, 0, (cudaStream_t)stream, 0);
}

void mspmv_host_merge_path_search(const int* row_offsets, int num_rows, int num_nonzeros,
                                  int64_t diagonal, int* out_x, int* out_y)
{
    host_search(row_offsets, num_rows, num_nonzeros, diagonal, out_x, out_y);
}

void mspmv_shard_partition(const int* row_offsets, int num_rows, int num_nonzeros, int num_shards,
                           int* coords)
{
    // cpu_spmv.cpp:311-321 with p = num_shards
    if (num_shards < 1 || !coords || !row_offsets) return;
    int64_t total = (int64_t)num_rows + num_nonzeros;
    int64_t share = (total + num_shards - 1) / num_shards;
    for (int g = 0; g <= num_shards; ++g) {
        int64_t d = share * g;
        if (d > total) d = total;
        host_search(row_offsets, num_rows, num_nonzeros, d, &coords[2 * g], &coords[2 * g + 1]);
    }
}

void mspmv_shard_row_offsets(const int* row_offsets, int x0, int y0, int x1, int y1,
                             int* local_row_offsets)
{
    int local_rows = x1 - x0 + 1;
    local_row_offsets[0] = 0;
    for (int i = 0; i < local_rows - 1; ++i) local_row_offsets[i + 1] = row_offsets[x0 + i + 1] - y0;
    local_row_offsets[local_rows] = y1 - y0;
}

int mspmv_apply_carries_f32(float* y, int b, int n, int rows_global, const int* cr, const float* cv,
                            int shards, mspmv_stream_t s)
{
    return apply_carries<float>(y, b, n, rows_global, cr, cv, shards, s);
}
int mspmv_apply_carries_f64(double* y, int b, int n, int rows_global, const int* cr, const double* cv,
                            int shards, mspmv_stream_t s)
{
    return apply_carries<double>(y, b, n, rows_global, cr, cv, shards, s);
}

size_t mspmv_exchange_buffer_bytes(int num_shards) { return sizeof(uint64_t) * 4u * (size_t)(num_shards > 0 ? num_shards : 0); }
int mspmv_exchange_carries_f32(float* y, int local_rows, int b, int n, int rows_global, const int* cr,
                               void* const* peer_bufs, int rank, int shards, unsigned long long* epoch,
                               mspmv_stream_t s)
{
    return exchange_carries<float>(y, local_rows, b, n, rows_global, cr, peer_bufs, rank, shards, epoch, s);
}
int mspmv_exchange_carries_f64(double* y, int local_rows, int b, int n, int rows_global, const int* cr,
                               void* const* peer_bufs, int rank, int shards, unsigned long long* epoch,
                               mspmv_stream_t s)
{
    return exchange_carries<double>(y, local_rows, b, n, rows_global, cr, peer_bufs, rank, shards, epoch, s);
}

// ---- session -----------------------------------------------------------------------------------
int mspmv_session_create(mspmv_session** out, int device, int value_bytes, int num_rows, int num_cols,
                         int num_nonzeros, const int* row_offsets, const int* column_indices,
                         const void* values)
{
    if (!out || (value_bytes != 4 && value_bytes != 8)) return (int)cudaErrorInvalidValue;
    MSPMV_TRY(cudaSetDevice(device));
    mspmv_session* s = new mspmv_session();
    s->device = device;
    s->value_bytes = value_bytes;
    s->rows = num_rows;
    s->cols = num_cols;
    s->nnz = num_nonzeros;
    int rc = 0;
    auto fail = [&](cudaError_t e) {
        rc = (int)e;
        return e != cudaSuccess;
    };
    size_t vb = (size_t)value_bytes;
    if (fail(cudaMalloc(&s->d_row_offsets, sizeof(int) * (size_t)(num_rows + 1))) ||
        fail(cudaMalloc(&s->d_col, sizeof(int) * (size_t)(num_nonzeros > 0 ? num_nonzeros : 1))) ||
        fail(cudaMalloc(&s->d_val, vb * (size_t)(num_nonzeros > 0 ? num_nonzeros : 1)))) {
        mspmv_session_destroy(s);
        return rc;
    }
    for (int i = 0; i < mspmv_session::kSlots; ++i) {
        if (fail(cudaMalloc(&s->d_x[i], vb * (size_t)(num_cols > 0 ? num_cols : 1))) ||
            fail(cudaMalloc(&s->d_y[i], vb * (size_t)(num_rows > 0 ? num_rows : 1)))) {
            mspmv_session_destroy(s);
            return rc;
        }
    }
    if (fail(cudaStreamCreateWithFlags(&s->s_h2d, cudaStreamNonBlocking)) ||
        fail(cudaStreamCreateWithFlags(&s->s_compute, cudaStreamNonBlocking)) ||
        fail(cudaStreamCreateWithFlags(&s->s_d2h, cudaStreamNonBlocking))) {
        mspmv_session_destroy(s);
        return rc;
    }
    for (int i = 0; i < mspmv_session::kSlots; ++i) {
        cudaEventCreateWithFlags(&s->ev_x[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&s->ev_k[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&s->ev_y[i], cudaEventDisableTiming);
    }
    s->events = true;
    if (fail(cudaMemcpyAsync(s->d_row_offsets, row_offsets, sizeof(int) * (size_t)(num_rows + 1),
                             cudaMemcpyHostToDevice, s->s_compute)) ||
        fail(cudaMemcpyAsync(s->d_col, column_indices, sizeof(int) * (size_t)num_nonzeros,
                             cudaMemcpyHostToDevice, s->s_compute)) ||
        fail(cudaMemcpyAsync(s->d_val, values, vb * (size_t)num_nonzeros, cudaMemcpyHostToDevice,
                             s->s_compute))) {
        mspmv_session_destroy(s);
        return rc;
    }
    // temp-storage size query, then allocate (gpu_spmv.cu:390-398)
    size_t bytes = 0;
    rc = value_bytes == 8
             ? mspmv_csrmv_f64(nullptr, &bytes, nullptr, nullptr, nullptr, nullptr, nullptr, num_rows,
                               num_cols, num_nonzeros, nullptr, 0)
             : mspmv_csrmv_f32(nullptr, &bytes, nullptr, nullptr, nullptr, nullptr, nullptr, num_rows,
                               num_cols, num_nonzeros, nullptr, 0);
    if (rc || fail(cudaMalloc(&s->d_temp, bytes)) || fail(cudaStreamSynchronize(s->s_compute))) {
        mspmv_session_destroy(s);
        return rc;
    }
    s->temp_bytes = bytes;
    *out = s;
    return 0;
}

int mspmv_session_apply(mspmv_session* s, const void* x_host, void* y_host)
{
    if (!s) return (int)cudaErrorInvalidValue;
    MSPMV_TRY(cudaSetDevice(s->device));
    size_t vb = (size_t)s->value_bytes;
    MSPMV_TRY(cudaMemcpyAsync(s->d_x[0], x_host, vb * (size_t)s->cols, cudaMemcpyHostToDevice,
                              s->s_compute));
    int rc = session_csrmv(s, 0, s->s_compute);
    if (rc) return rc;
    MSPMV_TRY(cudaMemcpyAsync(y_host, s->d_y[0], vb * (size_t)s->rows, cudaMemcpyDeviceToHost,
                              s->s_compute));
    MSPMV_TRY(cudaStreamSynchronize(s->s_compute));
    return 0;
}

int mspmv_session_apply_many(mspmv_session* s, int n, const void* xs_host, void* ys_host)
{
    if (!s || n < 0) return (int)cudaErrorInvalidValue;
    MSPMV_TRY(cudaSetDevice(s->device));
    constexpr int K = mspmv_session::kSlots;
    size_t vb = (size_t)s->value_bytes;
    size_t xbytes = vb * (size_t)s->cols, ybytes = vb * (size_t)s->rows;
    const char* xs = (const char*)xs_host;
    char* ys = (char*)ys_host;
    // three-stage pipeline over K device slots: H2D(i) | kernel(i-1) | D2H(i-2) run concurrently
    for (int i = 0; i < n; ++i) {
        int slot = i % K;
        if (i >= K) MSPMV_TRY(cudaStreamWaitEvent(s->s_h2d, s->ev_k[slot], 0));  // x slot consumed
        MSPMV_TRY(cudaMemcpyAsync(s->d_x[slot], xs + xbytes * (size_t)i, xbytes, cudaMemcpyHostToDevice,
                                  s->s_h2d));
        MSPMV_TRY(cudaEventRecord(s->ev_x[slot], s->s_h2d));
        MSPMV_TRY(cudaStreamWaitEvent(s->s_compute, s->ev_x[slot], 0));
        if (i >= K) MSPMV_TRY(cudaStreamWaitEvent(s->s_compute, s->ev_y[slot], 0));  // y slot drained
        int rc = session_csrmv(s, slot, s->s_compute);
        if (rc) return rc;
        MSPMV_TRY(cudaEventRecord(s->ev_k[slot], s->s_compute));
        MSPMV_TRY(cudaStreamWaitEvent(s->s_d2h, s->ev_k[slot], 0));
        MSPMV_TRY(cudaMemcpyAsync(ys + ybytes * (size_t)i, s->d_y[slot], ybytes, cudaMemcpyDeviceToHost,
                                  s->s_d2h));
        MSPMV_TRY(cudaEventRecord(s->ev_y[slot], s->s_d2h));
    }
    MSPMV_TRY(cudaStreamSynchronize(s->s_d2h));
    MSPMV_TRY(cudaStreamSynchronize(s->s_compute));
    MSPMV_TRY(cudaStreamSynchronize(s->s_h2d));
    return 0;
}

void mspmv_session_destroy(mspmv_session* s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->s_compute) cudaStreamSynchronize(s->s_compute);
    cudaFree(s->d_row_offsets);
    cudaFree(s->d_col);
    cudaFree(s->d_val);
    cudaFree(s->d_temp);
    for (int i = 0; i < mspmv_session::kSlots; ++i) {
        cudaFree(s->d_x[i]);
        cudaFree(s->d_y[i]);
        if (s->events) {
            cudaEventDestroy(s->ev_x[i]);
            cudaEventDestroy(s->ev_k[i]);
            cudaEventDestroy(s->ev_y[i]);
        }
    }
    if (s->s_h2d) cudaStreamDestroy(s->s_h2d);
    if (s->s_compute) cudaStreamDestroy(s->s_compute);
    if (s->s_d2h) cudaStreamDestroy(s->s_d2h);
    delete s;
}

int mspmv_host_alloc(void** out, size_t bytes) { return (int)cudaMallocHost(out, bytes); }
int mspmv_host_free(void* p) { return (int)cudaFreeHost(p); }

int mspmv_version(void) { return MSPMV_VERSION_MAJOR * 100 + MSPMV_VERSION_MINOR; }

int mspmv_ptx_version(int* ptx_version)
{
    if (!ptx_version) return (int)cudaErrorInvalidValue;
    cudaFuncAttributes attr;
    MSPMV_TRY(cudaFuncGetAttributes(&attr, spmv_pipe_kernel<PipeCfgA<double>, false, true>));
    *ptx_version = attr.ptxVersion * 10;
    return 0;
}
uint64_t mspmv_launch_count(void) { return g_launches.load(); }

int mspmv_csrmv_config(int value_bytes, int num_rows, int num_nonzeros, int* out)
{
    if (!out || (value_bytes != 4 && value_bytes != 8)) return (int)cudaErrorInvalidValue;
    auto fill = [&](auto tag) {
        using T = decltype(tag);
        Plan<T> p;
        int rc = make_plan<T>(num_rows, num_nonzeros, p);
        if (rc) return rc;
        out[0] = p.num_blocks;
        out[1] = p.pipe_cfg ? PipeCfgB<T>::THREADS : PipeCfgA<T>::THREADS;
        out[2] = p.tile_items;
        out[3] = (int)(p.pipe_cfg ? pipe_smem_bytes<PipeCfgB<T>>() : pipe_smem_bytes<PipeCfgA<T>>());
        out[4] = pipe_search() ? 1 : 2;
        out[5] = p.num_tiles;
        return 0;
    };
    return value_bytes == 8 ? fill(double()) : fill(float());
}

#if MSPMV_PIPE_PROFILE
// tuning builds only (not declared in include/mergespmv.h): per-phase SM cycles of thread 0 of every block
int mspmv_debug_profile(unsigned long long* out, int reset)
{
    if (out) MSPMV_TRY(cudaMemcpyFromSymbol(out, g_pipe_prof, sizeof(unsigned long long) * 8));
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        MSPMV_TRY(cudaMemcpyToSymbol(g_pipe_prof, z, sizeof(z)));
    }
    return 0;
}
#endif

const char* mspmv_error_string(int err) { return cudaGetErrorString((cudaError_t)err); }

int mspmv_set_engine(const char* name)
{
    // one engine since round 2 ("pipe", spmv_pipe.cuh); the entry point stays for callers that name it
    if (!name) return 1;
    return (!std::strcmp(name, "auto") || !std::strcmp(name, "pipe")) ? 0 : 1;
}

int mspmv_set_option(const char* name, int value)
{
    if (!name) return 1;
    if (!std::strcmp(name, "pipe_search")) {
        g_pipe_search = value < 0 ? -1 : (value != 0);
        return 0;
    }
    if (!std::strcmp(name, "pipe_blocks_per_sm")) {
        g_pipe_blocks_per_sm = value < 0 ? -1 : value;
        return 0;
    }
    if (!std::strcmp(name, "pipe_config")) {
        if (value > 2) return 1;
        g_pipe_config = value < 0 ? -1 : value;
        return 0;
    }
    if (!std::strcmp(name, "pipe_smem_kb")) {
        g_pipe_smem_kb = value <= 0 ? -1 : value;
        return 0;
    }
    if (!std::strcmp(name, "pipe_export_coords")) {
        g_pipe_export_coords = value > 0;
        return 0;
    }
    return 1;
}

}  // extern "C"
