// tile_emu.cpp -- runs the tile engine's three kernels (merge-spmv_b200/csrc/spmv_tile.cuh, compiled
// unchanged) under the host SIMT interpreter.  TEST INFRASTRUCTURE ONLY: built by
// tests/test_kernel_emu.py, never part of the product.  The launch sequence mirrors
// csrmv_launch() in merge-spmv_b200/csrc/mergespmv.cu (search -> tile -> carry fix-up).
#define MSPMV_PTX_HEADER "ptx_emu.cuh"
#include "simt_emu.hpp"

#include "spmv_tile.cuh"
#include "spmv_tile3.cuh"
#include "carry_exchange.cuh"
#include "spmv_stream.cuh"

#include <vector>

namespace {

using namespace mspmv;

template <typename T>
int shift_of(const void* p)
{
    return (int)((reinterpret_cast<uintptr_t>(p) & 15) / sizeof(T));
}

template <typename T, bool AXPBY>
int run(const T* values, const int* row_offsets, const int* col, const T* x, T* y, int num_rows,
        int num_nonzeros, T alpha, T beta, int prefetch_ahead, int* stats, int mode = 0)
{
    using C = TileCfg<T>;
    if (num_rows <= 0) return 0;
    const int64_t merge_items = (int64_t)num_rows + num_nonzeros;
    const int num_tiles = (int)((merge_items + C::TILE - 1) / C::TILE);
    const int num_fix_blocks = (num_tiles + C::FIX - 1) / C::FIX;
    // temporaries of exactly the size the dispatch carves out of the temp blob (so that AddressSanitizer
    // sees any access past them), filled with garbage on purpose (the product's temp blob is uninitialised)
    std::vector<int2> coords_buf((size_t)num_tiles + 1);
    std::vector<int> cr_buf((size_t)num_tiles), c2r_buf((size_t)num_fix_blocks);
    std::vector<T> cv_buf((size_t)num_tiles), c2v_buf((size_t)num_fix_blocks);
    int2* coords = coords_buf.data();
    int* carry_rows = cr_buf.data();
    int* carry2_rows = c2r_buf.data();
    T* carry_vals = cv_buf.data();
    T* carry2_vals = c2v_buf.data();
    std::memset(coords_buf.data(), 0xEE, coords_buf.size() * sizeof(int2));
    std::memset(cr_buf.data(), 0xEE, cr_buf.size() * sizeof(int));
    std::memset(cv_buf.data(), 0xEE, cv_buf.size() * sizeof(T));
    std::memset(c2r_buf.data(), 0xEE, c2r_buf.size() * sizeof(int));
    std::memset(c2v_buf.data(), 0xEE, c2v_buf.size() * sizeof(T));
    unsigned int ticket = 0xEEEEEEEEu;

    const int* row_end = row_offsets + 1;
    if (stats) {
        stats[0] = num_tiles;
        stats[1] = C::TILE;
        stats[2] = C::THREADS;
    }
    if (mode == 1) {  // single-launch path for small matrices (cudaMemsetAsync of the ticket + one kernel)
        ticket = 0u;
        emu::launch((unsigned)num_tiles, (unsigned)C::THREADS, [&] {
            spmv_tile_fused_kernel<T, AXPBY>(values, row_offsets, col, x, y, carry_rows, carry_vals, alpha, beta,
                                             num_rows, num_nonzeros, shift_of<T>(values), shift_of<int>(col),
                                             shift_of<int>(row_offsets), &ticket);
        });
        return 0;
    }
    emu::launch((unsigned)((num_tiles + 1 + 127) / 128), 128, [&] {
        tile_search_kernel(row_end, num_rows, num_nonzeros, C::TILE, num_tiles, coords, &ticket);
    });
    emu::launch((unsigned)num_tiles, (unsigned)C::THREADS, [&] {
        if (mode == 3)
            spmv_tile3_kernel<T, AXPBY>(values, row_offsets, col, x, y, coords, carry_rows, carry_vals, alpha, beta,
                                        num_rows, num_nonzeros, shift_of<T>(values), shift_of<int>(col),
                                        shift_of<int>(row_offsets), prefetch_ahead);
        else
            spmv_tile_kernel<T, AXPBY>(values, row_offsets, col, x, y, coords, carry_rows, carry_vals, alpha, beta,
                                       num_rows, num_nonzeros, shift_of<T>(values), shift_of<int>(col),
                                       shift_of<int>(row_offsets), prefetch_ahead);
    });
    if (num_tiles > 1) {
        emu::launch((unsigned)num_fix_blocks, (unsigned)C::FIX, [&] {
            carry_fixup_block_kernel<T, AXPBY>(carry_rows, carry_vals, num_tiles, num_rows, y, alpha, carry2_rows,
                                               carry2_vals, &ticket);
        });
    }
    return 0;
}

// the alternative "stream" engine (persistent swaths, three warp-specialised stages over mbarrier rings):
// stream_launch() + the runs fix-up of csrmv_launch(), with `sm_count` standing in for the device's SM count
template <typename T, bool AXPBY>
int run_stream(const T* values, const int* row_offsets, const int* col, const T* x, T* y, int num_rows,
               int num_nonzeros, T alpha, T beta, int sm_count, int* stats)
{
    if (num_rows <= 0) return 0;
    const StreamGeom g = stream_geometry<T>((int64_t)num_rows + num_nonzeros, sm_count);
    const int n = g.num_swaths;
    std::vector<int2> coords((size_t)n + 1);
    std::vector<int> carry_rows((size_t)n);
    std::vector<T> carry_vals((size_t)n);
    std::memset(coords.data(), 0xEE, coords.size() * sizeof(int2));
    std::memset(carry_rows.data(), 0xEE, carry_rows.size() * sizeof(int));
    std::memset(carry_vals.data(), 0xEE, carry_vals.size() * sizeof(T));
    const int sv = shift_of<T>(values), sc = shift_of<int>(col), sr = shift_of<int>(row_offsets);
    const bool vec = sv == 0 && sc == 0;
    emu::launch((unsigned)n, (unsigned)g.threads, [&] {
        if (vec)
            spmv_stream_kernel<T, AXPBY, true>(values, row_offsets, col, x, y, num_rows, num_nonzeros, g.swath_items,
                                               coords.data(), carry_rows.data(), carry_vals.data(), alpha, beta, sv, sc, sr);
        else
            spmv_stream_kernel<T, AXPBY, false>(values, row_offsets, col, x, y, num_rows, num_nonzeros, g.swath_items,
                                                coords.data(), carry_rows.data(), carry_vals.data(), alpha, beta, sv, sc, sr);
    });
    if (n > 1)
        emu::launch((unsigned)((n + 255) / 256), 256, [&] {
            carry_fixup_runs_kernel<T, AXPBY>(carry_rows.data(), carry_vals.data(), n, num_rows, y, alpha);
        });
    if (stats) {
        stats[0] = n;
        stats[1] = g.tile_items;
        stats[2] = g.threads;
        stats[3] = g.swath_items;
    }
    return 0;
}

}  // namespace

extern "C" {

int emu_csrmv_stream_f64(const double* v, const int* ro, const int* ci, const double* x, double* y, int rows, int nnz,
                         double alpha, double beta, int axpby, int sm_count, int* stats)
{
    return axpby ? run_stream<double, true>(v, ro, ci, x, y, rows, nnz, alpha, beta, sm_count, stats)
                 : run_stream<double, false>(v, ro, ci, x, y, rows, nnz, alpha, beta, sm_count, stats);
}
int emu_csrmv_stream_f32(const float* v, const int* ro, const int* ci, const float* x, float* y, int rows, int nnz,
                         float alpha, float beta, int axpby, int sm_count, int* stats)
{
    return axpby ? run_stream<float, true>(v, ro, ci, x, y, rows, nnz, alpha, beta, sm_count, stats)
                 : run_stream<float, false>(v, ro, ci, x, y, rows, nnz, alpha, beta, sm_count, stats);
}

// mode: 0 = search + tile + fix-up (shipped), 1 = single fused launch, 3 = search + tile variant 3 + fix-up
int emu_csrmv_f64(const double* v, const int* ro, const int* ci, const double* x, double* y, int rows, int nnz,
                  double alpha, double beta, int axpby, int prefetch_ahead, int* stats, int mode)
{
    return axpby ? run<double, true>(v, ro, ci, x, y, rows, nnz, alpha, beta, prefetch_ahead, stats, mode)
                 : run<double, false>(v, ro, ci, x, y, rows, nnz, alpha, beta, prefetch_ahead, stats, mode);
}
int emu_csrmv_f32(const float* v, const int* ro, const int* ci, const float* x, float* y, int rows, int nnz,
                  float alpha, float beta, int axpby, int prefetch_ahead, int* stats, int mode)
{
    return axpby ? run<float, true>(v, ro, ci, x, y, rows, nnz, alpha, beta, prefetch_ahead, stats, mode)
                 : run<float, false>(v, ro, ci, x, y, rows, nnz, alpha, beta, prefetch_ahead, stats, mode);
}

// The NVLink carry exchange (carry_exchange.cuh) with `world` ranks simulated in one process: every
// rank has its own y slice, exchange buffer and epoch counter; "peer memory" is just the other
// rank's buffer.  Per step all ranks run the push phase, then all ranks run the fold phase (the
// product launches both phases in one kernel; the interpreter runs one block at a time, so a
// rank polling for a flag another rank has not written yet would never be satisfied).
//   cuts[world+1]   first global row of every rank (rank g owns rows [cuts[g], cuts[g+1]))
//   carry_rows[world], carries[steps][world]; y_out[steps][rows_global] = the owned rows after the fold
int emu_exchange_f64(int world, int steps, const int* cuts, const int* carry_rows, const double* carries,
                     double* y_out, int use_f32)
{
    const int rows_global = cuts[world];
    std::vector<std::vector<uint64_t>> bufs(world, std::vector<uint64_t>(4 * (size_t)world, 0));
    std::vector<uint64_t*> ptrs(world);
    for (int g = 0; g < world; ++g) ptrs[g] = bufs[g].data();
    std::vector<unsigned long long> epoch(world, 0);
    for (int st = 0; st < steps; ++st) {
        std::vector<std::vector<double>> yd(world);
        std::vector<std::vector<float>> yf(world);
        for (int g = 0; g < world; ++g) {
            const int owned = cuts[g + 1] - cuts[g];
            yd[g].assign(owned + 1, 0.0);
            yf[g].assign(owned + 1, 0.f);
            for (int i = 0; i < owned; ++i) yd[g][i] = yf[g][i] = (float)(100 * st + cuts[g] + i);  // "A*x" of the owned rows
            yd[g][owned] = carries[(size_t)st * world + g];
            yf[g][owned] = (float)carries[(size_t)st * world + g];
        }
        for (int phase = 1; phase <= 2; ++phase)
            for (int g = 0; g < world; ++g) {
                const int owned = cuts[g + 1] - cuts[g];
                emu::launch(1, 32, [&] {
                    if (use_f32)
                        carry_exchange_kernel<float>(yf[g].data(), owned, cuts[g], owned, rows_global, carry_rows,
                                                     ptrs.data(), g, world, &epoch[g], phase);
                    else
                        carry_exchange_kernel<double>(yd[g].data(), owned, cuts[g], owned, rows_global, carry_rows,
                                                      ptrs.data(), g, world, &epoch[g], phase);
                });
            }
        for (int g = 0; g < world; ++g)
            for (int i = 0; i < cuts[g + 1] - cuts[g]; ++i)
                y_out[(size_t)st * rows_global + cuts[g] + i] = use_f32 ? (double)yf[g][i] : yd[g][i];
    }
    for (int g = 0; g < world; ++g)
        if (epoch[g] != (unsigned long long)steps) return 1;
    return 0;
}

// thread resume order inside a block: 0 ascending, 1 descending, 2 random per pass
void emu_set_schedule(int mode) { emu::set_schedule(mode); }

// coordinates of arbitrary diagonals through the device search routine (merge_common.cuh)
int emu_merge_path_search(const int* ro, int rows, int nnz, const int* diagonals, int n, int* coords)
{
    emu::launch((unsigned)((n + 127) / 128), 128, [&] {
        diagonal_search_kernel(ro + 1, rows, nnz, diagonals, n, reinterpret_cast<int2*>(coords));
    });
    return 0;
}
}
