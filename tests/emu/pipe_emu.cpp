// pipe_emu.cpp -- runs the pipe engine's kernel (merge-spmv_b200/csrc/spmv_pipe.cuh, compiled unchanged)
// under the host SIMT interpreter.  TEST INFRASTRUCTURE ONLY: built by tests/test_pipe_emu.py, never part
// of the product.  The launch mirrors pipe_launch_impl() in merge-spmv_b200/csrc/mergespmv.cu: a 4-byte
// memset of the ticket (or tile_search_kernel when search == 0), then ONE kernel.
#define MSPMV_PTX_HEADER "ptx_emu.cuh"
#include "simt_emu.hpp"

#include "merge_search.cuh"  // tile_search_kernel, diagonal_search_kernel
#include "spmv_pipe.cuh"
#include "carry_exchange.cuh"

#include <vector>

namespace {

using namespace mspmv;

template <typename T>
int shift_of(const void* p)
{
    return (int)((reinterpret_cast<uintptr_t>(p) & 15) / sizeof(T));
}

// kernel shape under test: <IPT fp64, IPT fp32, value-ring slots, column-ring slots, gather-ahead, consumer warps, first-in-register>
#ifndef EMU_PIPE_CFG
#define EMU_PIPE_CFG 9, 13, 2, 2, 0, 4, 0
#endif
template <typename T, int I64, int I32, int VST, int CST, int AHEAD, int NW, int FIR>
using CfgSel = PipeCfg<T, (sizeof(T) == 8 ? I64 : I32), VST, CST, AHEAD, NW, FIR>;

template <typename T, bool AXPBY>
int run(const T* values, const int* row_offsets, const int* col, const T* x, T* y, int num_rows, int num_nonzeros,
        T alpha, T beta, int max_blocks, int search, int* coords_out, int* stats)
{
    using C = CfgSel<T, EMU_PIPE_CFG>;
    if (num_rows <= 0) return 0;
    const int64_t merge_items = (int64_t)num_rows + num_nonzeros;
    const int num_tiles = (int)((merge_items + C::TILE - 1) / C::TILE);
    const int num_blocks = num_tiles < max_blocks ? num_tiles : max_blocks;
    // temporaries of exactly the size the kernel may touch, filled with garbage (the temp blob is uninitialised)
    std::vector<int2> coords_buf((size_t)num_tiles + 1);
    std::vector<int> cr_buf((size_t)num_blocks);
    std::vector<T> cv_buf((size_t)num_blocks);
    std::memset(coords_buf.data(), 0xEE, coords_buf.size() * sizeof(int2));
    std::memset(cr_buf.data(), 0xEE, cr_buf.size() * sizeof(int));
    std::memset(cv_buf.data(), 0xEE, cv_buf.size() * sizeof(T));
    unsigned int ticket = 0xEEEEEEEEu;
    if (stats) {
        stats[0] = num_tiles;
        stats[1] = C::TILE;
        stats[2] = C::THREADS;
        stats[3] = num_blocks;
    }
    if (!search) {
        emu::launch((unsigned)((num_tiles + 1 + 127) / 128), 128, [&] {
            tile_search_kernel(row_offsets + 1, num_rows, num_nonzeros, C::TILE, num_tiles, coords_buf.data(), &ticket);
        });
    } else {
        ticket = 0u;  // cudaMemsetAsync
    }
    int2* cout = reinterpret_cast<int2*>(coords_out);
    emu::launch((unsigned)num_blocks, (unsigned)C::THREADS, [&] {
        if (search)
            spmv_pipe_kernel<C, AXPBY, true>(values, row_offsets, col, x, y, nullptr, cout, cr_buf.data(), cv_buf.data(),
                                             &ticket, alpha, beta, num_rows, num_nonzeros, num_tiles,
                                             shift_of<T>(values), shift_of<int>(col), shift_of<int>(row_offsets));
        else
            spmv_pipe_kernel<C, AXPBY, false>(values, row_offsets, col, x, y, coords_buf.data(), cout, cr_buf.data(),
                                              cv_buf.data(), &ticket, alpha, beta, num_rows, num_nonzeros, num_tiles,
                                              shift_of<T>(values), shift_of<int>(col), shift_of<int>(row_offsets));
    });
    return 0;
}

}  // namespace

extern "C" {

int emu_pipe_f64(const double* v, const int* ro, const int* ci, const double* x, double* y, int rows, int nnz,
                 double alpha, double beta, int axpby, int max_blocks, int search, int* coords_out, int* stats)
{
    return axpby ? run<double, true>(v, ro, ci, x, y, rows, nnz, alpha, beta, max_blocks, search, coords_out, stats)
                 : run<double, false>(v, ro, ci, x, y, rows, nnz, alpha, beta, max_blocks, search, coords_out, stats);
}
int emu_pipe_f32(const float* v, const int* ro, const int* ci, const float* x, float* y, int rows, int nnz, float alpha,
                 float beta, int axpby, int max_blocks, int search, int* coords_out, int* stats)
{
    return axpby ? run<float, true>(v, ro, ci, x, y, rows, nnz, alpha, beta, max_blocks, search, coords_out, stats)
                 : run<float, false>(v, ro, ci, x, y, rows, nnz, alpha, beta, max_blocks, search, coords_out, stats);
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

int emu_merge_path_search(const int* ro, int rows, int nnz, const int* diagonals, int n, int* coords)
{
    emu::launch((unsigned)((n + 127) / 128), 128, [&] {
        diagonal_search_kernel(ro + 1, rows, nnz, diagonals, n, reinterpret_cast<int2*>(coords));
    });
    return 0;
}
}
