// tsan_check.cpp -- the pipe-engine kernel under the SIMT interpreter as a ThreadSanitizer subject
// (built with -fsanitize=thread -DEMU_TSAN by tests/test_pipe_emu.py; TEST INFRASTRUCTURE ONLY).
// Every CUDA thread is a TSan fiber, barriers / warp collectives / mbarrier phases are release-acquire
// edges, so any shared- or global-memory access of the kernels that is not ordered by one of them is
// reported as a data race (exit code 66).  Results are also checked against a sequential SpMV.
#include "pipe_emu.cpp"

#include <cstdio>
#include <random>

extern "C" int emu_csrmv_f32(const float*, const int*, const int*, const float*, float*, int, int, float, float, int, int,
                             int*, int);

template <typename T>
static int run_case(int rows, int cols, double mean_len, double empty, int long_rows, unsigned seed)
{
    std::mt19937 rng(seed);
    std::vector<int> ro(rows + 1, 0), col;
    std::poisson_distribution<int> len(mean_len > 0 ? mean_len : 1e-9);
    std::uniform_real_distribution<double> u(0, 1);
    std::vector<int> lens(rows);
    for (int r = 0; r < rows; ++r) lens[r] = (mean_len > 0 && u(rng) >= empty) ? std::min(len(rng), cols) : 0;
    for (int k = 0; k < long_rows; ++k) lens[rng() % rows] = std::max(1, cols / 2);
    for (int r = 0; r < rows; ++r) {
        ro[r + 1] = ro[r] + lens[r];
        std::vector<int> pick(cols);
        for (int c = 0; c < cols; ++c) pick[c] = c;
        for (int i = 0; i < lens[r]; ++i) std::swap(pick[i], pick[i + rng() % (cols - i)]);
        std::sort(pick.begin(), pick.begin() + lens[r]);
        col.insert(col.end(), pick.begin(), pick.begin() + lens[r]);
    }
    const int nnz = ro[rows];
    std::vector<T> val(std::max(nnz, 1)), x(cols), want(rows, T(0));
    for (int i = 0; i < nnz; ++i) val[i] = T(1 + rng() % 4);
    for (int c = 0; c < cols; ++c) x[c] = T(1 + rng() % 4);
    for (int r = 0; r < rows; ++r)
        for (int k = ro[r]; k < ro[r + 1]; ++k) want[r] += val[k] * x[col[k]];
    if (col.empty()) col.push_back(0);
    int bad = 0;
    for (int mode = 0; mode < 4; ++mode) {  // grid of 1 / 3 / one block per tile, in-kernel search; 2 blocks, search kernel
        const int blocks = mode == 0 ? 1 : mode == 1 ? 3 : mode == 2 ? (1 << 20) : 2;
        const int search = mode != 3;
        std::vector<T> y(rows, T(-1));
        int stats[4];
        int rc;
        if constexpr (sizeof(T) == 8)
            rc = emu_pipe_f64(val.data(), ro.data(), col.data(), x.data(), y.data(), rows, nnz, 1.0, 0.0, 0, blocks, search,
                              nullptr, stats);
        else
            rc = emu_pipe_f32(val.data(), ro.data(), col.data(), x.data(), y.data(), rows, nnz, 1.f, 0.f, 0, blocks, search,
                              nullptr, stats);
        if (rc != 0 || y != want) {
            std::printf("MISMATCH rows=%d cols=%d nnz=%d mode=%d\n", rows, cols, nnz, mode);
            ++bad;
        }
    }
    return bad;
}

int main()
{
    int bad = 0;
    const struct { int rows, cols; double mean, empty; int longs; } cases[] = {
        {1, 1, 1.0, 0.0, 0}, {17, 3, 1.0, 0.3, 0}, {100, 64, 0.0, 1.0, 0}, {3000, 300, 0.05, 0.9, 0}, {1500, 2000, 9, 0.1, 2},
        {40, 6000, 700, 0.0, 1}, {1, 9000, 0, 0.0, 1}, {4000, 128, 2, 0.5, 0}, {600, 600, 31, 0.01, 0}};
    unsigned seed = 1;
    for (const auto& c : cases) {
        bad += run_case<double>(c.rows, c.cols, c.mean, c.empty, c.longs, seed++);
        bad += run_case<float>(c.rows, c.cols, c.mean, c.empty, c.longs, seed++);
    }
    for (int world : {2, 3, 8}) {  // the NVLink carry exchange, ranks simulated one after another
        std::vector<int> cuts(world + 1), rows_of(world);
        for (int g = 0; g <= world; ++g) cuts[g] = g * 5;
        for (int g = 0; g < world; ++g) rows_of[g] = cuts[g + 1];
        std::vector<double> carries(3 * world, 2.0), y(3 * (size_t)cuts[world]);
        if (emu_exchange_f64(world, 3, cuts.data(), rows_of.data(), carries.data(), y.data(), 0)) ++bad;
    }
    std::printf(bad ? "tsan check: %d result mismatches\n" : "tsan check complete\n", bad);
    return bad ? 1 : 0;
}
