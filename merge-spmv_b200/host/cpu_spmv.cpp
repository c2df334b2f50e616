// cpu_spmv.cpp -- the cpu_spmv driver surface of dumerrill/merge-spmv (CPU-only plumbing,
// BASELINE.json configs[0]): OpenMP merge-based CsrMV on the host, same flags / protocol / output as
// the reference (file:line in /root/reference): main cpu_spmv.cpp:682-747, RunTests :537-675,
// TestOmpMergeCsrmv :362-406 (poison y, checked call, 3 warm calls, timed loop), DisplayPerf :502-528.
// This binary is a separate CPU tool; nothing in libmergespmv.so or gpu_spmv calls into it, and it
// does not use oracle/.  The MKL comparator column is gone (third-party, not in the image).
#include <omp.h>

#include "matrix.hpp"

using namespace mspmv_host;

static bool g_quiet = false;
static bool g_verbose2 = false;  // --v2: display the input matrix (cpu_spmv.cpp:74,604,722); --v is parsed and unused there (:73,721)
static int g_omp_threads = -1;

static inline void merge_path_search(int diagonal, const int* row_end_offsets, int a_len, int b_len, int& x, int& y)
{
    int lo = std::max(diagonal - b_len, 0), hi = std::min(diagonal, a_len);
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (row_end_offsets[mid] <= diagonal - mid - 1) lo = mid + 1;
        else hi = mid;
    }
    x = std::min(lo, a_len);
    y = diagonal - lo;
}

// One equal share of the rows+nnz merge path per thread; whole rows are stored, the trailing
// partial row is carried and folded in serially afterwards (cpu_spmv.cpp:292-353).
template <typename V>
static void omp_merge_csrmv(int num_threads, const Csr<V>& a, const V* __restrict x, V* __restrict y,
                            std::vector<int>& carry_row, std::vector<V>& carry_val)
{
    const int* __restrict row_end = a.row_offsets.data() + 1;
    const int* __restrict col = a.column_indices.data();
    const V* __restrict val = a.values.data();
    const int total = a.num_rows + a.num_nonzeros;
    const int share = (total + num_threads - 1) / num_threads;
#pragma omp parallel for schedule(static) num_threads(num_threads)
    for (int tid = 0; tid < num_threads; ++tid) {
        int d0 = (int)std::min<long long>((long long)share * tid, total);
        int d1 = (int)std::min<long long>((long long)d0 + share, total);
        int cx, cy, ex, ey;
        merge_path_search(d0, row_end, a.num_rows, a.num_nonzeros, cx, cy);
        merge_path_search(d1, row_end, a.num_rows, a.num_nonzeros, ex, ey);
        for (; cx < ex; ++cx) {
            V acc = V(0);
            for (; cy < row_end[cx]; ++cy) acc += val[cy] * x[col[cy]];
            y[cx] = acc;
        }
        V tail = V(0);
        for (; cy < ey; ++cy) tail += val[cy] * x[col[cy]];
        carry_row[tid] = ex;
        carry_val[tid] = tail;
    }
    for (int tid = 0; tid < num_threads - 1; ++tid)
        if (carry_row[tid] < a.num_rows) y[carry_row[tid]] += carry_val[tid];
}

template <typename V>
static void display_perf(double setup_ms, double avg_ms, const Csr<V>& a)  // cpu_spmv.cpp:502-528
{
    size_t total_bytes = (size_t)a.num_nonzeros * (sizeof(V) * 2 + sizeof(int)) + (size_t)a.num_rows * (sizeof(int) + sizeof(V));
    double nz_throughput = double(a.num_nonzeros) / avg_ms / 1.0e6;
    double effective_bandwidth = double(total_bytes) / avg_ms / 1.0e6;
    if (!g_quiet)
        std::printf("fp%d: %.4f setup ms, %.4f avg ms, %.5f gflops, %.3lf effective GB/s\n", (int)sizeof(V) * 8, setup_ms,
                    avg_ms, 2 * nz_throughput, effective_bandwidth);
    else
        std::printf("%.5f, %.5f, %.6f, %.3lf, ", setup_ms, avg_ms, 2 * nz_throughput, effective_bandwidth);
    std::fflush(stdout);
}

template <typename V>
static void run_tests(const CommandLineArgs& args, V alpha, V beta, int timing_iterations)
{
    Csr<V> a = build_from_args<V>(args, g_quiet, false);
    maybe_dump_csr(args, a);
    stats(a).display(!g_quiet);
    if (!g_quiet) {
        std::printf("\n");
        display_histogram(a);
        std::printf("\n");
        if (g_verbose2) display_matrix(a);
        std::printf("\n");
    }
    std::fflush(stdout);
    if (timing_iterations == -1) {  // cpu_spmv.cpp:609-615: printed after the statistics, only when adaptive
        timing_iterations = (int)std::min(200000ull, std::max(100ull, (16ull << 30) / (unsigned long long)std::max(a.num_nonzeros, 1)));
        if (!g_quiet) std::printf("\t%d timing iterations\n", timing_iterations);
    }
    std::vector<V> x(a.num_cols, V(1)), y_in(a.num_rows, V(1)), y_ref(a.num_rows), y(a.num_rows);
    if (args.CheckCmdLineFlag("randx"))
        for (int c = 0; c < a.num_cols; ++c) x[c] = (V)hashed_value((uint64_t)c, 0x5EED00FFull);
    spmv_gold(a, x.data(), y_in.data(), y_ref.data(), alpha, beta);

    if (!g_quiet) std::printf("\n\n");
    std::printf("Merge CsrMV, ");
    std::fflush(stdout);
    if (g_omp_threads == -1) g_omp_threads = omp_get_num_procs();
    if (!g_quiet) std::printf("\tUsing %d threads on %d procs\n", g_omp_threads, omp_get_num_procs());
    std::vector<int> carry_row(g_omp_threads);
    std::vector<V> carry_val(g_omp_threads);
    std::memset(y.data(), -1, sizeof(V) * a.num_rows);
    omp_merge_csrmv(g_omp_threads, a, x.data(), y.data(), carry_row, carry_val);
    if (!g_quiet) {
        int compare = compare_results(y.data(), y_ref.data(), a.num_rows, true);
        std::printf("\t%s\n", compare ? "FAIL" : "PASS");
        std::fflush(stdout);
    }
    for (int w = 0; w < 3; ++w) omp_merge_csrmv(g_omp_threads, a, x.data(), y.data(), carry_row, carry_val);
    double t0 = omp_get_wtime();
    for (int it = 0; it < timing_iterations; ++it) omp_merge_csrmv(g_omp_threads, a, x.data(), y.data(), carry_row, carry_val);
    double avg_ms = (omp_get_wtime() - t0) * 1000.0 / timing_iterations;
    display_perf(0.0, avg_ms, a);
}

int main(int argc, char** argv)
{
    CommandLineArgs args(argc, argv);
    if (args.CheckCmdLineFlag("help")) {
        std::printf("%s [--quiet] [--v] [--threads=<OMP threads>] [--i=<timing iterations>] [--fp32] "
                    "[--alpha=<alpha scalar (default: 1.0)>] [--beta=<beta scalar (default: 0.0)>]\n"
                    "\t--mtx=<matrix market file>\n\t--dense=<cols>\n\t--grid2d=<width>\n\t--grid3d=<width>\n\t--wheel=<spokes>\n"
                    "\t--uniform=<nnz per row> [--rows=] [--cols=]\n\t--powerlaw=<max row length> [--rows=] [--cols=] [--nnz=]\n"
                    "\t--banded=<half bandwidth> [--rows=]\n\t[--values=ones|random] [--randx] [--seed=]\n",
                    argv[0]);
        return 0;
    }
    int timing_iterations = -1;
    float alpha = 1.0f, beta = 0.0f;
    g_quiet = args.CheckCmdLineFlag("quiet");
    (void)args.CheckCmdLineFlag("v");
    g_verbose2 = args.CheckCmdLineFlag("v2");
    const bool fp32 = args.CheckCmdLineFlag("fp32");
    args.GetCmdLineArgument("i", timing_iterations);
    args.GetCmdLineArgument("threads", g_omp_threads);
    args.GetCmdLineArgument("alpha", alpha);
    args.GetCmdLineArgument("beta", beta);
    if (fp32) run_tests<float>(args, alpha, beta, timing_iterations);
    else run_tests<double>(args, (double)alpha, (double)beta, timing_iterations);
    std::printf("\n");
    return 0;
}
