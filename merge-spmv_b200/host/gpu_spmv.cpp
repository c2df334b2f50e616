// gpu_spmv.cpp -- the gpu_spmv driver surface of dumerrill/merge-spmv on top of libmergespmv.so.
//
// Same flag grammar, test protocol and output as the reference driver (file:line in
// /root/reference): main gpu_spmv.cu:671-741, RunTests :598-664, RunTest :484-590,
// TestGpuMergeCsrmv :376-435 (size query, warm-up + check, timed loop between two cudaEvents),
// DisplayPerf :445-474 (gflops = 2*nnz/ms/1e6; "effective" bytes = nnz*(2*sizeof(V)+4) +
// rows*(4+sizeof(V))), DeviceInit banner utils.h:451-515 (peak GB/s = busWidth*memClock*2/8).
// Differences: the legacy cuSPARSE csrmv/hybmv comparators (removed from CUDA 12) are replaced by
// one optional modern comparator, cusparseSpMV CSR_ALG2 (--cusparse); --alpha/--beta are honoured
// by the merge kernel (the reference's ignores them and would print FAIL); new synthetic
// families --uniform/--powerlaw/--banded; a line with the algorithmic-bytes roofline fraction.
#include <cuda_runtime.h>

#include <chrono>
#include <cusparse.h>

#include "../../include/mergespmv_cub_shim.hpp"
#include "matrix.hpp"

using namespace mspmv_host;

static bool g_quiet = false;
static bool g_verbose = false;   // --v: display the reference and computed vectors (gpu_spmv.cu:59,711; utils.h:785-798)
static bool g_verbose2 = false;  // --v2: display the input matrix (gpu_spmv.cu:60,510,712)

#define CUDA_EXIT(e)                                                                              \
    do {                                                                                          \
        cudaError_t _e = (e);                                                                     \
        if (_e != cudaSuccess) {                                                                  \
            std::fprintf(stderr, "CUDA error %d [%s, %d]: %s\n", (int)_e, __FILE__, __LINE__,    \
                         cudaGetErrorString(_e));                                                 \
            std::exit(1);                                                                         \
        }                                                                                         \
    } while (0)

struct WallTimer {  // host wall clock (utils.h:528-553 uses omp_get_wtime for the same purpose)
    std::chrono::steady_clock::time_point a, b;
    void Start() { a = std::chrono::steady_clock::now(); }
    void Stop() { b = std::chrono::steady_clock::now(); }
    float ElapsedMillis() const { return std::chrono::duration<float, std::milli>(b - a).count(); }
};

struct GpuTimer {  // utils.h:624-658
    cudaEvent_t start, stop;
    GpuTimer()
    {
        cudaEventCreate(&start);
        cudaEventCreate(&stop);
    }
    ~GpuTimer()
    {
        cudaEventDestroy(start);
        cudaEventDestroy(stop);
    }
    void Start() { cudaEventRecord(start, 0); }
    void Stop() { cudaEventRecord(stop, 0); }
    float ElapsedMillis()
    {
        float ms;
        cudaEventSynchronize(stop);
        cudaEventElapsedTime(&ms, start, stop);
        return ms;
    }
};

template <typename V>
struct DeviceProblem {
    V *d_values = nullptr, *d_x = nullptr, *d_y = nullptr;
    int *d_row_offsets = nullptr, *d_col = nullptr;
};

template <typename V>
static int compare_device(const V* h_reference, const V* d_data, int n)  // utils.h:771-808
{
    std::vector<V> h(n);
    CUDA_EXIT(cudaMemcpy(h.data(), d_data, sizeof(V) * n, cudaMemcpyDeviceToHost));
    if (g_verbose) {  // display_data of CompareDeviceResults (utils.h:785-798)
        std::printf("Reference:\n");
        for (int i = 0; i < n; ++i) std::printf("%g, ", (double)h_reference[i]);
        std::printf("\n\nComputed:\n");
        for (int i = 0; i < n; ++i) std::printf("%g, ", (double)h[i]);
        std::printf("\n\n");
    }
    return compare_results(h.data(), h_reference, n, true);
}

// TestGpuMergeCsrmv, gpu_spmv.cu:376-435
template <typename V>
static float test_merge_csrmv(const Csr<V>& a, const V* y_in, const V* y_ref, DeviceProblem<V>& p, V alpha, V beta,
                              int timing_iterations, float& setup_ms)
{
    setup_ms = 0.0f;
    const bool axpby = !(alpha == V(1) && beta == V(0));
    auto call = [&](void* temp, size_t& bytes, bool debug) {
        return axpby ? cub::DeviceSpmv::CsrMV(temp, bytes, p.d_values, p.d_row_offsets, p.d_col, p.d_x, p.d_y,
                                              a.num_rows, a.num_cols, a.num_nonzeros, alpha, beta, (cudaStream_t)0, debug)
                     : cub::DeviceSpmv::CsrMV(temp, bytes, p.d_values, p.d_row_offsets, p.d_col, p.d_x, p.d_y,
                                              a.num_rows, a.num_cols, a.num_nonzeros, (cudaStream_t)0, debug);
    };
    size_t temp_bytes = 0;
    void* d_temp = nullptr;
    CUDA_EXIT(call(nullptr, temp_bytes, false));
    CUDA_EXIT(cudaMalloc(&d_temp, temp_bytes));
    CUDA_EXIT(cudaMemcpy(p.d_y, y_in, sizeof(V) * a.num_rows, cudaMemcpyHostToDevice));
    CUDA_EXIT(call(d_temp, temp_bytes, !g_quiet));  // warm-up, debug_synchronous = !quiet
    if (!g_quiet) {
        int compare = compare_device(y_ref, p.d_y, a.num_rows);
        std::printf("\t%s\n", compare ? "FAIL" : "PASS");
        std::fflush(stdout);
    }
    GpuTimer timer;
    timer.Start();
    for (int it = 0; it < timing_iterations; ++it) CUDA_EXIT(call(d_temp, temp_bytes, false));
    timer.Stop();
    float elapsed = timer.ElapsedMillis();
    CUDA_EXIT(cudaFree(d_temp));
    return elapsed / timing_iterations;
}

// Same-box yardstick in the slot the reference used for cuSPARSE csrmv/hybmv (gpu_spmv.cu:568-578)
template <typename V>
static float test_cusparse_spmv(const Csr<V>& a, const V* y_in, const V* y_ref, DeviceProblem<V>& p, V alpha, V beta,
                                int timing_iterations, float& setup_ms)
{
    cusparseHandle_t h;
    cusparseCreate(&h);
    const cudaDataType dt = sizeof(V) == 8 ? CUDA_R_64F : CUDA_R_32F;
    cusparseSpMatDescr_t mat;
    cusparseDnVecDescr_t vx, vy;
    cusparseCreateCsr(&mat, a.num_rows, a.num_cols, a.num_nonzeros, p.d_row_offsets, p.d_col, p.d_values,
                      CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, dt);
    cusparseCreateDnVec(&vx, a.num_cols, p.d_x, dt);
    cusparseCreateDnVec(&vy, a.num_rows, p.d_y, dt);
    size_t bytes = 0;
    void* buf = nullptr;
    GpuTimer setup;
    setup.Start();
    cusparseSpMV_bufferSize(h, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, mat, vx, &beta, vy, dt,
                            CUSPARSE_SPMV_CSR_ALG2, &bytes);
    CUDA_EXIT(cudaMalloc(&buf, bytes ? bytes : 1));
    setup.Stop();
    setup_ms = setup.ElapsedMillis();
    CUDA_EXIT(cudaMemcpy(p.d_y, y_in, sizeof(V) * a.num_rows, cudaMemcpyHostToDevice));
    cusparseSpMV(h, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, mat, vx, &beta, vy, dt, CUSPARSE_SPMV_CSR_ALG2, buf);
    if (!g_quiet) {
        int compare = compare_device(y_ref, p.d_y, a.num_rows);
        std::printf("\t%s\n", compare ? "FAIL" : "PASS");
        std::fflush(stdout);
    }
    GpuTimer timer;
    timer.Start();
    for (int it = 0; it < timing_iterations; ++it)
        cusparseSpMV(h, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, mat, vx, &beta, vy, dt, CUSPARSE_SPMV_CSR_ALG2, buf);
    timer.Stop();
    float elapsed = timer.ElapsedMillis();
    cudaFree(buf);
    cusparseDestroySpMat(mat);
    cusparseDestroyDnVec(vx);
    cusparseDestroyDnVec(vy);
    cusparseDestroy(h);
    return elapsed / timing_iterations;
}

// Second yardstick: the toolkit's own cub::DeviceSpmv::CsrMV (host/cub_comparator.cu), y = A*x only
extern "C" int toolkit_cub_csrmv_f64(void*, size_t*, const double*, const int*, const int*, const double*, double*, int,
                                     int, int, cudaStream_t);
extern "C" int toolkit_cub_csrmv_f32(void*, size_t*, const float*, const int*, const int*, const float*, float*, int,
                                     int, int, cudaStream_t);
static int toolkit_cub_csrmv(void* t, size_t* b, const double* v, const int* ro, const int* ci, const double* x,
                             double* y, int r, int c, int n)
{
    return toolkit_cub_csrmv_f64(t, b, v, ro, ci, x, y, r, c, n, 0);
}
static int toolkit_cub_csrmv(void* t, size_t* b, const float* v, const int* ro, const int* ci, const float* x, float* y,
                             int r, int c, int n)
{
    return toolkit_cub_csrmv_f32(t, b, v, ro, ci, x, y, r, c, n, 0);
}

template <typename V>
static float test_toolkit_cub(const Csr<V>& a, const V* y_in, const V* y_ref, DeviceProblem<V>& p, int timing_iterations,
                              float& setup_ms)
{
    setup_ms = 0.0f;
    size_t temp_bytes = 0;
    void* d_temp = nullptr;
    auto call = [&](void* temp) {
        return (cudaError_t)toolkit_cub_csrmv(temp, &temp_bytes, p.d_values, p.d_row_offsets, p.d_col, p.d_x, p.d_y,
                                              a.num_rows, a.num_cols, a.num_nonzeros);
    };
    CUDA_EXIT(call(nullptr));
    CUDA_EXIT(cudaMalloc(&d_temp, temp_bytes ? temp_bytes : 1));
    CUDA_EXIT(cudaMemcpy(p.d_y, y_in, sizeof(V) * a.num_rows, cudaMemcpyHostToDevice));
    CUDA_EXIT(call(d_temp));
    if (!g_quiet) {
        int compare = compare_device(y_ref, p.d_y, a.num_rows);
        std::printf("\t%s\n", compare ? "FAIL" : "PASS");
        std::fflush(stdout);
    }
    GpuTimer timer;
    timer.Start();
    for (int it = 0; it < timing_iterations; ++it) CUDA_EXIT(call(d_temp));
    timer.Stop();
    float elapsed = timer.ElapsedMillis();
    CUDA_EXIT(cudaFree(d_temp));
    return elapsed / timing_iterations;
}

// ---- one process, several GPUs (new surface: the reference is single-GPU, README.md:5) ---------
// mspmv_mg_session_* (include/mergespmv.h section 5): the matrix is cut into merge-path shards exactly
// like OmpMergeCsrmv cuts it between CPU threads (cpu_spmv.cpp:311-321); device g runs the unchanged
// single-GPU CsrMV on its shard and then the NVLink carry-exchange kernel -- no NCCL.  The driver keeps
// the reference's protocol: one warm-up product checked against SpmvGold, then `timing_iterations`
// device-timed products (x resident), max over devices.
template <typename V>
static float test_merge_csrmv_multi(const Csr<V>& a, const V* x_host, const V* y_ref, int p, int timing_iterations,
                                    float& setup_ms)
{
    mspmv_mg_session* s = nullptr;
    WallTimer setup;
    setup.Start();
    CUDA_EXIT((cudaError_t)mspmv_mg_session_create(&s, p, nullptr, (int)sizeof(V), a.num_rows, a.num_cols, a.num_nonzeros,
                                                   a.row_offsets.data(), a.column_indices.data(), a.values.data()));
    setup.Stop();
    setup_ms = setup.ElapsedMillis();
    V *xh = nullptr, *yh = nullptr;
    CUDA_EXIT((cudaError_t)mspmv_host_alloc((void**)&xh, sizeof(V) * (size_t)a.num_cols));
    CUDA_EXIT((cudaError_t)mspmv_host_alloc((void**)&yh, sizeof(V) * (size_t)a.num_rows));
    std::memcpy(xh, x_host, sizeof(V) * (size_t)a.num_cols);
    CUDA_EXIT((cudaError_t)mspmv_mg_session_apply(s, xh, yh));  // warm-up + check
    if (!g_quiet) {
        int compare = compare_results(yh, y_ref, a.num_rows, true);
        std::printf("\t%s\n", compare ? "FAIL" : "PASS");
        std::fflush(stdout);
    }
    float ms = 0.f;
    CUDA_EXIT((cudaError_t)mspmv_mg_session_time_device(s, timing_iterations, &ms));
    // end to end with host buffers: 3-slot pipeline over the same x (what a solver's host loop would see)
    const int n_e2e = 24;
    V *xs = nullptr, *ys = nullptr;
    CUDA_EXIT((cudaError_t)mspmv_host_alloc((void**)&xs, sizeof(V) * (size_t)a.num_cols * n_e2e));
    CUDA_EXIT((cudaError_t)mspmv_host_alloc((void**)&ys, sizeof(V) * (size_t)a.num_rows * n_e2e));
    for (int i = 0; i < n_e2e; ++i) std::memcpy(xs + (size_t)i * a.num_cols, x_host, sizeof(V) * (size_t)a.num_cols);
    CUDA_EXIT((cudaError_t)mspmv_mg_session_apply_many(s, 3, xs, ys));
    WallTimer e2e;
    e2e.Start();
    CUDA_EXIT((cudaError_t)mspmv_mg_session_apply_many(s, n_e2e, xs, ys));
    e2e.Stop();
    if (!g_quiet)
        std::printf("\thost buffers, pipelined (mspmv_mg_session_apply_many): %.4f ms per product, %.2f gflops%s\n",
                    e2e.ElapsedMillis() / n_e2e, 2.0 * a.num_nonzeros / (e2e.ElapsedMillis() / n_e2e) / 1e6,
                    std::memcmp(ys + (size_t)(n_e2e - 1) * a.num_rows, yh, sizeof(V) * (size_t)a.num_rows) ? "  MISMATCH" : "");
    mspmv_host_free(xs), mspmv_host_free(ys), mspmv_host_free(xh), mspmv_host_free(yh);
    mspmv_mg_session_destroy(s);
    return ms;
}

template <typename V>
static void display_perf(float device_giga_bandwidth, double setup_ms, double avg_ms, const Csr<V>& a)
{
    size_t total_bytes = (size_t)a.num_nonzeros * (sizeof(V) * 2 + sizeof(int)) + (size_t)a.num_rows * (sizeof(int) + sizeof(V));
    double nz_throughput = double(a.num_nonzeros) / avg_ms / 1.0e6;
    double effective_bandwidth = double(total_bytes) / avg_ms / 1.0e6;
    if (!g_quiet) {
        std::printf("fp%d: %.4f setup ms, %.4f avg ms, %.5f gflops, %.3lf effective GB/s (%.2f%% peak)\n",
                    (int)sizeof(V) * 8, setup_ms, avg_ms, 2 * nz_throughput, effective_bandwidth,
                    effective_bandwidth / device_giga_bandwidth * 100);
        // compulsory-traffic roofline: every array once
        double alg = double(a.num_nonzeros) * (sizeof(V) + 4) + (a.num_rows + 1.0) * 4 + double(a.num_rows) * sizeof(V) +
                     double(a.num_cols) * sizeof(V);
        std::printf("\talgorithmic %.3f GB/s (%.2f%% of %.1f GB/s peak)\n", alg / avg_ms / 1.0e6,
                    alg / avg_ms / 1.0e6 / device_giga_bandwidth * 100, device_giga_bandwidth);
    } else {
        std::printf("%.5f, %.5f, %.6f, %.3lf, ", setup_ms, avg_ms, 2 * nz_throughput, effective_bandwidth);
    }
    std::fflush(stdout);
}

template <typename V>
static void run_tests(const CommandLineArgs& args, V alpha, V beta, int timing_iterations, const cudaDeviceProp& prop,
                      float device_giga_bandwidth)
{
    Csr<V> a = build_from_args<V>(args, g_quiet, true);
    maybe_dump_csr(args, a);
    if (timing_iterations == -1)  // gpu_spmv.cu:493
        timing_iterations = (int)std::min(50000ull, std::max(100ull, (16ull << 30) / (unsigned long long)std::max(a.num_nonzeros, 1)));
    if (!g_quiet) std::printf("\t%d timing iterations\n", timing_iterations);
    stats(a).display(!g_quiet);
    if (!g_quiet) {
        std::printf("\n");
        display_histogram(a);
        std::printf("\n");
        if (g_verbose2) display_matrix(a);
        std::printf("\n");
    }
    std::fflush(stdout);

    std::vector<V> x(a.num_cols, V(1)), y_in(a.num_rows, V(1)), y_ref(a.num_rows);
    if (args.CheckCmdLineFlag("randx"))
        for (int c = 0; c < a.num_cols; ++c) x[c] = (V)hashed_value((uint64_t)c, 0x5EED00FFull);
    spmv_gold(a, x.data(), y_in.data(), y_ref.data(), alpha, beta);

    if (g_quiet) {
        std::printf("%s, %s, ", prop.name, sizeof(V) > 4 ? "fp64" : "fp32");
        std::fflush(stdout);
    }
    DeviceProblem<V> p;
    CUDA_EXIT(cudaMalloc(&p.d_values, sizeof(V) * std::max(a.num_nonzeros, 1)));
    CUDA_EXIT(cudaMalloc(&p.d_row_offsets, sizeof(int) * (a.num_rows + 1)));
    CUDA_EXIT(cudaMalloc(&p.d_col, sizeof(int) * std::max(a.num_nonzeros, 1)));
    CUDA_EXIT(cudaMalloc(&p.d_x, sizeof(V) * a.num_cols));
    CUDA_EXIT(cudaMalloc(&p.d_y, sizeof(V) * a.num_rows));
    CUDA_EXIT(cudaMemcpy(p.d_values, a.values.data(), sizeof(V) * a.num_nonzeros, cudaMemcpyHostToDevice));
    CUDA_EXIT(cudaMemcpy(p.d_row_offsets, a.row_offsets.data(), sizeof(int) * (a.num_rows + 1), cudaMemcpyHostToDevice));
    CUDA_EXIT(cudaMemcpy(p.d_col, a.column_indices.data(), sizeof(int) * a.num_nonzeros, cudaMemcpyHostToDevice));
    CUDA_EXIT(cudaMemcpy(p.d_x, x.data(), sizeof(V) * a.num_cols, cudaMemcpyHostToDevice));

    float setup_ms, avg_ms;
    if (!g_quiet) std::printf("\n\n");
    std::printf("Merge-based CsrMV, ");
    std::fflush(stdout);
    avg_ms = test_merge_csrmv(a, y_in.data(), y_ref.data(), p, alpha, beta, timing_iterations, setup_ms);
    display_perf(device_giga_bandwidth, setup_ms, avg_ms, a);

    if (args.CheckCmdLineFlag("cusparse")) {
        if (!g_quiet) std::printf("\n\n");
        std::printf("cuSPARSE SpMV (CSR_ALG2), ");
        std::fflush(stdout);
        avg_ms = test_cusparse_spmv(a, y_in.data(), y_ref.data(), p, alpha, beta, timing_iterations, setup_ms);
        display_perf(device_giga_bandwidth, setup_ms, avg_ms, a);
    }
    if (args.CheckCmdLineFlag("cub")) {
        if (alpha != V(1) || beta != V(0)) {
            std::fprintf(stderr, "--cub: the toolkit's cub::DeviceSpmv computes y = A*x only; skipped\n");
        } else {
            if (!g_quiet) std::printf("\n\n");
            std::printf("CUDA toolkit cub::DeviceSpmv CsrMV, ");
            std::fflush(stdout);
            avg_ms = test_toolkit_cub(a, y_in.data(), y_ref.data(), p, timing_iterations, setup_ms);
            display_perf(device_giga_bandwidth, setup_ms, avg_ms, a);
        }
    }
    int gpus = 1;
    args.GetCmdLineArgument("gpus", gpus);
    if (gpus > 1) {
        int have = 0;
        CUDA_EXIT(cudaGetDeviceCount(&have));
        if (alpha != V(1) || beta != V(0) || have < gpus) {
            std::fprintf(stderr, "--gpus=%d: needs %d visible devices (have %d) and alpha=1, beta=0; skipped\n", gpus, gpus, have);
        } else {
            if (!g_quiet) std::printf("\n\n");
            std::printf("Merge-based CsrMV x%d GPUs (merge-path shards, NVLink carry exchange), ", gpus);
            std::fflush(stdout);
            avg_ms = test_merge_csrmv_multi(a, x.data(), y_ref.data(), gpus, timing_iterations, setup_ms);
            display_perf(device_giga_bandwidth * gpus, setup_ms, avg_ms, a);
        }
    }
    cudaFree(p.d_values);
    cudaFree(p.d_row_offsets);
    cudaFree(p.d_col);
    cudaFree(p.d_x);
    cudaFree(p.d_y);
}

int main(int argc, char** argv)
{
    CommandLineArgs args(argc, argv);
    if (args.CheckCmdLineFlag("help")) {
        std::printf("%s [--device=<device-id>] [--quiet] [--v] [--i=<timing iterations>] [--fp32] "
                    "[--alpha=<alpha scalar (default: 1.0)>] [--beta=<beta scalar (default: 0.0)>] [--cusparse] [--cub] [--gpus=<n>]\n"
                    "\t--mtx=<matrix market file>\n\t--dense=<cols> [--size=<nnz>]\n\t--grid2d=<width>\n\t--grid3d=<width>\n"
                    "\t--wheel=<spokes>\n\t--uniform=<nnz per row> [--rows=] [--cols=]\n"
                    "\t--powerlaw=<max row length> [--rows=] [--cols=] [--nnz=]\n\t--banded=<half bandwidth> [--rows=]\n"
                    "\t[--values=ones|random] [--randx] [--seed=]\n",
                    argv[0]);
        return 0;
    }
    int timing_iterations = -1, dev = 0;
    float alpha = 1.0f, beta = 0.0f;
    g_quiet = args.CheckCmdLineFlag("quiet");
    g_verbose = args.CheckCmdLineFlag("v");
    g_verbose2 = args.CheckCmdLineFlag("v2");
    const bool fp32 = args.CheckCmdLineFlag("fp32");
    args.GetCmdLineArgument("i", timing_iterations);
    args.GetCmdLineArgument("alpha", alpha);
    args.GetCmdLineArgument("beta", beta);
    args.GetCmdLineArgument("device", dev);

    // DeviceInit, utils.h:451-515
    int device_count = 0;
    CUDA_EXIT(cudaGetDeviceCount(&device_count));
    if (device_count == 0) {
        std::fprintf(stderr, "No devices supporting CUDA.\n");
        return 1;
    }
    if (dev > device_count - 1 || dev < 0) dev = 0;
    CUDA_EXIT(cudaSetDevice(dev));
    size_t free_mem = 0, total_mem = 0;
    CUDA_EXIT(cudaMemGetInfo(&free_mem, &total_mem));
    cudaDeviceProp prop;
    CUDA_EXIT(cudaGetDeviceProperties(&prop, dev));
    int mem_clock_khz = 0, bus_width = 0;
    cudaDeviceGetAttribute(&mem_clock_khz, cudaDevAttrMemoryClockRate, dev);
    cudaDeviceGetAttribute(&bus_width, cudaDevAttrGlobalMemoryBusWidth, dev);
    float device_giga_bandwidth = float(bus_width) * mem_clock_khz * 2 / 8 / 1000 / 1000;  // utils.h:491
    if (!g_quiet) {
        int ptx_version = 0;
        mspmv_ptx_version(&ptx_version);
        std::printf("Using device %d: %s (PTX version %d, SM%d, %d SMs, %lld free / %lld total MB physmem, %.3f GB/s @ %d kHz mem clock, ECC %s)\n",
                    dev, prop.name, ptx_version, prop.major * 100 + prop.minor * 10, prop.multiProcessorCount,
                    (long long)free_mem / 1024 / 1024, (long long)total_mem / 1024 / 1024, device_giga_bandwidth,
                    mem_clock_khz, prop.ECCEnabled ? "on" : "off");
        std::fflush(stdout);
    }
    if (fp32) run_tests<float>(args, alpha, beta, timing_iterations, prop, device_giga_bandwidth);
    else run_tests<double>(args, (double)alpha, (double)beta, timing_iterations, prop, device_giga_bandwidth);
    CUDA_EXIT(cudaDeviceSynchronize());
    std::printf("\n");
    return 0;
}
