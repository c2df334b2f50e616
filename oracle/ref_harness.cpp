/*
 * oracle/ref_harness.cpp -- links the REFERENCE's own CPU code as a library.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  This translation unit contains no SpMV logic of its
 * own: it #includes /root/reference/cpu_spmv.cpp where it lies (found through -I, never
 * copied into this repository) with `main` renamed, and exports thin extern "C" wrappers
 * around the reference's MergePathSearch (cpu_spmv.cpp:223), OmpMergeCsrmv (:292),
 * SpmvGold (:257), TestOmpMergeCsrmv (:362), CompareResults (utils.h:692,720) and the
 * CooMatrix/CsrMatrix builders (sparse_matrix.h:217-617,666-728).
 *
 * Built only where /root/reference exists (this container) by oracle/Makefile into
 * oracle/_ref/libref_cpu_spmv*.so; the .so travels to the GPU box, the sources do not.
 * Used by tests/ (to pin oracle/merge_oracle.c and to produce tests/golden/) and by
 * bench.py's cpu_baseline / --impl reference leg ("kind": "reference").
 */
#define main reference_cpu_spmv_main
#include "cpu_spmv.cpp"
#undef main

#include <cstring>

namespace {

// CsrMatrix only has a from-COO constructor (sparse_matrix.h:766-771).  Wrap caller arrays
// without copying: build from an empty COO, free its zero-length arrays, point the fields
// at the caller's memory, and detach again before destruction.
template <typename V>
struct BorrowedCsr {
    CooMatrix<V, int> empty;
    CsrMatrix<V, int> csr;
    BorrowedCsr(int rows, int cols, int nnz, int* row_offsets, int* col, V* val) : csr(empty) {
        csr.Clear();
        csr.num_rows = rows;
        csr.num_cols = cols;
        csr.num_nonzeros = nnz;
        csr.row_offsets = row_offsets;
        csr.column_indices = col;
        csr.values = val;
    }
    ~BorrowedCsr() {
        csr.row_offsets = NULL;
        csr.column_indices = NULL;
        csr.values = NULL;
    }
};

template <typename V>
void csrmv(int threads, int rows, int cols, int nnz, int* row_offsets, int* col, V* val, V* x, V* y) {
    BorrowedCsr<V> m(rows, cols, nnz, row_offsets, col, val);
    OmpMergeCsrmv(threads, m.csr, row_offsets + 1, col, val, x, y);
}

template <typename V>
void gold(int rows, int cols, int nnz, int* row_offsets, int* col, V* val, V* x, V* y_in, V* y_out,
          V alpha, V beta) {
    BorrowedCsr<V> m(rows, cols, nnz, row_offsets, col, val);
    SpmvGold(m.csr, x, y_in, y_out, alpha, beta);
}

template <typename V>
double timed(int threads, int rows, int cols, int nnz, int* row_offsets, int* col, V* val, V* x,
             V* y_ref, V* y, int iterations) {
    BorrowedCsr<V> m(rows, cols, nnz, row_offsets, col, val);
    bool quiet = g_quiet;
    int saved_threads = g_omp_threads;
    g_quiet = true;
    g_omp_threads = threads;
    float setup_ms = 0;
    float avg_ms = TestOmpMergeCsrmv(m.csr, x, y_ref, y, iterations, setup_ms);
    g_quiet = quiet;
    g_omp_threads = saved_threads;
    return avg_ms;
}

// COO -> CSR through the reference's own CsrMatrix::Init, copied out to caller buffers.
template <typename V>
struct BuiltCsr {
    CooMatrix<V, int> coo;
    CsrMatrix<V, int>* csr;
    BuiltCsr() : csr(NULL) {}
    ~BuiltCsr() { delete csr; }
    void finish() { csr = new CsrMatrix<V, int>(coo); }
};

BuiltCsr<double>* g_built = NULL;

}  // namespace

extern "C" {

void ref_merge_path_search(int diagonal, int* row_end_offsets, int a_len, int b_len, int* out_xy) {
    CountingInputIterator<int> nonzero_indices(0);
    int2 c;
    MergePathSearch(diagonal, row_end_offsets, nonzero_indices, a_len, b_len, c);
    out_xy[0] = c.x;
    out_xy[1] = c.y;
}

void ref_omp_merge_csrmv_f32(int threads, int rows, int cols, int nnz, int* row_offsets, int* col,
                             float* val, float* x, float* y) {
    csrmv<float>(threads, rows, cols, nnz, row_offsets, col, val, x, y);
}
void ref_omp_merge_csrmv_f64(int threads, int rows, int cols, int nnz, int* row_offsets, int* col,
                             double* val, double* x, double* y) {
    csrmv<double>(threads, rows, cols, nnz, row_offsets, col, val, x, y);
}

void ref_spmv_gold_f32(int rows, int cols, int nnz, int* row_offsets, int* col, float* val, float* x,
                       float* y_in, float* y_out, float alpha, float beta) {
    gold<float>(rows, cols, nnz, row_offsets, col, val, x, y_in, y_out, alpha, beta);
}
void ref_spmv_gold_f64(int rows, int cols, int nnz, int* row_offsets, int* col, double* val,
                       double* x, double* y_in, double* y_out, double alpha, double beta) {
    gold<double>(rows, cols, nnz, row_offsets, col, val, x, y_in, y_out, alpha, beta);
}

// Average ms per call in the reference's own TestOmpMergeCsrmv protocol (cpu_spmv.cpp:362-406).
double ref_time_omp_merge_csrmv_f32(int threads, int rows, int cols, int nnz, int* row_offsets,
                                    int* col, float* val, float* x, float* y_ref, float* y,
                                    int iterations) {
    return timed<float>(threads, rows, cols, nnz, row_offsets, col, val, x, y_ref, y, iterations);
}
double ref_time_omp_merge_csrmv_f64(int threads, int rows, int cols, int nnz, int* row_offsets,
                                    int* col, double* val, double* x, double* y_ref, double* y,
                                    int iterations) {
    return timed<double>(threads, rows, cols, nnz, row_offsets, col, val, x, y_ref, y, iterations);
}

int ref_compare_results_f32(float* computed, float* reference, int len) {
    return CompareResults(computed, reference, len, false);
}
int ref_compare_results_f64(double* computed, double* reference, int len) {
    return CompareResults(computed, reference, len, false);
}

int ref_num_procs(void) { return omp_get_num_procs(); }

// The reference driver's own main() (cpu_spmv.cpp:682-747), for pinning the CLI / CSV contract.
int ref_cpu_spmv_main(int argc, char** argv) { return reference_cpu_spmv_main(argc, argv); }

// ---- matrix builders (fp64 values): kind 0 = Matrix-Market file, 1 = grid2d(w), 2 = grid3d(w),
// 3 = wheel(spokes), 4 = dense(rows=a, cols=b).  Returns 0 and fills dims; then ref_built_copy.
int ref_build(int kind, const char* path, int a, int b, int* dims /* rows, cols, nnz */) {
    delete g_built;
    g_built = new BuiltCsr<double>();
    switch (kind) {
        case 0: g_built->coo.InitMarket(std::string(path), 1.0, false); break;
        case 1: g_built->coo.InitGrid2d(a, false); break;
        case 2: g_built->coo.InitGrid3d(a, false); break;
        case 3: g_built->coo.InitWheel(a); break;
        case 4: g_built->coo.InitDense(a, b); break;
        default: return 1;
    }
    g_built->finish();
    dims[0] = g_built->csr->num_rows;
    dims[1] = g_built->csr->num_cols;
    dims[2] = g_built->csr->num_nonzeros;
    return 0;
}
void ref_built_copy(int* row_offsets, int* col, double* val) {
    CsrMatrix<double, int>& m = *g_built->csr;
    std::memcpy(row_offsets, m.row_offsets, sizeof(int) * (size_t)(m.num_rows + 1));
    std::memcpy(col, m.column_indices, sizeof(int) * (size_t)m.num_nonzeros);
    std::memcpy(val, m.values, sizeof(double) * (size_t)m.num_nonzeros);
}
// Row-length statistics as the reference prints them (sparse_matrix.h:786-913).
void ref_built_stats(double* out /* mean, std_dev, variation, skewness */) {
    GraphStats s = g_built->csr->Stats();
    out[0] = s.row_length_mean;
    out[1] = s.row_length_std_dev;
    out[2] = s.row_length_variation;
    out[3] = s.row_length_skewness;
}
void ref_built_free(void) {
    delete g_built;
    g_built = NULL;
}

}  // extern "C"
