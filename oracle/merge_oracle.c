/*
 * oracle/merge_oracle.c -- CPU restatement of the reference's merge-based CsrMV.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this.  The product path
 * (libmergespmv.so, the gpu_spmv driver) never links or calls anything in oracle/.
 *
 * Parity is PINNED (see tests/test_oracle.py): this file is checked against
 *   - the reference's own code compiled from /root/reference (oracle/_ref, built by
 *     oracle/Makefile from ref_harness.cpp) on randomized CSR structures,
 *   - the known answers the reference carries: the 3x3 lattice in
 *     cub/device/device_spmv.cuh:90-123 and the paper's Fig. 8 example
 *     (coordinates (0,0),(2,2),(3,5),(4,8); y = [2,0,6,16]),
 *   - committed golden vectors under tests/golden/ generated from oracle/_ref by
 *     tests/golden/make_golden.py.
 *
 * All citations are file:line in /root/reference (dumerrill/merge-spmv @ 18895571).
 * Offsets are 32-bit ints like the reference (cpu_spmv.cpp:311,317 use int).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "merge_oracle.h"

/* ------------------------------------------------------------------------------------
 * MergePathSearch -- cpu_spmv.cpp:223-245 (same logic as cub/thread/thread_search.cuh:53-84).
 * List A = row_end_offsets[0..a_len), list B = the natural numbers 0..b_len (the
 * CountingInputIterator of cpu_spmv.cpp:94-206, so b[i] == i).  Finds where the given
 * diagonal crosses the merge path; ties consume A first (cpu_spmv.cpp:237).
 * ---------------------------------------------------------------------------------- */
void oracle_merge_path_search(int diagonal, const int* row_end_offsets, int a_len, int b_len,
                              int* out_x, int* out_y)
{
    int lo = diagonal - b_len;
    if (lo < 0) lo = 0;
    int hi = diagonal < a_len ? diagonal : a_len;

    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        /* b[diagonal - mid - 1] is simply the integer diagonal - mid - 1 */
        if (row_end_offsets[mid] <= diagonal - mid - 1)
            lo = mid + 1;
        else
            hi = mid;
    }
    *out_x = lo < a_len ? lo : a_len;
    *out_y = diagonal - lo;
}

/* Per-thread start/end diagonals -- cpu_spmv.cpp:311-318. */
static void thread_diagonals(int tid, int num_threads, int num_rows, int num_nonzeros,
                             int* start_diag, int* end_diag)
{
    int total = num_rows + num_nonzeros;
    int share = (total + num_threads - 1) / num_threads;
    long long s = (long long)share * tid; /* the reference multiplies in int; same value when it fits */
    int sd = s < total ? (int)s : total;
    long long e = (long long)sd + share;
    int ed = e < total ? (int)e : total;
    *start_diag = sd;
    *end_diag = ed;
}

/* coords[2*t], coords[2*t+1] = (x,y) where thread t starts; entry num_threads = end of the
 * last thread.  cpu_spmv.cpp:320-321. */
void oracle_merge_thread_coords(int num_threads, int num_rows, int num_nonzeros,
                                const int* row_end_offsets, int* coords)
{
    for (int t = 0; t <= num_threads; ++t) {
        int sd, ed;
        thread_diagonals(t < num_threads ? t : num_threads - 1, num_threads, num_rows,
                         num_nonzeros, &sd, &ed);
        int diag = t < num_threads ? sd : ed;
        oracle_merge_path_search(diag, row_end_offsets, num_rows, num_nonzeros,
                                 &coords[2 * t], &coords[2 * t + 1]);
    }
}

/* ------------------------------------------------------------------------------------
 * OmpMergeCsrmv -- cpu_spmv.cpp:292-353, for value type T.
 *   - each thread owns an equal share of the rows+nnz merge path (:311-321)
 *   - whole rows: y[row] = running sum, accumulated left to right in T (:324-333)
 *   - trailing partial row -> (row_carry_out, value_carry_out) (:336-344)
 *   - serial fix-up for tid < p-1, guarded by row < num_rows (:348-352)
 * The reference keeps the carries in 256-entry stack arrays (:302-303); here they are
 * heap arrays so any thread count works, the arithmetic is unchanged.
 * ---------------------------------------------------------------------------------- */
#define DEFINE_MERGE_CSRMV(NAME, T)                                                              \
    void NAME(int num_threads, int num_rows, int num_nonzeros, const int* row_end_offsets,       \
              const int* column_indices, const T* values, const T* x, T* y)                      \
    {                                                                                            \
        if (num_threads < 1) num_threads = 1;                                                    \
        int* carry_row = (int*)malloc(sizeof(int) * (size_t)num_threads);                        \
        T* carry_val = (T*)malloc(sizeof(T) * (size_t)num_threads);                              \
        _Pragma("omp parallel for schedule(static) num_threads(num_threads)")                    \
        for (int tid = 0; tid < num_threads; ++tid) {                                            \
            int sd, ed, cx, cy, ex, ey;                                                          \
            thread_diagonals(tid, num_threads, num_rows, num_nonzeros, &sd, &ed);                \
            oracle_merge_path_search(sd, row_end_offsets, num_rows, num_nonzeros, &cx, &cy);     \
            oracle_merge_path_search(ed, row_end_offsets, num_rows, num_nonzeros, &ex, &ey);     \
            for (; cx < ex; ++cx) {                                                              \
                T acc = (T)0.0;                                                                  \
                for (; cy < row_end_offsets[cx]; ++cy) acc += values[cy] * x[column_indices[cy]]; \
                y[cx] = acc;                                                                     \
            }                                                                                    \
            T tail = (T)0.0;                                                                     \
            for (; cy < ey; ++cy) tail += values[cy] * x[column_indices[cy]];                    \
            carry_row[tid] = ex;                                                                 \
            carry_val[tid] = tail;                                                               \
        }                                                                                        \
        for (int tid = 0; tid < num_threads - 1; ++tid)                                          \
            if (carry_row[tid] < num_rows) y[carry_row[tid]] += carry_val[tid];                  \
        free(carry_row);                                                                         \
        free(carry_val);                                                                         \
    }

DEFINE_MERGE_CSRMV(oracle_merge_csrmv_f32, float)
DEFINE_MERGE_CSRMV(oracle_merge_csrmv_f64, double)

/* ------------------------------------------------------------------------------------
 * SpmvGold -- cpu_spmv.cpp:257-277 / gpu_spmv.cu:72-92: y = alpha*A*x + beta*y_in,
 * sequential, partial starts at beta*y_in[row] and adds alpha*v*x[col] term by term.
 * ---------------------------------------------------------------------------------- */
#define DEFINE_GOLD(NAME, T)                                                                     \
    void NAME(int num_rows, const int* row_offsets, const int* column_indices, const T* values,  \
              const T* x, const T* y_in, T* y_out, T alpha, T beta)                              \
    {                                                                                            \
        for (int r = 0; r < num_rows; ++r) {                                                     \
            T p = beta * y_in[r];                                                                \
            for (int k = row_offsets[r]; k < row_offsets[r + 1]; ++k)                            \
                p += alpha * values[k] * x[column_indices[k]];                                   \
            y_out[r] = p;                                                                        \
        }                                                                                        \
    }

DEFINE_GOLD(oracle_spmv_gold_f32, float)
DEFINE_GOLD(oracle_spmv_gold_f64, double)

/* ------------------------------------------------------------------------------------
 * CompareResults -- utils.h:692-713 (float) and :720-742 (double, which first narrows
 * both operands to float, :728-729).  Distance = |bits(a) - bits(b)| on the fp32
 * patterns; FAIL (1) iff sqrt(distance) > len.  Returns the index of the first failing
 * element + 1, or 0 for PASS (the reference returns just 0/1).
 * ---------------------------------------------------------------------------------- */
static int fails_ulps(float a, float b, long long len)
{
    int ia, ib;
    memcpy(&ia, &a, sizeof ia);
    memcpy(&ib, &b, sizeof ib);
    int d = abs(ia - ib);
    float s = sqrtf((float)d);
    return s > (float)len;
}

long long oracle_compare_results_f32(const float* computed, const float* reference, long long len)
{
    for (long long i = 0; i < len; ++i)
        if (fails_ulps(computed[i], reference[i], len)) return i + 1;
    return 0;
}

long long oracle_compare_results_f64(const double* computed, const double* reference, long long len)
{
    for (long long i = 0; i < len; ++i)
        if (fails_ulps((float)computed[i], (float)reference[i], len)) return i + 1;
    return 0;
}

/* ------------------------------------------------------------------------------------
 * Timing in the reference's protocol -- TestOmpMergeCsrmv, cpu_spmv.cpp:362-406:
 * poison y, one checked call, three warm calls, then `iterations` timed calls; returns
 * average milliseconds.  Wall clock = omp_get_wtime like utils.h:533-553 (-DCUB_MKL).
 * ---------------------------------------------------------------------------------- */
static double wall_seconds(void)
{
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

#define DEFINE_TIMED(NAME, CSRMV, T)                                                             \
    double NAME(int num_threads, int num_rows, int num_nonzeros, const int* row_end_offsets,     \
                const int* column_indices, const T* values, const T* x, T* y, int iterations)    \
    {                                                                                            \
        memset(y, -1, sizeof(T) * (size_t)num_rows);                                             \
        for (int w = 0; w < 4; ++w)                                                              \
            CSRMV(num_threads, num_rows, num_nonzeros, row_end_offsets, column_indices, values,  \
                  x, y);                                                                         \
        double t0 = wall_seconds();                                                              \
        for (int it = 0; it < iterations; ++it)                                                  \
            CSRMV(num_threads, num_rows, num_nonzeros, row_end_offsets, column_indices, values,  \
                  x, y);                                                                         \
        double t1 = wall_seconds();                                                              \
        return (t1 - t0) * 1000.0 / (iterations > 0 ? iterations : 1);                           \
    }

DEFINE_TIMED(oracle_time_merge_csrmv_f32, oracle_merge_csrmv_f32, float)
DEFINE_TIMED(oracle_time_merge_csrmv_f64, oracle_merge_csrmv_f64, double)

int oracle_num_procs(void)
{
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
