/* Stand-in for Intel MKL's <mkl.h>, which the reference's cpu_spmv.cpp:61 includes
 * unconditionally but which is not installed in this image.  TEST INFRASTRUCTURE ONLY:
 * it exists so that oracle/_ref can compile the reference's own OmpMergeCsrmv /
 * MergePathSearch / SpmvGold from /root/reference unmodified.  The two csrgemv entry
 * points are only the "MKL CsrMV" comparator column (cpu_spmv.cpp:426,442); here they
 * are a plain sequential zero-based CSR loop. */
#ifndef MSPMV_ORACLE_SHIM_MKL_H
#define MSPMV_ORACLE_SHIM_MKL_H
#include <stdlib.h>
#ifdef __cplusplus
extern "C" {
#endif
static inline void* mkl_malloc(size_t bytes, int align) {
    void* p = NULL;
    if (align < (int)sizeof(void*)) align = (int)sizeof(void*);
    if (posix_memalign(&p, (size_t)align, bytes ? bytes : 1)) return NULL;
    return p;
}
static inline void mkl_free(void* p) { free(p); }
static inline void mkl_cspblas_scsrgemv(const char* trans, const int* m, const float* a,
                                        const int* ia, const int* ja, const float* x, float* y) {
    (void)trans;
    for (int r = 0; r < *m; ++r) {
        float s = 0.f;
        for (int k = ia[r]; k < ia[r + 1]; ++k) s += a[k] * x[ja[k]];
        y[r] = s;
    }
}
static inline void mkl_cspblas_dcsrgemv(const char* trans, const int* m, const double* a,
                                        const int* ia, const int* ja, const double* x, double* y) {
    (void)trans;
    for (int r = 0; r < *m; ++r) {
        double s = 0.;
        for (int k = ia[r]; k < ia[r + 1]; ++k) s += a[k] * x[ja[k]];
        y[r] = s;
    }
}
#ifdef __cplusplus
}
#endif
#endif
