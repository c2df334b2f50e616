/* Stand-in for libnuma's <numa.h> (sparse_matrix.h:49 under -DCUB_MKL).  TEST
 * INFRASTRUCTURE ONLY.  Reports "NUMA unavailable" so the reference takes its
 * mkl_malloc branch (sparse_matrix.h:657,693-698; cpu_spmv.cpp:628-634). */
#ifndef MSPMV_ORACLE_SHIM_NUMA_H
#define MSPMV_ORACLE_SHIM_NUMA_H
#include <stdlib.h>
#ifdef __cplusplus
extern "C" {
#endif
static inline int   numa_available(void) { return -1; }
static inline void  numa_set_strict(int s) { (void)s; }
static inline int   numa_num_task_nodes(void) { return 1; }
static inline void* numa_alloc_onnode(size_t bytes, int node) { (void)node; return malloc(bytes ? bytes : 1); }
static inline void  numa_free(void* p, size_t bytes) { (void)bytes; free(p); }
#ifdef __cplusplus
}
#endif
#endif
