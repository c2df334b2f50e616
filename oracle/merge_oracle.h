/* oracle/merge_oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE, see merge_oracle.c). */
#ifndef MSPMV_MERGE_ORACLE_H
#define MSPMV_MERGE_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

void oracle_merge_path_search(int diagonal, const int* row_end_offsets, int a_len, int b_len,
                              int* out_x, int* out_y);
void oracle_merge_thread_coords(int num_threads, int num_rows, int num_nonzeros,
                                const int* row_end_offsets, int* coords /* 2*(num_threads+1) */);

void oracle_merge_csrmv_f32(int num_threads, int num_rows, int num_nonzeros,
                            const int* row_end_offsets, const int* column_indices,
                            const float* values, const float* x, float* y);
void oracle_merge_csrmv_f64(int num_threads, int num_rows, int num_nonzeros,
                            const int* row_end_offsets, const int* column_indices,
                            const double* values, const double* x, double* y);

void oracle_spmv_gold_f32(int num_rows, const int* row_offsets, const int* column_indices,
                          const float* values, const float* x, const float* y_in, float* y_out,
                          float alpha, float beta);
void oracle_spmv_gold_f64(int num_rows, const int* row_offsets, const int* column_indices,
                          const double* values, const double* x, const double* y_in, double* y_out,
                          double alpha, double beta);

long long oracle_compare_results_f32(const float* computed, const float* reference, long long len);
long long oracle_compare_results_f64(const double* computed, const double* reference, long long len);

double oracle_time_merge_csrmv_f32(int num_threads, int num_rows, int num_nonzeros,
                                   const int* row_end_offsets, const int* column_indices,
                                   const float* values, const float* x, float* y, int iterations);
double oracle_time_merge_csrmv_f64(int num_threads, int num_rows, int num_nonzeros,
                                   const int* row_end_offsets, const int* column_indices,
                                   const double* values, const double* x, double* y, int iterations);
int oracle_num_procs(void);

#ifdef __cplusplus
}
#endif
#endif
