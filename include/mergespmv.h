/*
 * mergespmv.h -- C ABI of libmergespmv.so: Blackwell-native (sm_100a) merge-based CSR SpMV.
 *
 * This is the drop-in boundary for the one hot path of dumerrill/merge-spmv:
 *
 *     cub::DeviceSpmv::CsrMV<ValueT>(d_temp_storage, temp_storage_bytes, d_values,
 *                                    d_row_offsets, d_column_indices, d_vector_x, d_vector_y,
 *                                    num_rows, num_cols, num_nonzeros, stream, debug_synchronous)
 *                                                    -- cub/device/device_spmv.cuh:129-164
 *
 * Plain pointers and sizes only; no C++/torch types.  All citations are file:line in the
 * reference tree.  Every function returns a cudaError_t value as int (0 == cudaSuccess), the
 * reference's own error convention (dispatch_spmv_orig.cuh:563-749).
 *
 * There is NO CPU fallback anywhere behind this interface: without a CUDA device the compute
 * entry points return the CUDA error they hit.
 */
#ifndef MERGESPMV_H
#define MERGESPMV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSPMV_VERSION_MAJOR 0
#define MSPMV_VERSION_MINOR 2

/* cudaStream_t is passed as void* so that this header needs no CUDA headers. */
typedef void* mspmv_stream_t;

/* -------------------------------------------------------------------------------------------
 * 1. y = A*x -- replaces cub::DeviceSpmv::CsrMV<float|double> (device_spmv.cuh:129-164) as
 *    called by TestGpuMergeCsrmv (gpu_spmv.cu:390-395 size query, :404-409 warm-up, :424-429
 *    timed loop).  Same argument order and meaning:
 *      - d_temp_storage == NULL: write the required byte count to *temp_storage_bytes, launch
 *        nothing, return 0 (dispatch_spmv_orig.cuh:651-655).
 *      - otherwise *temp_storage_bytes must be >= that count, else cudaErrorInvalidValue
 *        (util_device.cuh:90-93).  The blob needs no initialisation and is reusable.
 *      - d_row_offsets has num_rows+1 zero-based entries, last == num_nonzeros.
 *      - alpha = 1, beta = 0 (device_spmv.cuh:155-156): every y[r] is overwritten, empty rows
 *        get 0.  Inputs are not modified.  Asynchronous on `stream` unless debug_synchronous.
 *      - debug_synchronous != 0 logs each launch configuration to stdout and synchronises the
 *        stream after every kernel (dispatch_spmv_orig.cuh:579,590,685,698,702,718,724,739).
 *    Differences, all deliberate:
 *      - never reads row_offsets[num_rows+1] and never adds a carry to y[num_rows]
 *        (the reference's out-of-bounds read/write, App. A items 5-6);
 *      - num_cols == 1 takes the general path and sums duplicates (the reference's
 *        DeviceSpmv1ColKernel, dispatch_spmv_orig.cuh:68-93, keeps only the first nonzero);
 *      - carry fix-up is deterministic for fp32 too (the reference uses atomics there,
 *        agent_segment_fixup.cuh:226-260).
 *    num_rows + num_nonzeros must be < 2^31 like the reference (dispatch_spmv_orig.cuh:608).
 * ----------------------------------------------------------------------------------------- */
int mspmv_csrmv_f32(void* d_temp_storage, size_t* temp_storage_bytes, const float* d_values,
                    const int* d_row_offsets, const int* d_column_indices, const float* d_vector_x,
                    float* d_vector_y, int num_rows, int num_cols, int num_nonzeros,
                    mspmv_stream_t stream, int debug_synchronous);
int mspmv_csrmv_f64(void* d_temp_storage, size_t* temp_storage_bytes, const double* d_values,
                    const int* d_row_offsets, const int* d_column_indices, const double* d_vector_x,
                    double* d_vector_y, int num_rows, int num_cols, int num_nonzeros,
                    mspmv_stream_t stream, int debug_synchronous);

/* y = alpha*A*x + beta*y -- the epilogue the reference's CLI (--alpha/--beta, gpu_spmv.cu:721-722)
 * and SpmvGold (gpu_spmv.cu:72-92) accept but its merge kernels leave disabled
 * (dispatch_spmv_orig.cuh:789-838; agent_spmv_orig.cuh:386-396).  beta == 0 never reads y. */
int mspmv_csrmv_axpby_f32(void* d_temp_storage, size_t* temp_storage_bytes, const float* d_values,
                          const int* d_row_offsets, const int* d_column_indices,
                          const float* d_vector_x, float* d_vector_y, int num_rows, int num_cols,
                          int num_nonzeros, float alpha, float beta, mspmv_stream_t stream,
                          int debug_synchronous);
int mspmv_csrmv_axpby_f64(void* d_temp_storage, size_t* temp_storage_bytes, const double* d_values,
                          const int* d_row_offsets, const int* d_column_indices,
                          const double* d_vector_x, double* d_vector_y, int num_rows, int num_cols,
                          int num_nonzeros, double alpha, double beta, mspmv_stream_t stream,
                          int debug_synchronous);

/* -------------------------------------------------------------------------------------------
 * 2. Merge-path coordinates (bit-exact against the reference's MergePathSearch,
 *    cpu_spmv.cpp:223-245 == cub/thread/thread_search.cuh:53-84; what DeviceSpmvSearchKernel,
 *    dispatch_spmv_orig.cuh:104-143, computes per tile).
 * ----------------------------------------------------------------------------------------- */

/* Device: for each of n diagonals (device array) write (x = row index, y = nonzero index) to
 * d_coords[2*i], d_coords[2*i+1].  List A is row_end_offsets = d_row_offsets + 1. */
int mspmv_merge_path_search(const int* d_row_offsets, int num_rows, int num_nonzeros,
                            const int* d_diagonals, int n, int* d_coords, mspmv_stream_t stream);

/* Device: the start coordinate of every threadblock's diagonal swath for this problem shape,
 * exactly as mspmv_csrmv_* computes them in-kernel, plus the final end coordinate.
 * Size query: d_coords == NULL writes the swath count to *num_swaths.  Otherwise writes
 * 2*(*num_swaths+1) ints to d_coords (device memory). */
int mspmv_csrmv_swath_coords(const int* d_row_offsets, int num_rows, int num_nonzeros,
                             int value_bytes, int* num_swaths, int* d_coords,
                             mspmv_stream_t stream);

/* Host: same search on host memory (used to cut the matrix into per-GPU shards). */
void mspmv_host_merge_path_search(const int* row_offsets, int num_rows, int num_nonzeros,
                                  int64_t diagonal, int* out_x, int* out_y);

/* -------------------------------------------------------------------------------------------
 * 3. Multi-GPU sharding by the same merge decomposition (new surface: the reference is
 *    single-GPU; README.md:5 and the paper's section III.A only assert that the
 *    decomposition partitions hierarchically).  One process per GPU; shard g of p owns
 *    diagonals [g*ceil((rows+nnz)/p), ...) exactly like thread g of OmpMergeCsrmv
 *    (cpu_spmv.cpp:311-321).
 * ----------------------------------------------------------------------------------------- */

/* Host: cut points.  coords[2*g], coords[2*g+1] = (row, nonzero) where shard g starts,
 * g = 0..num_shards (last entry = (num_rows, num_nonzeros)). */
void mspmv_shard_partition(const int* row_offsets, int num_rows, int num_nonzeros, int num_shards,
                           int* coords);

/* Host: local CSR row offsets of one shard.  The shard holds rows [x0, x1) that END in it plus
 * one trailing partial row (row x1, possibly empty): local_rows = x1 - x0 + 1 entries of y, the
 * last of which is the carry-out (cpu_spmv.cpp:336-344).  Writes local_rows+1 offsets rebased
 * by -y0 to local_row_offsets. */
void mspmv_shard_row_offsets(const int* row_offsets, int x0, int y0, int x1, int y1,
                             int* local_row_offsets);

/* Device: after all shards' (row, carry) records have been all-gathered (one NCCL call; a
 * record is the int32 row followed by the value), fold the carries of shards 0..p-2 into the
 * local y slice in shard order, skipping rows >= num_rows_global -- the serial fix-up of
 * cpu_spmv.cpp:348-352.  d_y_local[i] is global row y_row_begin + i, y_rows entries. */
int mspmv_apply_carries_f32(float* d_y_local, int y_row_begin, int y_rows, int num_rows_global,
                            const int* d_carry_rows, const float* d_carry_vals, int num_shards,
                            mspmv_stream_t stream);
int mspmv_apply_carries_f64(double* d_y_local, int y_row_begin, int y_rows, int num_rows_global,
                            const int* d_carry_rows, const double* d_carry_vals, int num_shards,
                            mspmv_stream_t stream);

/* The same exchange step over NVLink peer memory, fused with the fold, in ONE launch per rank and
 * without NCCL (csrc/carry_exchange.cuh): every rank stores its carry -- the last of the
 * local_rows entries of d_y_local -- directly into each peer's exchange buffer and releases a flag;
 * it then waits for the p flags in its own buffer and folds carries 0..p-2 exactly like
 * mspmv_apply_carries_*.  d_peer_bufs is a DEVICE array of num_shards pointers, entry g = rank g's
 * exchange buffer as mapped into this process (symmetric memory: cudaIpc / VMM peer mapping, e.g.
 * torch.distributed._symmetric_memory), each mspmv_exchange_buffer_bytes(num_shards) bytes and
 * zero-filled before the first call; d_epoch is a zero-initialised device counter private to the
 * rank.  Every rank must make the same sequence of calls; safe to capture into a CUDA graph. */
size_t mspmv_exchange_buffer_bytes(int num_shards);
int mspmv_exchange_carries_f32(float* d_y_local, int local_rows, int y_row_begin, int y_rows,
                               int num_rows_global, const int* d_carry_rows, void* const* d_peer_bufs,
                               int rank, int num_shards, unsigned long long* d_epoch, mspmv_stream_t stream);
int mspmv_exchange_carries_f64(double* d_y_local, int local_rows, int y_row_begin, int y_rows,
                               int num_rows_global, const int* d_carry_rows, void* const* d_peer_bufs,
                               int rank, int num_shards, unsigned long long* d_epoch, mspmv_stream_t stream);

/* -------------------------------------------------------------------------------------------
 * 4. Host-buffer operator ("session"): the gpu_spmv driver's own protocol -- upload the CSR
 *    once (gpu_spmv.cu:542-556), then apply it repeatedly (:421-432) -- behind one handle, so
 *    that callers with host memory do not touch CUDA.  create() uploads A (setup); apply()
 *    copies x host->device, runs mspmv_csrmv_*, copies y device->host, and returns when y is
 *    valid.  value_bytes is 4 (float) or 8 (double).
 * ----------------------------------------------------------------------------------------- */
typedef struct mspmv_session mspmv_session;

int mspmv_session_create(mspmv_session** out, int device, int value_bytes, int num_rows,
                         int num_cols, int num_nonzeros, const int* row_offsets,
                         const int* column_indices, const void* values);
int mspmv_session_apply(mspmv_session* s, const void* x_host, void* y_host);
/* Pipelined stream of n right-hand sides (xs: n*num_cols values, ys: n*num_rows values, both
 * ideally pinned): copies, kernels and read-backs of consecutive vectors overlap. */
int mspmv_session_apply_many(mspmv_session* s, int n, const void* xs_host, void* ys_host);
void mspmv_session_destroy(mspmv_session* s);

/* Pinned host memory helpers for the session's callers. */
int mspmv_host_alloc(void** out, size_t bytes);
int mspmv_host_free(void* p);

/* -------------------------------------------------------------------------------------------
 * 5. Multi-GPU host-buffer operator: sections 3 and 4 combined behind one handle -- ONE process,
 *    num_shards GPUs with peer access (NVLink / NVSwitch), no NCCL, no Python (csrc/mg_session.cu).
 *    New surface: the reference is single-GPU.  create() cuts the CSR into merge-path shards
 *    (mspmv_shard_partition, cpu_spmv.cpp:311-321 with p = num_shards) and uploads shard g to
 *    device_ids[g] (NULL: devices 0..num_shards-1; a device may hold several shards).  apply():
 *    x host -> shard 0's device (one PCIe crossing) -> peer copies to the other devices; every
 *    device runs mspmv_csrmv_* on its shard plus ONE exchange kernel (mspmv_exchange_carries_*);
 *    every device copies the rows it owns into their place in y_host.  apply_many() pipelines
 *    n right-hand sides over three slots and three streams per device like
 *    mspmv_session_apply_many.  Host buffers should be pinned (mspmv_host_alloc).
 * ----------------------------------------------------------------------------------------- */
typedef struct mspmv_mg_session mspmv_mg_session;

int mspmv_mg_session_create(mspmv_mg_session** out, int num_shards, const int* device_ids,
                            int value_bytes, int num_rows, int num_cols, int num_nonzeros,
                            const int* row_offsets, const int* column_indices, const void* values);
int mspmv_mg_session_apply(mspmv_mg_session* s, const void* x_host, void* y_host);
int mspmv_mg_session_apply_many(mspmv_mg_session* s, int n, const void* xs_host, void* ys_host);
/* Device-side time of one sharded product (CsrMV + carry exchange on every device, x of the last
 * apply already resident): `iterations` back-to-back products between two events per device,
 * the maximum over devices divided by iterations is written to *ms_per_step. */
int mspmv_mg_session_time_device(mspmv_mg_session* s, int iterations, float* ms_per_step);
/* Shard g's cut: out[0..3] = (x0, y0, x1, y1) -- rows [x0, x1) end in it, nonzeros [y0, y1) --
 * and out[4] = its device. */
int mspmv_mg_session_shard(const mspmv_mg_session* s, int shard, int* out);
void mspmv_mg_session_destroy(mspmv_mg_session* s);

/* -------------------------------------------------------------------------------------------
 * 6. Introspection.
 * ----------------------------------------------------------------------------------------- */
int mspmv_version(void); /* major*100 + minor */
/* PTX/SASS version the kernels were compiled for, in cub::PtxVersion units (util_device.cuh:118-160:
 * major*100 + minor*10, e.g. 1000 for sm_100a); written to *ptx_version. */
int mspmv_ptx_version(int* ptx_version);
/* Kernel launches issued by this library since load (all entry points). */
uint64_t mspmv_launch_count(void);
/* Launch geometry mspmv_csrmv_* uses for a shape (6 ints): out[0] = threadblocks (each owns a
 * contiguous run of tiles), out[1] = threads per block, out[2] = merge items per tile,
 * out[3] = dynamic smem bytes per block, out[4] = kernels per call (1), out[5] = tiles
 * (equal-length diagonal swaths of the merge path). */
int mspmv_csrmv_config(int value_bytes, int num_rows, int num_nonzeros, int* out);
const char* mspmv_error_string(int err);
/* Kept for callers that name an engine: "auto" and "pipe" (the one engine, csrc/spmv_pipe.cuh)
 * return 0, anything else 1. */
int mspmv_set_engine(const char* name);
/* Tuning / test options for subsequent calls in this process.  Returns 0, or 1 for an unknown name.
 * value -1 returns an option to its environment value / default.
 *   "pipe_search"         1 (default): the producer warp of every threadblock finds its tile
 *                         coordinates itself -- ONE launch per CsrMV (plus a 4-byte memset node).
 *                         0: DeviceSpmvSearchKernel's analogue runs first (dispatch_spmv_orig.cuh:689),
 *                         two launches.  Same tiles, same bits.
 *   "pipe_config"         0 (default): kernel shape by average row length -- B for
 *                         (rows + nnz) / rows <= 16, else A (the analogue of the reference's
 *                         per-type tile policies, dispatch_spmv_orig.cuh:393-441).  1: A.  2: B.
 *   "pipe_smem_kb"        shared-memory budget per SM that sizes the grid (default A 160, B 192; the
 *                         rest of the 228 KB stays L1 for the x gathers' misses in flight).
 *   "pipe_blocks_per_sm"  cap on resident threadblocks per SM (0 = what the budget allows).
 *   "pipe_export_coords"  1: the producer warps also write the coordinates they found to the start
 *                         of the temp blob (tiles + 1 int2; the bit-exact MergePathSearch check). */
int mspmv_set_option(const char* name, int value);

#ifdef __cplusplus
}
#endif
#endif /* MERGESPMV_H */
