// carry_exchange.cuh -- multi-GPU carry exchange over NVLink peer memory, fused with the fold.
//
// After the sharded CsrMV every rank holds ONE carry (the partial sum of the row that continues on
// the next rank, cpu_spmv.cpp:336-344).  The NCCL path moves the p carries with an all_gather and
// folds them with a second kernel (mspmv_apply_carries_*).  This kernel does both in one launch,
// writing straight into the other GPUs' memory:
//   push   thread g stores my carry into rank g's exchange buffer (slot [parity][my rank]) and then
//          releases a flag (= the epoch number) next to it -- 16 bytes per peer over NVLink;
//   fold   thread g polls the flag rank g set in MY buffer (acquire, system scope); when all p have
//          arrived, thread 0 folds carries 0..p-2 into the rows this rank owns in shard order, with
//          the row < num_rows guard -- the serial fix-up of cpu_spmv.cpp:348-352.
// The buffer of every rank is symmetric memory of 4*p 64-bit words: values[2][p], flags[2][p].
// Slots are double-buffered by epoch parity: a rank can be at most one exchange ahead of a peer (it
// cannot finish epoch e+1 before that peer has pushed e+1, which the peer does only after folding e),
// so the slot of epoch e is never overwritten before it has been read.  The epoch counter lives in
// device memory, so a CUDA graph that contains this kernel can be replayed.
// No rank waits for a peer's *kernel to start* before pushing, so there is no co-scheduling
// requirement and no deadlock as long as every rank eventually launches the kernel.
#pragma once

#include <stdio.h>
#include <string.h>

#include "merge_common.cuh"

namespace mspmv {

constexpr int kExchangePhasePush = 1, kExchangePhaseFold = 2;
constexpr uint64_t kExchangeTimeoutNs = 10ull * 1000 * 1000 * 1000;  // 10 s

template <typename T>
__global__ void carry_exchange_kernel(T* __restrict__ y_local, int carry_index, int y_row_begin, int y_rows,
                                      int num_rows_global, const int* __restrict__ carry_rows,
                                      uint64_t* const* __restrict__ peer_bufs, int rank, int world,
                                      unsigned long long* epoch_ctr, int phase)
{
    __shared__ unsigned long long s_epoch;
    if (threadIdx.x == 0) s_epoch = *epoch_ctr + 1ull;
    __syncthreads();
    const uint64_t epoch = s_epoch;
    const int par = (int)(epoch & 1ull);
    uint64_t* mine = peer_bufs[rank];
    if (phase & kExchangePhasePush) {
        const T v = y_local[carry_index];
        uint64_t bits = 0;
        memcpy(&bits, &v, sizeof(T));
        for (int g = threadIdx.x; g < world; g += blockDim.x) {
            uint64_t* dst = peer_bufs[g];
            st_relaxed_sys_u64(dst + par * world + rank, bits);
            st_release_sys_u64(dst + 2 * world + par * world + rank, epoch);  // orders the value before the flag
        }
    }
    if (phase & kExchangePhaseFold) {
        // A peer that never launches its kernel (an exception on one rank, mismatched call counts) must not
        // hang this GPU forever: after kExchangeTimeoutNs the kernel reports and traps, and the host sees a
        // launch failure at its next synchronisation instead of a silent hang.
        const uint64_t t_start = global_timer_ns();
        for (int g = threadIdx.x; g < world; g += blockDim.x)
            while (ld_acquire_sys_u64(mine + 2 * world + par * world + g) != epoch) {
                if (global_timer_ns() - t_start > kExchangeTimeoutNs) {
                    printf("mspmv carry exchange: rank %d timed out waiting for rank %d (epoch %llu)\n", rank, g,
                           (unsigned long long)epoch);
                    trap_kernel();
                }
            }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int g = 0; g < world - 1; ++g) {
                const int row = carry_rows[g];
                if (row < num_rows_global && row >= y_row_begin && row < y_row_begin + y_rows) {
                    const uint64_t bits = ld_acquire_sys_u64(mine + par * world + g);
                    T v;
                    memcpy(&v, &bits, sizeof(T));
                    y_local[row - y_row_begin] += v;
                }
            }
            *epoch_ctr = epoch;
        }
    }
}

}  // namespace mspmv
