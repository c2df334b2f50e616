// ptx_emu.cuh -- host stand-ins for the wrappers of merge-spmv_b200/csrc/ptx_sm100.cuh, used only by
// the SIMT interpreter in tests/emu/ (see simt_emu.hpp).  They keep the *contracts* of the
// instructions so that misuse is caught on the CPU:
//   * cp.async.bulk: 16-byte aligned source and destination, size a multiple of 16; the destination
//     is poisoned at issue and the data lands only when the mbarrier phase completes, so a kernel
//     that reads staged data without waiting on the barrier fails its parity test;
//   * mbarrier: arrival count + transaction bytes, phase parity, waits yield to the other threads.
#pragma once

#include <stdint.h>

#include "simt_emu.hpp"

// dynamic shared memory: one static buffer of the largest size a block can ask for
#define MSPMV_DYNAMIC_SHARED(name) alignas(1024) static unsigned char name[232448]

namespace mspmv {

struct EmuMbar {  // lives in the kernel's 8-byte mbarrier word
    int64_t tx : 36;
    uint64_t pending : 13;
    uint64_t expected : 13;
    uint64_t phase : 1;
};
static_assert(sizeof(EmuMbar) == 8, "mbarrier word");

struct EmuBulkCopy {
    void* dst;
    const void* src;
    uint32_t bytes;
    uint64_t* bar;
};
// in-flight copies: a plain array handled only inside EMU_INTERNAL functions, so that the interpreter's own
// bookkeeping (touched by whichever CUDA thread issues a copy or completes a phase) is invisible to TSan
struct EmuInflight {
    EmuBulkCopy item[4096];
    size_t n = 0;
};
EMU_INTERNAL inline EmuInflight& emu_inflight()
{
    static EmuInflight v;
    return v;
}
// the data movement of a bulk copy, visible to the sanitizers as ordinary writes of the thread that
// completes the phase (ordered before the waiters by the release / acquire on the barrier)
inline void emu_land_copy(void* dst, const void* src, uint32_t bytes) { std::memcpy(dst, src, bytes); }
inline void emu_poison_copy(void* dst, uint32_t bytes) { std::memset(dst, 0xCD, bytes); }

EMU_INTERNAL inline EmuMbar* emu_bar(uint64_t* bar) { return reinterpret_cast<EmuMbar*>(bar); }

EMU_INTERNAL inline void emu_mbar_check_complete(uint64_t* bar)
{
    EmuMbar* b = emu_bar(bar);
    if (b->pending != 0 || b->tx != 0) return;
    // the phase completes: the bulk copies that signal this barrier become visible now
    EmuInflight& fl = emu_inflight();
    for (size_t i = 0; i < fl.n;) {
        if (fl.item[i].bar == bar) {
            emu_land_copy(fl.item[i].dst, fl.item[i].src, fl.item[i].bytes);
            fl.item[i] = fl.item[fl.n - 1];
            --fl.n;
        } else {
            ++i;
        }
    }
    b->phase ^= 1;
    b->pending = b->expected;
    emu::S().progress = true;
    EMU_HB_RELEASE(bar);  // the landed bytes and everything the arriving threads did before
}

EMU_INTERNAL inline void mbar_init(uint64_t* bar, int count)
{
    EmuMbar* b = emu_bar(bar);
    b->tx = 0;
    b->pending = (uint64_t)count;
    b->expected = (uint64_t)count;
    b->phase = 0;
}
inline void fence_mbar_init() {}
EMU_INTERNAL inline void mbar_arrive(uint64_t* bar)
{
    EmuMbar* b = emu_bar(bar);
    if (b->pending == 0) emu::die("mbarrier: more arrivals than expected");
    EMU_HB_RELEASE(bar);
    b->pending -= 1;
    emu_mbar_check_complete(bar);
}
EMU_INTERNAL inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    EmuMbar* b = emu_bar(bar);
    b->tx += (int64_t)bytes;
    mbar_arrive(bar);
}
EMU_INTERNAL inline bool mbar_test_wait(uint64_t* bar, uint32_t parity) { return emu_bar(bar)->phase != (parity & 1u); }
EMU_INTERNAL inline void mbar_wait(uint64_t* bar, uint32_t parity)
{
    emu::S().progress = true;
    while (!mbar_test_wait(bar, parity)) emu::yield();
    EMU_HB_ACQUIRE(bar);
}
inline uint64_t l2_policy_evict_first() { return 0; }
inline uint64_t l2_policy_evict_last() { return 0; }

EMU_INTERNAL inline void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t)
{
    if (((uintptr_t)dst_smem & 15) || ((uintptr_t)src_gmem & 15) || (bytes & 15) || bytes == 0)
        emu::die("cp.async.bulk: addresses must be 16-byte aligned and the size a non-zero multiple of 16");
    emu_poison_copy(dst_smem, bytes);  // not visible before the barrier phase completes
    EmuInflight& fl = emu_inflight();
    if (fl.n == 4096) emu::die("too many bulk copies in flight");
    fl.item[fl.n++] = EmuBulkCopy{dst_smem, src_gmem, bytes, bar};
    emu_bar(bar)->tx -= (int64_t)bytes;  // complete_tx
    emu_mbar_check_complete(bar);
}
inline uint64_t global_timer_ns()
{
    static uint64_t t = 0;
    return t += 1000;  // a microsecond per look at the clock: the 10 s limit is ~10 million polls
}
inline void trap_kernel() { emu::die("kernel trapped"); }
inline void st_relaxed_sys_u64(uint64_t* p, uint64_t v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
inline void st_release_sys_u64(uint64_t* p, uint64_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
EMU_INTERNAL inline uint64_t ld_acquire_sys_u64(const uint64_t* p)
{
    // a thread that reads the same address and gets the same value again is spinning: give the other
    // threads a turn without counting it as progress (so a flag nobody sets ends as a reported deadlock)
    static const uint64_t* last_addr[emu::kMaxThreads];
    static uint64_t last_val[emu::kMaxThreads];
    const int t = emu::S().cur;
    const uint64_t v = __atomic_load_n(p, __ATOMIC_ACQUIRE);
    if (last_addr[t] == p && last_val[t] == v) {
        emu::yield();
        return __atomic_load_n(p, __ATOMIC_ACQUIRE);
    }
    last_addr[t] = p;
    last_val[t] = v;
    return v;
}
// cp.async (LDGSTS): the copy is performed at issue; the thread's arrival on the barrier (which the
// hardware defers until its copies have landed) follows it in program order
inline void fence_proxy_async() {}
inline void named_bar_sync(int id, int threads) { emu::block_barrier(id, threads); }

#ifndef MSPMV_GATHER_FLAVOUR
#define MSPMV_GATHER_FLAVOUR 5
#endif
template <typename T>
inline T ld_gather(const T* p, uint64_t)
{
    return *p;
}
template <typename T>
inline T ld_gather_l1(const T* p, uint64_t)
{
    return *p;
}
}  // namespace mspmv
