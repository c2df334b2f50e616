// simt_emu.hpp -- a small host-side SIMT interpreter, TEST INFRASTRUCTURE ONLY.
//
// Purpose: run the *logic* of the sm_100a kernels (bitmap construction, popcount prefix, merge walk,
// segmented scans, carry fix-up, staging-range arithmetic) on a box without a GPU, so that
// `pytest -m "not gpu"` can check it against the oracle.  It is compiled only by
// tests/test_kernel_emu.py into tests/emu/_build/, is never linked into libmergespmv.so and is not
// importable from the merge_spmv_b200 package: the product has no CPU path.
//
// Model: one CUDA thread = one ucontext coroutine; the threads of a block are resumed round-robin by
// a scheduler in a single OS thread and give up control only inside a synchronising primitive
// (__syncthreads, named barriers, warp collectives, mbarrier waits).  Blocks run one after another,
// so `__shared__` variables are plain statics.  A pass over all threads in which nobody made
// progress is a deadlock (e.g. a barrier not reached by every thread) and aborts with a message.
// Warp collectives require the full mask -- all the kernels use.
#pragma once

#include <cuda_runtime.h>  // vector types; __device__ / __global__ expand to nothing for a host compiler
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)

// EMU_TSAN: the same interpreter as a data-race detector.  Every CUDA thread is registered with
// ThreadSanitizer as a fiber; barriers, warp collectives and mbarrier phases are annotated as
// release/acquire pairs, atomics map to real atomics, and everything else -- every shared- or
// global-memory access of the kernels -- is checked by TSan for a happens-before order.  A missing
// __syncthreads / mbarrier wait is then reported as a race no matter which schedule ran.
#ifdef EMU_TSAN
extern "C" {
void* __tsan_get_current_fiber(void);
void* __tsan_create_fiber(unsigned flags);
void __tsan_destroy_fiber(void* fiber);
void __tsan_switch_to_fiber(void* fiber, unsigned flags);
void __tsan_acquire(void* addr);
void __tsan_release(void* addr);
}
#define EMU_HB_RELEASE(p) __tsan_release((void*)(p))
#define EMU_HB_ACQUIRE(p) __tsan_acquire((void*)(p))
#define EMU_INTERNAL __attribute__((no_sanitize("thread")))  // the interpreter's own bookkeeping is not the subject
#else
#define EMU_HB_RELEASE(p) ((void)0)
#define EMU_HB_ACQUIRE(p) ((void)0)
#define EMU_INTERNAL
#endif

namespace emu {

constexpr int kMaxThreads = 1024;
constexpr size_t kStack = 256 * 1024;

struct Lane {
    ucontext_t ctx;
    bool done = true;
};
struct WarpState {
    uint64_t slot[32];
    int arrived = 0;
    unsigned gen = 0;
};
struct BarState {
    int arrived = 0;
    unsigned gen = 0;
};
struct State {
    uint3 tidx{0, 0, 0}, bidx{0, 0, 0};
    dim3 bdim{1, 1, 1}, gdim{1, 1, 1};
    int nthreads = 0, cur = 0;
    bool progress = false;
    ucontext_t sched;
    Lane lanes[kMaxThreads];
    WarpState warps[kMaxThreads / 32];
    BarState bars[16];
    std::function<void()> body;
    char* stacks = nullptr;
    uint64_t collectives = 0, block_barriers = 0;  // statistics (per launch)
    int schedule = 0;                               // see run_block()
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    char block_token = 0;                           // happens-before edge between consecutive blocks / launches
#ifdef EMU_TSAN
    void* sched_fiber = nullptr;
    void* fibers[kMaxThreads] = {};
#endif
};
EMU_INTERNAL inline State& S()
{
    static State* s = new State();
    return *s;
}

EMU_INTERNAL [[noreturn]] inline void die(const char* what)
{
    State& s = S();
    std::fprintf(stderr, "simt_emu: %s (block %u, thread %d)\n", what, s.bidx.x, s.cur);
    std::abort();
}

EMU_INTERNAL inline void yield()
{
    State& s = S();
#ifdef EMU_TSAN
    __tsan_switch_to_fiber(s.sched_fiber, 1);  // 1 = no_sync: a context switch is not a synchronisation
#endif
    swapcontext(&s.lanes[s.cur].ctx, &s.sched);
}

EMU_INTERNAL inline void trampoline()
{
    State& s = S();
    EMU_HB_ACQUIRE(&s.block_token);  // blocks run one after another: ordered after the previous block and the host
    s.body();
    EMU_HB_RELEASE(&s.block_token);
    s.lanes[s.cur].done = true;
    s.progress = true;
#ifdef EMU_TSAN
    __tsan_switch_to_fiber(s.sched_fiber, 1);  // 1 = no_sync: a context switch is not a synchronisation
#endif
    // returning switches to uc_link == &s.sched
}

EMU_INTERNAL inline void run_block(int nthreads)
{
    State& s = S();
    if (nthreads > kMaxThreads) die("block too large");
    if (!s.stacks) s.stacks = (char*)std::malloc(kStack * kMaxThreads);
    s.nthreads = nthreads;
    for (int w = 0; w < kMaxThreads / 32; ++w) s.warps[w] = WarpState();
    for (int b = 0; b < 16; ++b) s.bars[b] = BarState();
    for (int t = 0; t < nthreads; ++t) {
        Lane& l = s.lanes[t];
        getcontext(&l.ctx);
        l.ctx.uc_stack.ss_sp = s.stacks + kStack * (size_t)t;
        l.ctx.uc_stack.ss_size = kStack;
        l.ctx.uc_link = &s.sched;
        makecontext(&l.ctx, trampoline, 0);
        l.done = false;
#ifdef EMU_TSAN
        if (!s.fibers[t]) s.fibers[t] = __tsan_create_fiber(0);
#endif
    }
#ifdef EMU_TSAN
    s.sched_fiber = __tsan_get_current_fiber();
#endif
    EMU_HB_RELEASE(&s.block_token);  // what the host wrote before the launch is visible to the block
    // Resume order of the threads within a pass: 0 = ascending, 1 = descending, 2 = a fresh pseudo-random
    // permutation every pass (set_schedule()).  Results of a correctly synchronised kernel do not depend
    // on it; a read that is not ordered after its write by a barrier shows up as a difference.
    int remaining = nthreads;
    std::vector<int> order(nthreads);
    for (int t = 0; t < nthreads; ++t) order[t] = t;
    while (remaining) {
        s.progress = false;
        if (s.schedule == 2) {
            for (int i = nthreads - 1; i > 0; --i) {
                s.rng = s.rng * 6364136223846793005ull + 1442695040888963407ull;
                std::swap(order[i], order[(int)((s.rng >> 33) % (uint64_t)(i + 1))]);
            }
        }
        for (int k = 0; k < nthreads; ++k) {
            const int t = s.schedule == 1 ? nthreads - 1 - k : order[k];
            Lane& l = s.lanes[t];
            if (l.done) continue;
            s.cur = t;
            s.tidx = uint3{(unsigned)t, 0, 0};
#ifdef EMU_TSAN
            __tsan_switch_to_fiber(s.fibers[t], 1);
#endif
            swapcontext(&s.sched, &l.ctx);
            if (l.done) --remaining;
        }
        if (remaining && !s.progress) {
            s.cur = -1;
            die("deadlock: every live thread waits and nobody can make progress");
        }
    }
    EMU_HB_ACQUIRE(&s.block_token);  // the host reads results after the block
}

EMU_INTERNAL inline void set_schedule(int mode) { S().schedule = mode; }

// kernel<<<grid, block>>>(args...)  ==  emu::launch(grid, block, [&] { kernel(args...); })
template <typename F>
EMU_INTERNAL inline void launch(unsigned grid, unsigned block, F&& f)
{
    State& s = S();
    s.gdim = dim3(grid, 1, 1);
    s.bdim = dim3(block, 1, 1);
    s.body = std::forward<F>(f);
    for (unsigned b = 0; b < grid; ++b) {
        s.bidx = uint3{b, 0, 0};
        run_block((int)block);
    }
}

// ---- barriers ---------------------------------------------------------------------------------
EMU_INTERNAL inline void block_barrier(int id, int expected)
{
    State& s = S();
    if (id < 0 || id >= 16) die("bad barrier id");
    BarState& b = s.bars[id];
    const unsigned g = b.gen;
    s.progress = true;
    ++s.block_barriers;
    EMU_HB_RELEASE(&b);
    if (++b.arrived == expected) {
        b.arrived = 0;
        ++b.gen;
    } else {
        while (b.gen == g) yield();
    }
    EMU_HB_ACQUIRE(&b);
}
EMU_INTERNAL inline int warp_width()
{
    State& s = S();
    const int base = (s.cur >> 5) << 5;
    return s.nthreads - base < 32 ? s.nthreads - base : 32;
}
EMU_INTERNAL inline void warp_barrier()
{
    State& s = S();
    WarpState& w = s.warps[s.cur >> 5];
    const unsigned g = w.gen;
    s.progress = true;
    EMU_HB_RELEASE(&w);
    if (++w.arrived == warp_width()) {
        w.arrived = 0;
        ++w.gen;
    } else {
        while (w.gen == g) yield();
    }
    EMU_HB_ACQUIRE(&w);
}
EMU_INTERNAL inline void need_full(unsigned mask)
{
    if (mask != 0xffffffffu) die("warp collective with a partial mask (not modelled)");
    if (warp_width() != 32) die("warp collective in a partial warp");
}

// every lane deposits a value, f(slots) is evaluated by every lane after all have arrived
template <typename T, typename F>
EMU_INTERNAL inline auto collective(unsigned mask, T v, F&& f)
{
    static_assert(sizeof(T) <= 8, "collectives carry at most 8 bytes");
    need_full(mask);
    State& s = S();
    WarpState& w = s.warps[s.cur >> 5];
    const int lane = s.cur & 31;
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    w.slot[lane] = bits;
    ++s.collectives;
    warp_barrier();
    auto get = [&](int l) {
        T r;
        std::memcpy(&r, &w.slot[l & 31], sizeof(T));
        return r;
    };
    auto r = f(get, lane);
    warp_barrier();  // slots may be overwritten only after everyone has read them
    return r;
}

}  // namespace emu

// ---- CUDA built-ins ------------------------------------------------------------------------------
#define threadIdx (emu::S().tidx)
#define blockIdx (emu::S().bidx)
#define blockDim (emu::S().bdim)
#define gridDim (emu::S().gdim)

inline void __syncthreads() { emu::block_barrier(0, emu::S().nthreads); }
inline void __syncwarp(unsigned mask = 0xffffffffu)
{
    emu::need_full(mask);
    emu::warp_barrier();
}
inline void __threadfence() {}
inline void __threadfence_block() {}

template <typename T>
inline T __ldg(const T* p)
{
    return *p;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __brev(unsigned v)
{
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift)
{
    return (unsigned)((((uint64_t)hi << 32) | lo) >> (shift & 31));
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift)
{
    return (unsigned)(((((uint64_t)hi << 32) | lo) << (shift & 31)) >> 32);
}

template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src)
{
    return emu::collective(mask, v, [&](auto get, int) { return get(src); });
}
template <typename T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned d)
{
    return emu::collective(mask, v, [&](auto get, int lane) { return lane >= (int)d ? get(lane - (int)d) : v; });
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned d)
{
    return emu::collective(mask, v, [&](auto get, int lane) { return lane + (int)d < 32 ? get(lane + (int)d) : v; });
}
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int x)
{
    return emu::collective(mask, v, [&](auto get, int lane) { return get(lane ^ x); });
}
inline unsigned __ballot_sync(unsigned mask, int pred)
{
    return emu::collective(mask, pred ? 1 : 0, [&](auto get, int) {
        unsigned b = 0;
        for (int l = 0; l < 32; ++l) b |= (unsigned)(get(l) != 0) << l;
        return b;
    });
}
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
inline unsigned __activemask() { return 0xffffffffu; }

#define EMU_REDUCE(name, T, init, expr)                                  \
    inline T name(unsigned mask, T v)                                    \
    {                                                                    \
        return emu::collective(mask, v, [&](auto get, int) {             \
            T acc = init;                                                \
            for (int l = 0; l < 32; ++l) {                               \
                T e = get(l);                                            \
                acc = expr;                                              \
            }                                                            \
            return acc;                                                  \
        });                                                              \
    }
EMU_REDUCE(__reduce_add_sync, int, 0, acc + e)
EMU_REDUCE(__reduce_add_sync, unsigned, 0u, acc + e)
EMU_REDUCE(__reduce_min_sync, int, INT32_MAX, (e < acc ? e : acc))
EMU_REDUCE(__reduce_min_sync, unsigned, UINT32_MAX, (e < acc ? e : acc))
EMU_REDUCE(__reduce_max_sync, int, INT32_MIN, (e > acc ? e : acc))
EMU_REDUCE(__reduce_max_sync, unsigned, 0u, (e > acc ? e : acc))
EMU_REDUCE(__reduce_or_sync, unsigned, 0u, acc | e)
EMU_REDUCE(__reduce_and_sync, unsigned, 0xffffffffu, acc& e)
#undef EMU_REDUCE

template <typename T>
inline T atomicAdd(T* p, T v)
{
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
template <typename T>
inline T atomicOr(T* p, T v)
{
    return __atomic_fetch_or(p, v, __ATOMIC_RELAXED);
}
template <typename T>
inline T atomicMax(T* p, T v)
{
    T old = *p;
    if (v > old) *p = v;
    return old;
}
template <typename T>
inline T atomicMin(T* p, T v)
{
    T old = *p;
    if (v < old) *p = v;
    return old;
}
template <typename T>
inline T atomicExch(T* p, T v)
{
    T old = *p;
    *p = v;
    return old;
}

inline float __fmul_rn(float a, float b)
{
    volatile float r = a * b;  // volatile: the product is rounded before any use
    return r;
}
inline double __dmul_rn(double a, double b)
{
    volatile double r = a * b;
    return r;
}

// CUDA's global-namespace integer min / max
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline long min(long a, long b) { return a < b ? a : b; }  // int64_t is long on LP64
inline long max(long a, long b) { return a > b ? a : b; }
using std::fma;  // <cmath>: float and double overloads
