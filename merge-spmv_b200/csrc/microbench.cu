// microbench.cu -- the machine limits that bound merge-based CsrMV on one B200:
//   (1) streaming read bandwidth with 128-bit LDG,
//   (2) streaming read bandwidth with cp.async.bulk (TMA) into a shared-memory ring,
//   (3) x-gather throughput from an L2-resident vector with coalesced index stream,
//   (4) stream + gather + multiply-add with no row structure (an upper bound for SpMV).
// Prints one line per measurement.  Not part of the library; used to set expectations in
// DESIGN.md and profiles/.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            std::exit(1);                                                              \
        }                                                                              \
    } while (0)

static __host__ __device__ inline uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void fill_indices(int* idx, size_t n, int table, int per_row, int mode, uint64_t seed)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mode == 0) {  // uniform random
        idx[i] = (int)(splitmix64(seed + i) % (uint64_t)table);
    } else if (mode == 1) {  // stratified sorted: per_row entries per row, one per stratum
        size_t j = i % per_row;
        uint64_t lo = (uint64_t)table * j / per_row, hi = (uint64_t)table * (j + 1) / per_row;
        idx[i] = (int)(lo + splitmix64(seed + i) % (hi - lo));
    } else {  // banded: row r = i / per_row, col = r - per_row/2 + j
        long long r = (long long)(i / per_row), j = (long long)(i % per_row);
        long long c = r - per_row / 2 + j;
        if (c < 0) c = 0;
        if (c >= table) c = table - 1;
        idx[i] = (int)c;
    }
}
template <typename T>
__global__ void fill_vals(T* v, size_t n, uint64_t seed)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] = (T)(0.5 + (double)(splitmix64(seed + i) >> 11) * (1.0 / 9007199254740992.0));
}

// (1) 128-bit LDG streaming read
__global__ void stream_read_ldg(const int4* __restrict__ p, size_t n16, unsigned long long* sink)
{
    size_t stride = (size_t)gridDim.x * blockDim.x;
    int acc = 0;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        int4 a = __ldcs(p + i), b = __ldcs(p + i + stride), c = __ldcs(p + i + 2 * stride),
             d = __ldcs(p + i + 3 * stride);
        acc += a.x ^ b.y ^ c.z ^ d.w;
    }
    for (; i < n16; i += stride) acc += __ldcs(p + i).x;
    if (acc == 0x7fffffff) atomicAdd(sink, 1ull);
}

// (2) TMA bulk streaming read into an smem ring
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int CHUNK_BYTES, int SLOTS>
__global__ void stream_read_tma(const char* __restrict__ p, size_t bytes, unsigned long long* sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[SLOTS], empty[SLOTS];
    size_t nchunks = bytes / CHUNK_BYTES;
    size_t per = (nchunks + gridDim.x - 1) / gridDim.x;
    size_t c0 = per * blockIdx.x, c1 = c0 + per < nchunks ? c0 + per : nchunks;
    if (threadIdx.x == 0) {
        for (int i = 0; i < SLOTS; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[i])), "r"(4));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 4) {
        if (lane == 0) {
            for (size_t c = c0; c < c1; ++c) {
                size_t k = c - c0;
                int slot = (int)(k % SLOTS);
                if (k >= SLOTS) {
                    uint32_t ok, par = (uint32_t)((k / SLOTS - 1) & 1);
                    do {
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                            "selp.u32 %0, 1, 0, p;\n\t}"
                            : "=r"(ok)
                            : "r"(smem_u32(&empty[slot])), "r"(par)
                            : "memory");
                    } while (!ok);
                }
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[slot])),
                             "r"(CHUNK_BYTES)
                             : "memory");
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        smem_u32(smem + (size_t)slot * CHUNK_BYTES)),
                    "l"(p + c * CHUNK_BYTES), "r"(CHUNK_BYTES), "r"(smem_u32(&full[slot]))
                    : "memory");
            }
        }
        return;
    }
    int acc = 0;
    for (size_t c = c0; c < c1; ++c) {
        size_t k = c - c0;
        int slot = (int)(k % SLOTS);
        uint32_t ok, par = (uint32_t)((k / SLOTS) & 1);
        do {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(smem_u32(&full[slot])), "r"(par)
                : "memory");
        } while (!ok);
        const int4* s = reinterpret_cast<const int4*>(smem + (size_t)slot * CHUNK_BYTES);
        for (int i = threadIdx.x; i < CHUNK_BYTES / 16; i += 128) acc += s[i].x;
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[slot])) : "memory");
    }
    if (acc == 0x7fffffff) atomicAdd(sink, 1ull);
}

// (3)/(4) gather (+ optional value stream)
template <typename T, int U, bool WITH_VALUES>
__global__ void gather_kernel(const int* __restrict__ idx, const T* __restrict__ vals,
                              const T* __restrict__ table, size_t n, T* out)
{
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    T acc = 0;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        int c[U];
        T v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = __ldcs(idx + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = WITH_VALUES ? __ldcs(vals + i + u * stride) : T(1);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u] * __ldg(table + c[u]);
    }
    for (; i < n; i += stride) acc += (WITH_VALUES ? vals[i] : T(1)) * __ldg(table + idx[i]);
    if (acc == T(-12345)) out[0] = acc;
}

// gather flavours vs. L1 capacity: MODE 0 = ld.global.nc, 1 = ld.global.cg, 2 = ld.global.nc.L1::no_allocate,
// 3 = ld.global.cv (volatile)
template <int MODE>
__device__ __forceinline__ double ld_flavour(const double* p)
{
    double v;
    if (MODE == 0) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    if (MODE == 1) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    if (MODE == 2) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    if (MODE == 3) asm volatile("ld.global.cv.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
template <int MODE, int U>
__global__ void gather_flavour_kernel(const int* __restrict__ idx, const double* __restrict__ table, size_t n,
                                      double* out)
{
    extern __shared__ unsigned char pad[];
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    double acc = 0;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        int c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = __ldcs(idx + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += ld_flavour<MODE>(table + c[u]);
    }
    if (acc == -12345.0) out[0] = acc + pad[threadIdx.x];
}

template <typename F>
static float time_ms(F launch, int iters = 5)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    launch();
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < iters; ++i) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

template <typename T>
static void gather_suite(const char* tname, int sms)
{
    const size_t n = 64ull << 20;  // 64 Mi gathers
    int* idx;
    T *vals, *table, *out;
    CK(cudaMalloc(&idx, n * sizeof(int)));
    CK(cudaMalloc(&vals, n * sizeof(T)));
    CK(cudaMalloc(&out, 64));
    const int tables[] = {1 << 12, 1 << 14, 1 << 20, 2000000, 1 << 23, 20000000};
    const char* modes[] = {"uniform", "stratified64", "banded7"};
    fill_vals<T><<<(unsigned)((n + 255) / 256), 256>>>(vals, n, 7);
    for (int table_n : tables) {
        CK(cudaMalloc(&table, (size_t)table_n * sizeof(T)));
        fill_vals<T><<<(table_n + 255) / 256, 256>>>(table, table_n, 3);
        for (int mode = 0; mode < 3; ++mode) {
            fill_indices<<<(unsigned)((n + 255) / 256), 256>>>(idx, n, table_n, mode == 2 ? 7 : 64, mode, 11);
            CK(cudaDeviceSynchronize());
            for (int with_vals = 0; with_vals < 2; ++with_vals) {
                float ms;
                dim3 grid(sms * 8), block(256);
                if (with_vals)
                    ms = time_ms([&] { gather_kernel<T, 8, true><<<grid, block>>>(idx, vals, table, n, out); });
                else
                    ms = time_ms([&] { gather_kernel<T, 8, false><<<grid, block>>>(idx, vals, table, n, out); });
                double bytes = (double)n * (4 + (with_vals ? sizeof(T) : 0));
                std::printf("gather %s table=%d (%.1f MB) idx=%s values=%d: %.3f ms  %.1f Ggather/s  stream %.0f GB/s\n",
                            tname, table_n, table_n * sizeof(T) / 1e6, modes[mode], with_vals, ms,
                            n / ms / 1e6, bytes / ms / 1e6);
            }
        }
        CK(cudaFree(table));
    }
    CK(cudaFree(idx));
    CK(cudaFree(vals));
    CK(cudaFree(out));
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    std::printf("device %s, %d SMs, L2 %.1f MB, smem/SM %zu KB\n", prop.name, sms, prop.l2CacheSize / 1e6,
                prop.sharedMemPerMultiprocessor / 1024);

    unsigned long long* sink;
    CK(cudaMalloc(&sink, 8));
    const size_t bytes = 2ull << 30;
    char* buf;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 1, bytes));

    for (int mult : {2, 4, 8, 16}) {
        float ms = time_ms([&] {
            stream_read_ldg<<<sms * mult, 256>>>(reinterpret_cast<const int4*>(buf), bytes / 16, sink);
        });
        std::printf("stream_read_ldg grid=%dxSM x256thr: %.3f ms  %.0f GB/s\n", mult, ms, bytes / ms / 1e6);
    }
    {
        auto k = stream_read_tma<8192, 8>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
        for (int mult : {1, 2, 3}) {
            float ms = time_ms([&] { k<<<sms * mult, 160, 8192 * 8>>>(buf, bytes, sink); });
            std::printf("stream_read_tma chunk=8K slots=8 grid=%dxSM: %.3f ms  %.0f GB/s\n", mult, ms,
                        bytes / ms / 1e6);
        }
        auto k2 = stream_read_tma<4096, 8>;
        CK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 8));
        for (int mult : {2, 4}) {
            float ms = time_ms([&] { k2<<<sms * mult, 160, 4096 * 8>>>(buf, bytes, sink); });
            std::printf("stream_read_tma chunk=4K slots=8 grid=%dxSM: %.3f ms  %.0f GB/s\n", mult, ms,
                        bytes / ms / 1e6);
        }
        auto k3 = stream_read_tma<16384, 6>;
        CK(cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 6));
        for (int mult : {1, 2}) {
            float ms = time_ms([&] { k3<<<sms * mult, 160, 16384 * 6>>>(buf, bytes, sink); });
            std::printf("stream_read_tma chunk=16K slots=6 grid=%dxSM: %.3f ms  %.0f GB/s\n", mult, ms,
                        bytes / ms / 1e6);
        }
    }
    CK(cudaFree(buf));
    {   // gather flavour x shared-memory carve-out (L1 capacity) experiment
        const size_t n = 64ull << 20;
        int* idx;
        double *table, *out;
        CK(cudaMalloc(&idx, n * sizeof(int)));
        CK(cudaMalloc(&table, (1 << 20) * sizeof(double)));
        CK(cudaMalloc(&out, 64));
        fill_vals<double><<<(1 << 20) / 256, 256>>>(table, 1 << 20, 3);
        fill_indices<<<(unsigned)((n + 255) / 256), 256>>>(idx, n, 1 << 20, 64, 0, 11);
        CK(cudaDeviceSynchronize());
        const int smem_per_block[] = {0, 8 << 10, 16 << 10, 20 << 10, 24 << 10, 27 << 10};
        auto run = [&](auto kern, const char* name) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 27 << 10));
            for (int sb : smem_per_block) {
                float ms = time_ms([&] { kern<<<sms * 8, 256, sb>>>(idx, table, n, out); });
                std::printf("gather_flavour %s smem/SM=%3d KB: %.3f ms  %.1f Ggather/s\n", name, sb * 8 >> 10, ms,
                            n / ms / 1e6);
            }
        };
        run(gather_flavour_kernel<0, 8>, "ld.global.nc      ");
        run(gather_flavour_kernel<1, 8>, "ld.global.cg      ");
        run(gather_flavour_kernel<2, 8>, "ld.nc.no_allocate ");
        run(gather_flavour_kernel<3, 8>, "ld.global.cv      ");
        CK(cudaFree(idx));
        CK(cudaFree(table));
        CK(cudaFree(out));
    }
    if (std::getenv("MICROBENCH_FLAVOURS_ONLY")) return 0;
    gather_suite<double>("f64", sms);
    gather_suite<float>("f32", sms);
    std::printf("microbench done\n");
    return 0;
}
