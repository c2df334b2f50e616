// mg_session.cu -- multi-GPU host-buffer operator behind the C ABI (include/mergespmv.h section 6):
// one process, p devices with peer access (NVLink / NVSwitch), no NCCL and no Python.
//
// New surface: the reference is single-GPU (README.md:5).  The cut is the one OmpMergeCsrmv makes
// between CPU threads (cpu_spmv.cpp:311-321) with p = number of shards: shard g owns diagonals
// [g*ceil((rows+nnz)/p), ...) of the merge path, i.e. nonzeros [y_g, y_{g+1}) and the rows
// [x_g, x_{g+1}) that END inside its span; its local CSR has one extra last row -- the leading part
// of global row x_{g+1} -- whose result is the carry-out (cpu_spmv.cpp:336-344).  Device g runs the
// unchanged single-GPU CsrMV (mspmv_csrmv_*) on its shard and then ONE exchange kernel
// (mspmv_exchange_carries_*, carry_exchange.cuh): it stores its carry into every peer's exchange
// buffer over NVLink, waits for the peers' flags and folds carries 0..p-2 into the rows it owns in
// shard order with the row < num_rows guard -- the serial fix-up of cpu_spmv.cpp:348-352.
//
// Per right-hand side (apply_many, K = 3 slots, three streams per device):
//   x host -> device of shard 0 (one PCIe crossing) -> peer copies to the other devices (NVLink)
//   | CsrMV + carry exchange on every device | every device's y slice -> its place in y host.
// This file uses nothing but the C ABI above it and the CUDA runtime.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/mergespmv.h"

#define MG_TRY(expr)                            \
    do {                                        \
        cudaError_t _e = (cudaError_t)(expr);   \
        if (_e != cudaSuccess) return (int)_e;  \
    } while (0)

namespace {
constexpr int kSlots = 3;

// the caller's current device is restored on every exit path
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { cudaGetDevice(&prev); }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct MgDev {
    int device = 0;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0, local_rows = 0, owned = 0, nnz = 0;
    int *ro = nullptr, *col = nullptr, *carry_rows = nullptr;
    void *val = nullptr, *temp = nullptr, *xbuf = nullptr;
    void** peers = nullptr;
    unsigned long long* epoch = nullptr;
    size_t temp_bytes = 0;
    void* x[kSlots] = {nullptr, nullptr, nullptr};
    void* y[kSlots] = {nullptr, nullptr, nullptr};
    cudaStream_t s_in = nullptr, s_compute = nullptr, s_out = nullptr;
    cudaEvent_t ev_x[kSlots], ev_k[kSlots], ev_y[kSlots], t0 = nullptr, t1 = nullptr;
    bool events = false;
};
}  // namespace

struct mspmv_mg_session {
    int p = 0, value_bytes = 0, rows = 0, cols = 0, nnz = 0;
    std::vector<MgDev> d;
    int last_slot = 0;
};

static int mg_csrmv(mspmv_mg_session* s, MgDev& D, int slot)
{
    if (s->value_bytes == 8)
        return mspmv_csrmv_f64(D.temp, &D.temp_bytes, (const double*)D.val, D.ro, D.col, (const double*)D.x[slot],
                               (double*)D.y[slot], D.local_rows, s->cols, D.nnz, D.s_compute, 0);
    return mspmv_csrmv_f32(D.temp, &D.temp_bytes, (const float*)D.val, D.ro, D.col, (const float*)D.x[slot],
                           (float*)D.y[slot], D.local_rows, s->cols, D.nnz, D.s_compute, 0);
}

static int mg_exchange(mspmv_mg_session* s, MgDev& D, int g, int slot)
{
    if (s->p == 1) return 0;
    if (s->value_bytes == 8)
        return mspmv_exchange_carries_f64((double*)D.y[slot], D.local_rows, D.x0, D.owned, s->rows, D.carry_rows, D.peers,
                                          g, s->p, D.epoch, D.s_compute);
    return mspmv_exchange_carries_f32((float*)D.y[slot], D.local_rows, D.x0, D.owned, s->rows, D.carry_rows, D.peers, g,
                                      s->p, D.epoch, D.s_compute);
}

// the device part of one product on every shard, x already in slot `slot` of every device
static int mg_launch_products(mspmv_mg_session* s, int slot)
{
    for (int g = 0; g < s->p; ++g) {
        MgDev& D = s->d[g];
        MG_TRY(cudaSetDevice(D.device));
        int rc = mg_csrmv(s, D, slot);
        if (rc) return rc;
        rc = mg_exchange(s, D, g, slot);
        if (rc) return rc;
    }
    return 0;
}

extern "C" {

int mspmv_mg_session_create(mspmv_mg_session** out, int num_shards, const int* device_ids, int value_bytes,
                            int num_rows, int num_cols, int num_nonzeros, const int* row_offsets,
                            const int* column_indices, const void* values)
{
    if (!out || num_shards < 1 || num_shards > 64 || (value_bytes != 4 && value_bytes != 8) || num_rows < 0 ||
        num_nonzeros < 0 || !row_offsets)
        return (int)cudaErrorInvalidValue;
    int prev = 0;
    MG_TRY(cudaGetDevice(&prev));
    mspmv_mg_session* s = new mspmv_mg_session();
    s->p = num_shards, s->value_bytes = value_bytes, s->rows = num_rows, s->cols = num_cols, s->nnz = num_nonzeros;
    s->d.resize(num_shards);
    const size_t vb = (size_t)value_bytes;
    std::vector<int> coords(2 * (size_t)(num_shards + 1));
    mspmv_shard_partition(row_offsets, num_rows, num_nonzeros, num_shards, coords.data());
    std::vector<int> carry_rows(num_shards);
    for (int g = 0; g < num_shards; ++g) carry_rows[g] = coords[2 * (g + 1)];
    const size_t xbytes = mspmv_exchange_buffer_bytes(num_shards);
    int rc = 0;
    auto fail = [&](int e) {
        rc = e;
        return e != 0;
    };
    for (int g = 0; g < num_shards && !rc; ++g) {
        MgDev& D = s->d[g];
        D.device = device_ids ? device_ids[g] : g;
        if (fail(cudaSetDevice(D.device))) break;
        for (int h = 0; h < num_shards; ++h) {  // peer access to every other device that holds a shard
            const int other = device_ids ? device_ids[h] : h;
            if (other == D.device) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(other, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (fail(e)) break;
        }
        if (rc) break;
        D.x0 = coords[2 * g], D.y0 = coords[2 * g + 1], D.x1 = coords[2 * g + 2], D.y1 = coords[2 * g + 3];
        D.owned = D.x1 - D.x0, D.local_rows = D.owned + 1, D.nnz = D.y1 - D.y0;
        std::vector<int> lro((size_t)D.local_rows + 1);
        mspmv_shard_row_offsets(row_offsets, D.x0, D.y0, D.x1, D.y1, lro.data());
        const size_t n1 = (size_t)(D.nnz > 0 ? D.nnz : 1);
        if (fail(cudaMalloc(&D.val, vb * n1)) || fail(cudaMalloc(&D.col, sizeof(int) * n1)) ||
            fail(cudaMalloc(&D.ro, sizeof(int) * ((size_t)D.local_rows + 1))) ||
            fail(cudaMalloc(&D.carry_rows, sizeof(int) * (size_t)num_shards)) || fail(cudaMalloc(&D.xbuf, xbytes)) ||
            fail(cudaMalloc(&D.peers, sizeof(void*) * (size_t)num_shards)) ||
            fail(cudaMalloc(&D.epoch, sizeof(unsigned long long))))
            break;
        for (int k = 0; k < kSlots && !rc; ++k)
            if (fail(cudaMalloc(&D.x[k], vb * (size_t)(num_cols > 0 ? num_cols : 1))) ||
                fail(cudaMalloc(&D.y[k], vb * (size_t)D.local_rows)))
                break;
        if (rc) break;
        if (fail(cudaStreamCreateWithFlags(&D.s_in, cudaStreamNonBlocking)) ||
            fail(cudaStreamCreateWithFlags(&D.s_compute, cudaStreamNonBlocking)) ||
            fail(cudaStreamCreateWithFlags(&D.s_out, cudaStreamNonBlocking)))
            break;
        for (int k = 0; k < kSlots; ++k) {
            cudaEventCreateWithFlags(&D.ev_x[k], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&D.ev_k[k], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&D.ev_y[k], cudaEventDisableTiming);
        }
        cudaEventCreate(&D.t0);
        cudaEventCreate(&D.t1);
        D.events = true;
        if (fail(cudaMemset(D.xbuf, 0, xbytes)) || fail(cudaMemset(D.epoch, 0, sizeof(unsigned long long))) ||
            fail(cudaMemcpy(D.val, (const char*)values + vb * (size_t)D.y0, vb * (size_t)D.nnz, cudaMemcpyHostToDevice)) ||
            fail(cudaMemcpy(D.col, column_indices + D.y0, sizeof(int) * (size_t)D.nnz, cudaMemcpyHostToDevice)) ||
            fail(cudaMemcpy(D.ro, lro.data(), sizeof(int) * ((size_t)D.local_rows + 1), cudaMemcpyHostToDevice)) ||
            fail(cudaMemcpy(D.carry_rows, carry_rows.data(), sizeof(int) * (size_t)num_shards, cudaMemcpyHostToDevice)))
            break;
        // temp-storage size query, then allocate (gpu_spmv.cu:390-398)
        size_t bytes = 0;
        rc = value_bytes == 8 ? mspmv_csrmv_f64(nullptr, &bytes, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                D.local_rows, num_cols, D.nnz, nullptr, 0)
                              : mspmv_csrmv_f32(nullptr, &bytes, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                D.local_rows, num_cols, D.nnz, nullptr, 0);
        if (rc || fail(cudaMalloc(&D.temp, bytes))) break;
        D.temp_bytes = bytes;
    }
    if (!rc) {
        std::vector<void*> table(num_shards);
        for (int g = 0; g < num_shards; ++g) table[g] = s->d[g].xbuf;
        for (int g = 0; g < num_shards && !rc; ++g) {
            if (fail(cudaSetDevice(s->d[g].device)) ||
                fail(cudaMemcpy(s->d[g].peers, table.data(), sizeof(void*) * (size_t)num_shards, cudaMemcpyHostToDevice)) ||
                fail(cudaDeviceSynchronize()))
                break;
        }
    }
    cudaSetDevice(prev);
    if (rc) {
        mspmv_mg_session_destroy(s);
        return rc;
    }
    *out = s;
    return 0;
}

int mspmv_mg_session_apply_many(mspmv_mg_session* s, int n, const void* xs_host, void* ys_host)
{
    if (!s || n < 0 || (n > 0 && (!xs_host || !ys_host))) return (int)cudaErrorInvalidValue;
    DeviceGuard guard;
    const size_t vb = (size_t)s->value_bytes;
    const size_t xbytes = vb * (size_t)s->cols, ybytes = vb * (size_t)s->rows;
    const char* xs = (const char*)xs_host;
    char* ys = (char*)ys_host;
    MgDev& D0 = s->d[0];
    for (int i = 0; i < n; ++i) {
        const int sl = i % kSlots;
        // x: host -> shard 0's device once the slot is free there (its product AND the peers' copies of the
        // previous content are done), then device -> device for the other shards
        MG_TRY(cudaSetDevice(D0.device));
        if (i >= kSlots) {
            MG_TRY(cudaStreamWaitEvent(D0.s_in, D0.ev_k[sl], 0));
            for (int g = 1; g < s->p; ++g) MG_TRY(cudaStreamWaitEvent(D0.s_in, s->d[g].ev_x[sl], 0));
        }
        MG_TRY(cudaMemcpyAsync(D0.x[sl], xs + xbytes * (size_t)i, xbytes, cudaMemcpyHostToDevice, D0.s_in));
        MG_TRY(cudaEventRecord(D0.ev_x[sl], D0.s_in));
        for (int g = 1; g < s->p; ++g) {
            MgDev& D = s->d[g];
            MG_TRY(cudaSetDevice(D.device));
            if (i >= kSlots) MG_TRY(cudaStreamWaitEvent(D.s_in, D.ev_k[sl], 0));
            MG_TRY(cudaStreamWaitEvent(D.s_in, D0.ev_x[sl], 0));
            if (D.device == D0.device)
                MG_TRY(cudaMemcpyAsync(D.x[sl], D0.x[sl], xbytes, cudaMemcpyDeviceToDevice, D.s_in));
            else
                MG_TRY(cudaMemcpyPeerAsync(D.x[sl], D.device, D0.x[sl], D0.device, xbytes, D.s_in));
            MG_TRY(cudaEventRecord(D.ev_x[sl], D.s_in));
        }
        // product + carry exchange on every shard, then its slice of y back to the host
        for (int g = 0; g < s->p; ++g) {
            MgDev& D = s->d[g];
            MG_TRY(cudaSetDevice(D.device));
            MG_TRY(cudaStreamWaitEvent(D.s_compute, D.ev_x[sl], 0));
            if (i >= kSlots) MG_TRY(cudaStreamWaitEvent(D.s_compute, D.ev_y[sl], 0));
            int rc = mg_csrmv(s, D, sl);
            if (rc) return rc;
            rc = mg_exchange(s, D, g, sl);
            if (rc) return rc;
            MG_TRY(cudaEventRecord(D.ev_k[sl], D.s_compute));
            MG_TRY(cudaStreamWaitEvent(D.s_out, D.ev_k[sl], 0));
            if (D.owned > 0)
                MG_TRY(cudaMemcpyAsync(ys + ybytes * (size_t)i + vb * (size_t)D.x0, D.y[sl], vb * (size_t)D.owned,
                                       cudaMemcpyDeviceToHost, D.s_out));
            MG_TRY(cudaEventRecord(D.ev_y[sl], D.s_out));
        }
        s->last_slot = sl;
    }
    for (int g = 0; g < s->p; ++g) {
        MgDev& D = s->d[g];
        MG_TRY(cudaSetDevice(D.device));
        MG_TRY(cudaStreamSynchronize(D.s_out));
        MG_TRY(cudaStreamSynchronize(D.s_compute));
        MG_TRY(cudaStreamSynchronize(D.s_in));
    }
    return 0;
}

int mspmv_mg_session_apply(mspmv_mg_session* s, const void* x_host, void* y_host)
{
    return mspmv_mg_session_apply_many(s, 1, x_host, y_host);
}

int mspmv_mg_session_time_device(mspmv_mg_session* s, int iterations, float* ms_per_step)
{
    if (!s || iterations < 1 || !ms_per_step) return (int)cudaErrorInvalidValue;
    DeviceGuard guard;
    const int sl = s->last_slot;  // the x of the last apply (before any apply: whatever the slot holds -- timing only)
    int rc = mg_launch_products(s, sl);  // warm-up
    if (rc) return rc;
    for (int g = 0; g < s->p; ++g) {
        MG_TRY(cudaSetDevice(s->d[g].device));
        MG_TRY(cudaStreamSynchronize(s->d[g].s_compute));
    }
    for (int g = 0; g < s->p; ++g) {
        MG_TRY(cudaSetDevice(s->d[g].device));
        MG_TRY(cudaEventRecord(s->d[g].t0, s->d[g].s_compute));
    }
    for (int it = 0; it < iterations; ++it) {
        rc = mg_launch_products(s, sl);
        if (rc) return rc;
    }
    for (int g = 0; g < s->p; ++g) {
        MG_TRY(cudaSetDevice(s->d[g].device));
        MG_TRY(cudaEventRecord(s->d[g].t1, s->d[g].s_compute));
    }
    float worst = 0.f;
    for (int g = 0; g < s->p; ++g) {  // device-side time, max over devices
        MG_TRY(cudaSetDevice(s->d[g].device));
        MG_TRY(cudaEventSynchronize(s->d[g].t1));
        float ms = 0.f;
        MG_TRY(cudaEventElapsedTime(&ms, s->d[g].t0, s->d[g].t1));
        if (ms > worst) worst = ms;
    }
    *ms_per_step = worst / (float)iterations;
    return 0;
}

int mspmv_mg_session_shard(const mspmv_mg_session* s, int shard, int* out)
{
    if (!s || !out || shard < 0 || shard >= s->p) return (int)cudaErrorInvalidValue;
    const MgDev& D = s->d[shard];
    out[0] = D.x0, out[1] = D.y0, out[2] = D.x1, out[3] = D.y1, out[4] = D.device;
    return 0;
}

void mspmv_mg_session_destroy(mspmv_mg_session* s)
{
    if (!s) return;
    int prev = 0;
    cudaGetDevice(&prev);
    for (MgDev& D : s->d) {
        if (cudaSetDevice(D.device) != cudaSuccess) continue;
        if (D.s_compute) cudaStreamSynchronize(D.s_compute);
        if (D.s_in) cudaStreamSynchronize(D.s_in);
        if (D.s_out) cudaStreamSynchronize(D.s_out);
        cudaFree(D.val), cudaFree(D.col), cudaFree(D.ro), cudaFree(D.carry_rows), cudaFree(D.xbuf), cudaFree(D.peers);
        cudaFree(D.epoch), cudaFree(D.temp);
        for (int k = 0; k < kSlots; ++k) {
            cudaFree(D.x[k]), cudaFree(D.y[k]);
            if (D.events) cudaEventDestroy(D.ev_x[k]), cudaEventDestroy(D.ev_k[k]), cudaEventDestroy(D.ev_y[k]);
        }
        if (D.events) cudaEventDestroy(D.t0), cudaEventDestroy(D.t1);
        if (D.s_in) cudaStreamDestroy(D.s_in);
        if (D.s_compute) cudaStreamDestroy(D.s_compute);
        if (D.s_out) cudaStreamDestroy(D.s_out);
    }
    cudaSetDevice(prev);
    delete s;
}

}  // extern "C"
