// common.cuh — error handling, stream-ordered device buffers, small device helpers (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <ctime>
#include <map>
#include <stdexcept>
#include <string>

#include "../../include/asgart_b200.h"

namespace ab200 {

using u8 = uint8_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i64 = int64_t;

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

struct CudaError : std::runtime_error {
    int code;
    CudaError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
        int code = (e == cudaErrorMemoryAllocation) ? ASGART_B200_ENOMEM
                   : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? ASGART_B200_ENODEVICE
                                                                                   : ASGART_B200_ECUDA;
        throw CudaError(code, buf);
    }
}
#define CUDA_CHECK(x) ::ab200::cuda_check((x), #x, __FILE__, __LINE__)
#define KERNEL_CHECK() ::ab200::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__)

// Launch counter (bench.py's gpu_launches claim is read from here)
struct LaunchCounter {
    u64 total = 0;
};
extern thread_local LaunchCounter* g_launch_counter;
inline void count_launch(u64 n = 1) {
    if (g_launch_counter) g_launch_counter->total += n;
}

// cudaFuncSetAttribute is per device: a kernel that needs more than 48 KB of dynamic shared memory has to be prepared on
// every device that launches it (several contexts of one process may sit on different devices: build_index_group).
// `prepared` is one bit per device; returns true when the caller still has to set the attribute on the current device —
// it marks the device with device_prepared() AFTER setting it, so a second thread that sees the bit finds the attribute set.
inline bool device_needs_prepare(const std::atomic<unsigned long long>& prepared, unsigned long long& bit) {
    int dev = 0;
    cuda_check(cudaGetDevice(&dev), "cudaGetDevice", __FILE__, __LINE__);
    bit = 1ull << (unsigned(dev) & 63u);
    return (prepared.load(std::memory_order_acquire) & bit) == 0;
}
inline void device_prepared(std::atomic<unsigned long long>& prepared, unsigned long long bit) {
    prepared.fetch_or(bit, std::memory_order_release);
}

// Host-side stall accounting (developer aid: ASGART_B200_DEBUG_TIMING=1 prints it when a context is destroyed)
struct HostStalls {
    double alloc_ms = 0, sync_ms = 0;
    u64 allocs = 0, syncs = 0;
};
extern thread_local HostStalls g_host_stalls;
inline double host_now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// Per-context cache of device blocks. Every step of the pipeline allocates the same few dozen buffers again; handing
// them out of a size-keyed free list costs nothing, while cudaMallocAsync was measured at 27 ms per step once NCCL
// has enabled peer access on the device (each pool allocation is then mapped for the peers). All work of a context
// runs on one stream, so a block released by one kernel's buffer may be handed to the next launch right away.
struct DevicePool {
    std::multimap<size_t, void*> free_blocks;
    size_t held = 0, live = 0;
    static size_t round_size(size_t bytes) {   // size classes of 1/8 octave: at most 12.5 % slack, so blocks get reused
        if (bytes < 4096) return 4096;
        int lg = 63 - __builtin_clzll((unsigned long long)bytes);
        const size_t g = size_t(1) << (lg - 3);
        return (bytes + g - 1) / g * g;
    }
    void* get(size_t bytes) {
        const size_t want = round_size(bytes);
        auto it = free_blocks.lower_bound(want);
        if (it != free_blocks.end() && it->first <= want + want / 4) {
            void* p = it->second;
            held -= it->first;
            live += it->first;
            sizes[p] = it->first;
            free_blocks.erase(it);
            return p;
        }
        void* p = nullptr;
        const double t0 = host_now_ms();
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaErrorMemoryAllocation) {   // give the cached blocks back and try once more
            cudaGetLastError();
            trim();
            e = cudaMalloc(&p, want);
        }
        g_host_stalls.alloc_ms += host_now_ms() - t0;
        g_host_stalls.allocs++;
        cuda_check(e, "cudaMalloc (device pool)", __FILE__, __LINE__);
        sizes[p] = want;
        live += want;
        return p;
    }
    void put(void* p) {
        auto it = sizes.find(p);
        if (it == sizes.end()) return;
        free_blocks.emplace(it->second, p);
        held += it->second;
        live -= it->second;
        sizes.erase(it);
    }
    void trim() {
        cudaDeviceSynchronize();
        for (auto& kv : free_blocks) cudaFree(kv.second);
        free_blocks.clear();
        held = 0;
    }
    ~DevicePool() { trim(); }
    std::map<void*, size_t> sizes;
};
extern thread_local DevicePool* g_device_pool;

// Device buffer: from the calling context's DevicePool when there is one, else stream-ordered (cudaMallocAsync)
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    DevicePool* pool = nullptr;
    DevBuf() = default;
    DevBuf(size_t count, cudaStream_t stream) { alloc(count, stream); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s), pool(o.pool) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; s = o.s; pool = o.pool; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count, cudaStream_t stream) {
        release();
        s = stream;
        n = count;
        if (count == 0) { p = nullptr; return; }
        pool = g_device_pool;
        if (pool) { p = static_cast<T*>(pool->get(count * sizeof(T))); return; }
        const double t0 = host_now_ms();
        CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&p), count * sizeof(T), stream));
        g_host_stalls.alloc_ms += host_now_ms() - t0;
        g_host_stalls.allocs++;
    }
    void release() {
        if (p) {
            if (pool) pool->put(p); else cudaFreeAsync(p, s);
            p = nullptr; n = 0;
        }
    }
    void zero() { if (p) CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    size_t bytes() const { return n * sizeof(T); }
};

// NVTX ranges around the phases of the pipeline (header-only NVTX3: a no-op unless a profiler is attached)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
// consecutive phases of one function: next() closes the running range and opens the named one
struct NvtxPhases {
    bool open = false;
    void next(const char* name) { if (open) nvtxRangePop(); nvtxRangePushA(name); open = true; }
    void close() { if (open) nvtxRangePop(); open = false; }
    ~NvtxPhases() { close(); }
};

inline int ceil_div_i(i64 a, i64 b) { return int((a + b - 1) / b); }
inline u64 ceil_div(u64 a, u64 b) { return (a + b - 1) / b; }
inline int bit_width_u64(u64 v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }

// CUDA-event stopwatch on one stream
struct EventTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t s = nullptr;
    explicit EventTimer(cudaStream_t stream) : s(stream) {
        CUDA_CHECK(cudaEventCreate(&a));
        CUDA_CHECK(cudaEventCreate(&b));
    }
    ~EventTimer() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    void start() { CUDA_CHECK(cudaEventRecord(a, s)); }
    // records the stop event; ms() synchronises on it
    void stop() { CUDA_CHECK(cudaEventRecord(b, s)); }
    double ms() {
        CUDA_CHECK(cudaEventSynchronize(b));
        float t = 0;
        CUDA_CHECK(cudaEventElapsedTime(&t, a, b));
        return double(t);
    }
};

// Accumulating timer for a kernel family launched many times (radix passes, gathers): a ring of event pairs
// that is drained lazily so the stream is never synchronised inside the hot loop.
struct FamilyTimer {
    static constexpr int kRing = 64;
    cudaEvent_t a[kRing], b[kRing];
    int used = 0;
    cudaStream_t s = nullptr;
    double total_ms = 0;
    u64 launches = 0, bytes = 0;
    bool inited = false;
    void init(cudaStream_t stream) {
        s = stream;
        for (int i = 0; i < kRing; ++i) { CUDA_CHECK(cudaEventCreate(&a[i])); CUDA_CHECK(cudaEventCreate(&b[i])); }
        inited = true;
    }
    void destroy() {
        if (!inited) return;
        for (int i = 0; i < kRing; ++i) { cudaEventDestroy(a[i]); cudaEventDestroy(b[i]); }
        inited = false;
    }
    void drain() {
        for (int i = 0; i < used; ++i) {
            CUDA_CHECK(cudaEventSynchronize(b[i]));
            float t = 0;
            CUDA_CHECK(cudaEventElapsedTime(&t, a[i], b[i]));
            total_ms += t;
        }
        used = 0;
    }
    void begin() {
        if (used == kRing) drain();
        CUDA_CHECK(cudaEventRecord(a[used], s));
    }
    void end(u64 n_launches, u64 alg_bytes) {
        CUDA_CHECK(cudaEventRecord(b[used], s));
        ++used;
        launches += n_launches;
        bytes += alg_bytes;
    }
    void reset() { drain(); total_ms = 0; launches = 0; bytes = 0; }
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
#endif

}  // namespace ab200
