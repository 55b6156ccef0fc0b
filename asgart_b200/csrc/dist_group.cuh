// dist_group.cuh — the two back ends of SaGroup (sa_group.h).
//   ThreadGroup  members are host threads of one process (asgart_b200_build_index_group): host-side barrier, peers'
//                buffers addressed directly (same device, or peer access between devices).
//   NcclGroup    one process per GPU (asgart_b200_ctx_dist_init): NCCL collectives on the context's stream, peers'
//                buffers mapped with CUDA IPC. libnccl.so.2 is resolved at run time (dlopen) so that single-GPU users
//                do not need it; inside a torch process this is the NCCL torch itself loaded.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "sa_group.h"

namespace ab200 {

template <typename T>
__global__ void max_into_kernel(T* __restrict__ mine, const T* __restrict__ other, u64 n) {
    const u64 stride = u64(gridDim.x) * blockDim.x;
    for (u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T o = other[i];
        if (o > mine[i]) mine[i] = o;
    }
}

// ---------------------------------------------------------------------------------------------------- threads
struct ThreadGroupShared {
    int world;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    u64 generation = 0;
    bool failed = false;
    std::vector<std::vector<u64>> slots;
    std::vector<void*> ptrs;
    explicit ThreadGroupShared(int w) : world(w), slots(w), ptrs(w, nullptr) {}
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        if (failed) throw std::runtime_error("a member of the index group failed");
        const u64 gen = generation;
        if (++arrived == world) { arrived = 0; ++generation; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != gen || failed; });
        if (failed) throw std::runtime_error("a member of the index group failed");
    }
    void fail() {   // a member is leaving with an error: release the others
        std::lock_guard<std::mutex> lk(m);
        failed = true;
        cv.notify_all();
    }
};

struct ThreadGroup : SaGroup {
    ThreadGroupShared* sh;
    cudaStream_t stream;
    ThreadGroup(ThreadGroupShared* s, int r, cudaStream_t st) : sh(s), stream(st) { rank = r; world = s->world; }
    void allreduce_sum_host(u64* vals, int n) override {
        sh->slots[rank].assign(vals, vals + n);
        sh->barrier();
        for (int i = 0; i < n; ++i) {
            u64 t = 0;
            for (int r = 0; r < world; ++r) t += sh->slots[r][i];
            vals[i] = t;
        }
        sh->barrier();
    }
    void exchange_ptr(void* mine, size_t, void** all) override {
        sh->ptrs[rank] = mine;
        sh->barrier();
        for (int r = 0; r < world; ++r) all[r] = sh->ptrs[r];
        sh->barrier();
    }
    // max is idempotent and the arrays only grow towards the result, so members may read each other's arrays while those
    // are being updated: whatever version is read lies between the peer's input and the global maximum
    void allreduce_max_dev(void* buf, size_t count, int elem_bytes, cudaStream_t st) override {
        void* all[kMaxWorld];
        CUDA_CHECK(cudaStreamSynchronize(st));
        exchange_ptr(buf, count * elem_bytes, all);
        const unsigned grid = unsigned(std::min<u64>(ceil_div(std::max<u64>(count, 1), 256), u64(kNumSMs) * 8));
        for (int r = 0; r < world && count; ++r) {
            if (r == rank) continue;
            if (elem_bytes == 4) max_into_kernel<u32><<<grid, 256, 0, st>>>(static_cast<u32*>(buf), static_cast<const u32*>(all[r]), count);
            else max_into_kernel<u64><<<grid, 256, 0, st>>>(static_cast<u64*>(buf), static_cast<const u64*>(all[r]), count);
            KERNEL_CHECK();
            count_launch();
        }
        CUDA_CHECK(cudaStreamSynchronize(st));
        sh->barrier();
    }
    void share_pieces(void* base, const u64* offs, int elem_bytes, cudaStream_t st) override {
        void* all[kMaxWorld];
        CUDA_CHECK(cudaStreamSynchronize(st));
        exchange_ptr(base, 0, all);
        for (int r = 0; r < world; ++r) {
            if (r == rank || offs[r + 1] == offs[r]) continue;
            const size_t o = size_t(offs[r]) * elem_bytes, bytes = size_t(offs[r + 1] - offs[r]) * elem_bytes;
            CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(base) + o, static_cast<const char*>(all[r]) + o, bytes, cudaMemcpyDefault, st));
        }
        CUDA_CHECK(cudaStreamSynchronize(st));
        sh->barrier();
    }
};

// ---------------------------------------------------------------------------------------------------- NCCL + IPC
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool load() {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { error = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
        auto sym = [&](const char* name) -> void* {
            void* p = dlsym(lib, name);
            if (!p) error = std::string("libnccl lacks ") + name;
            return p;
        };
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
        Broadcast = reinterpret_cast<decltype(Broadcast)>(sym("ncclBroadcast"));
        AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        if (!error.empty()) { lib = nullptr; return false; }
        return true;
    }
};
inline NcclApi& nccl_api() {
    static NcclApi api;
    return api;
}
inline void nccl_check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) {
        const char* msg = nccl_api().GetErrorString ? nccl_api().GetErrorString(r) : "?";
        throw CudaError(ASGART_B200_ECUDA, std::string(what) + ": NCCL error: " + msg);
    }
}
#define NCCL_CHECK(x) ::ab200::nccl_check((x), #x)

struct NcclGroup : SaGroup {
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    u64* d_stage = nullptr;   // 64 words: host-value reductions, IPC handles
    u64* h_stage = nullptr;   // pinned
    std::map<std::string, void*> opened;   // IPC handle bytes -> mapped base
    static constexpr int kStageWords = 256 * kMaxWorld + 64;   // host-value reductions up to one 256-bin table per member

    NcclGroup(int r, int w, const ncclUniqueId& id, cudaStream_t st) : stream(st) {
        rank = r; world = w;
        NCCL_CHECK(nccl_api().CommInitRank(&comm, w, id, r));
        CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_stage), sizeof(u64) * kStageWords * (kMaxWorld + 1)));
        CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&h_stage), sizeof(u64) * kStageWords * (kMaxWorld + 1)));
    }
    ~NcclGroup() override {
        for (auto& kv : opened) cudaIpcCloseMemHandle(kv.second);
        if (d_stage) cudaFree(d_stage);
        if (h_stage) cudaFreeHost(h_stage);
        if (comm) nccl_api().CommDestroy(comm);
    }
    void allreduce_sum_host(u64* vals, int n) override {
        if (n > kStageWords) throw std::runtime_error("allreduce_sum_host: too many values");
        memcpy(h_stage, vals, sizeof(u64) * n);
        CUDA_CHECK(cudaMemcpyAsync(d_stage, h_stage, sizeof(u64) * n, cudaMemcpyHostToDevice, stream));
        NCCL_CHECK(nccl_api().AllReduce(d_stage, d_stage, size_t(n), ncclUint64, ncclSum, comm, stream));
        CUDA_CHECK(cudaMemcpyAsync(h_stage, d_stage, sizeof(u64) * n, cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        memcpy(vals, h_stage, sizeof(u64) * n);
    }
    void allreduce_max_dev(void* buf, size_t count, int elem_bytes, cudaStream_t st) override {
        if (count == 0) return;
        NCCL_CHECK(nccl_api().AllReduce(buf, buf, count, elem_bytes == 4 ? ncclUint32 : ncclUint64, ncclMax, comm, st));
    }
    // Every member pulls the other members' pieces straight out of their copies of the array (CUDA IPC mappings, copy
    // engines over NVLink). `base` must be the start of a cudaMalloc block. Measured at two GPUs, 6.2 GB per direction:
    // 26 ms as grouped ncclBroadcasts of the pieces.
    void share_pieces(void* base, const u64* offs, int elem_bytes, cudaStream_t st) override {
        void* all[kMaxWorld] = {};
        exchange_ptr(base, 0, all);     // stream-ordered collective inside: the peers' pieces are complete when it returns
        for (int i = 1; i < world; ++i) {
            const int r = (rank + i) % world;    // staggered, so that the members do not all read from the same peer at once
            const size_t cnt = size_t(offs[r + 1] - offs[r]);
            if (cnt == 0) continue;
            const size_t o = size_t(offs[r]) * elem_bytes;
            CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(base) + o, static_cast<const char*>(all[r]) + o, cnt * elem_bytes, cudaMemcpyDefault, st));
        }
        CUDA_CHECK(cudaStreamSynchronize(st));
        barrier();   // nobody touches its copy again before everyone has read its piece
    }
    void exchange_ptr(void* mine, size_t, void** all) override {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        CUDA_CHECK(cudaIpcGetMemHandle(&h, mine));
        constexpr int W8 = 8;   // 64 bytes = 8 words per member
        memcpy(h_stage, &h, sizeof h);
        u64* d_all = d_stage + kStageWords;
        u64* h_all = h_stage + kStageWords;
        CUDA_CHECK(cudaMemcpyAsync(d_stage, h_stage, sizeof h, cudaMemcpyHostToDevice, stream));
        NCCL_CHECK(nccl_api().AllGather(d_stage, d_all, W8, ncclUint64, comm, stream));
        CUDA_CHECK(cudaMemcpyAsync(h_all, d_all, sizeof(u64) * W8 * world, cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        for (int r = 0; r < world; ++r) {
            if (r == rank) { all[r] = mine; continue; }
            const std::string key(reinterpret_cast<const char*>(h_all + size_t(r) * W8), sizeof h);
            auto it = opened.find(key);
            if (it == opened.end()) {
                cudaIpcMemHandle_t ph;
                memcpy(&ph, key.data(), sizeof ph);
                void* p = nullptr;
                CUDA_CHECK(cudaIpcOpenMemHandle(&p, ph, cudaIpcMemLazyEnablePeerAccess));
                it = opened.emplace(key, p).first;
            }
            all[r] = it->second;
        }
    }
};

}  // namespace ab200
