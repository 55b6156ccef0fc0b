// tools/rs_bench.cu — developer microbenchmark for the radix-sort scatter/histogram kernels (not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --extended-lambda -lineinfo -o tools/rs_bench tools/rs_bench.cu
// Times one pass (hist + scan + scatter) over n random (u64 key, u32 value) pairs for several kernel variants and
// checks each against a host-side stable counting sort of the same digit.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../asgart_b200/csrc/common.cuh"
#include "../asgart_b200/csrc/scan.cuh"

namespace ab200 { thread_local LaunchCounter* g_launch_counter = nullptr; HostStalls g_host_stalls; thread_local DevicePool* g_device_pool = nullptr; }
using namespace ab200;

__device__ __forceinline__ u32 digit_of(u64 k, int shift) { return u32(k >> shift) & 255u; }

__device__ __forceinline__ unsigned match_ballot(u32 d, bool valid) {
    unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const unsigned vote = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? vote : ~vote;
    }
    return peers;
}

// VAR bit0: 1 = ballot match, 0 = __match_any_sync; bit1: 1 = leader atomicAdd + deferred shuffles, 0 = LDS/STS + shfl per round
// bit2: 1 = values staged through shared memory with cp.async at kernel start, 0 = values loaded into registers up front
template <int THREADS, int ITEMS, int VAR, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) scatter_variant(const u64* __restrict__ kin, const u32* __restrict__ vin, u64* __restrict__ kout,
                                                                 u32* __restrict__ vout, u64 n, int shift, const u32* __restrict__ offs,
                                                                 u64 num_tiles) {
    constexpr int TILE = THREADS * ITEMS;
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* stage = reinterpret_cast<u64*>(smem_raw);
    u32* whist = reinterpret_cast<u32*>(stage + TILE);          // [WARPS][256]
    u32* tile_off = whist + WARPS * 256;                        // [256]
    u32* delta = tile_off + 256;                                // [256]
    u32* scan_smem = delta + 256;                               // [32]
    u32* vbuf = scan_smem + 32;                                 // [TILE] (VAR bit2)

    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const u64 base = u64(blockIdx.x) * TILE;
    const u32 cnt = u32(min(u64(TILE), n - base));
    const bool full = cnt == TILE;

    if (VAR & 4) {
        if (full) {
            for (u32 c = tid; c < TILE / 4; c += THREADS) {
                const u32 dst = u32(__cvta_generic_to_shared(vbuf + c * 4));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(vin + base + c * 4));
            }
            asm volatile("cp.async.commit_group;");
        } else {
            for (u32 i = tid; i < cnt; i += THREADS) vbuf[i] = vin[base + i];
        }
    }
    for (u32 i = tid; i < WARPS * 256; i += THREADS) whist[i] = 0;

    u64 keys[ITEMS];
    u32 vals[ITEMS];
    u32 rnk[ITEMS];
    const u32 wbase = warp * (32 * ITEMS);
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const u32 li = wbase + r * 32 + lane;
        if (li < cnt) {
            keys[r] = kin[base + li];
            if (!(VAR & 4)) vals[r] = vin[base + li];
        }
    }
    __syncthreads();

    if (VAR & 2) {
        unsigned peers_a[ITEMS];
        u32 old_a[ITEMS];
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const u32 li = wbase + r * 32 + lane;
            const bool valid = li < cnt;
            const u32 d = valid ? digit_of(keys[r], shift) : 0u;
            unsigned peers;
            if (VAR & 1) peers = match_ballot(d, valid) | (valid ? 0u : (1u << lane));
            else peers = __match_any_sync(0xffffffffu, valid ? d : (256u + lane));
            peers_a[r] = peers;
            old_a[r] = 0;
            if (valid && (peers & lt) == 0) old_a[r] = atomicAdd(&whist[warp * 256 + d], u32(__popc(peers)));
            __syncwarp();
        }
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const int leader = __ffs(peers_a[r]) - 1;
            rnk[r] = __shfl_sync(0xffffffffu, old_a[r], leader) + u32(__popc(peers_a[r] & lt));
        }
    } else {
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const u32 li = wbase + r * 32 + lane;
            const bool valid = li < cnt;
            const u32 d = valid ? digit_of(keys[r], shift) : 0u;
            unsigned peers;
            if (VAR & 1) peers = match_ballot(d, valid) | (valid ? 0u : (1u << lane));
            else peers = __match_any_sync(0xffffffffu, valid ? d : (256u + lane));
            const int leader = __ffs(peers) - 1;
            u32 old = 0;
            if (valid && int(lane) == leader) {
                old = whist[warp * 256 + d];
                whist[warp * 256 + d] = old + u32(__popc(peers));
            }
            old = __shfl_sync(0xffffffffu, old, leader);
            rnk[r] = old + u32(__popc(peers & lt));
            __syncwarp();
        }
    }
    __syncthreads();

    for (u32 d = tid; d < 256; d += THREADS) {  // THREADS >= 256
        u32 sum = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            u32 t = whist[w * 256 + d];
            whist[w * 256 + d] = sum;
            sum += t;
        }
        tile_off[d] = sum;  // count for now
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of 256 counts by one warp: 8 per lane
        u32 c[8], s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { c[j] = tile_off[lane * 8 + j]; s += c[j]; }
        u32 inc = s;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, inc, dd); if (lane >= u32(dd)) inc += o; }
        u32 run = inc - s;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const u32 d = lane * 8 + j;
            tile_off[d] = run;
            delta[d] = offs[u64(d) * num_tiles + blockIdx.x] - run;
            run += c[j];
        }
    }
    __syncthreads();

#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const u32 li = wbase + r * 32 + lane;
        if (li < cnt) {
            const u32 d = digit_of(keys[r], shift);
            const u32 lp = tile_off[d] + whist[warp * 256 + d] + rnk[r];
            rnk[r] = lp;
            stage[lp] = keys[r];
        }
    }
    __syncthreads();

    u32 gp[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const u32 lp = j * THREADS + tid;
        if (lp < cnt) {
            const u64 k = stage[lp];
            gp[j] = delta[digit_of(k, shift)] + lp;
            kout[gp[j]] = k;
        }
    }
    if (VAR & 4) asm volatile("cp.async.wait_group 0;");
    __syncthreads();

    u32* vstage = reinterpret_cast<u32*>(stage);
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const u32 li = wbase + r * 32 + lane;
        if (li < cnt) vstage[rnk[r]] = (VAR & 4) ? vbuf[li] : vals[r];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const u32 lp = j * THREADS + tid;
        if (lp < cnt) vout[gp[j]] = vstage[lp];
    }
}

// HV 0: match_any + leader atomic; 1: ballot match + leader atomic; 2: plain shared atomicAdd per key
template <int THREADS, int ITEMS, int HV>
__global__ void __launch_bounds__(THREADS) hist_variant(const u64* __restrict__ kin, u64 n, int shift, u32* __restrict__ hist, u64 num_tiles) {
    constexpr int TILE = THREADS * ITEMS;
    __shared__ u32 h[256];
    for (u32 i = threadIdx.x; i < 256; i += THREADS) h[i] = 0;
    __syncthreads();
    const u64 base = u64(blockIdx.x) * TILE;
    const u32 cnt = u32(min(u64(TILE), n - base));
    u64 keys[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const u32 li = j * THREADS + threadIdx.x;
        if (li < cnt) keys[j] = kin[base + li];
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const u32 li = j * THREADS + threadIdx.x;
        const bool valid = li < cnt;
        const u32 d = valid ? digit_of(keys[j], shift) : 0u;
        if (HV == 2) {
            if (valid) atomicAdd(&h[d], 1u);
        } else {
            unsigned peers = HV == 1 ? match_ballot(d, valid) : __match_any_sync(0xffffffffu, valid ? d : 256u + (threadIdx.x & 31));
            if (valid && (__ffs(peers) - 1) == int(threadIdx.x & 31)) atomicAdd(&h[d], u32(__popc(peers)));
        }
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < 256; i += THREADS) hist[u64(i) * num_tiles + blockIdx.x] = h[i];
}

template <int THREADS, int ITEMS, int VAR, int MINB, int HV>
void run_variant(const char* name, const u64* d_k, const u32* d_v, u64* d_ko, u32* d_vo, u64 n, int shift, const std::vector<u64>& want_k,
                 const std::vector<u32>& want_v) {
    constexpr int TILE = THREADS * ITEMS;
    constexpr int WARPS = THREADS / 32;
    const u64 tiles = ceil_div(n, u64(TILE));
    const u64 table = tiles * 256;
    cudaStream_t s = nullptr;
    DevBuf<u32> hist(table, s), offs(table, s);
    size_t smem = size_t(TILE) * 8 + (WARPS * 256 + 256 + 256 + 32) * 4 + ((VAR & 4) ? size_t(TILE) * 4 : 0);
    auto kern = scatter_variant<THREADS, ITEMS, VAR, MINB>;
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    cudaEvent_t e0, e1, e2, e3;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    float best_h = 1e9f, best_s = 1e9f, best_c = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        hist_variant<THREADS, ITEMS, HV><<<unsigned(tiles), THREADS>>>(d_k, n, shift, hist.p, tiles);
        cudaEventRecord(e1);
        const u32* hp = hist.p;
        u32* op = offs.p;
        device_scan<u32, SumOp>([hp] __device__(u64 i) { return hp[i]; }, [op] __device__(u64 i, u32 exc, u32) { op[i] = exc; }, table,
                                (u32*)nullptr, s);
        cudaEventRecord(e2);
        kern<<<unsigned(tiles), THREADS, smem>>>(d_k, d_v, d_ko, d_vo, n, shift, offs.p, tiles);
        cudaEventRecord(e3);
        CUDA_CHECK(cudaEventSynchronize(e3));
        KERNEL_CHECK();
        float th, tc, ts;
        cudaEventElapsedTime(&th, e0, e1); cudaEventElapsedTime(&tc, e1, e2); cudaEventElapsedTime(&ts, e2, e3);
        best_h = std::min(best_h, th); best_c = std::min(best_c, tc); best_s = std::min(best_s, ts);
    }
    std::vector<u64> got_k(n);
    std::vector<u32> got_v(n);
    CUDA_CHECK(cudaMemcpy(got_k.data(), d_ko, n * 8, cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(got_v.data(), d_vo, n * 4, cudaMemcpyDeviceToHost));
    const bool ok = got_k == want_k && got_v == want_v;
    const double gb = double(n) * 24 / 1e9;
    printf("%-34s thr=%d items=%d occ=%d smem=%zuK  hist %.1f us (%.0f GB/s)  scan %.1f us  scatter %.1f us (%.0f GB/s alg)  %s\n", name, THREADS, ITEMS,
           occ, smem / 1024, best_h * 1e3, double(n) * 8 / 1e9 / (best_h * 1e-3), best_c * 1e3, best_s * 1e3, gb / (best_s * 1e-3), ok ? "OK" : "MISMATCH");
    fflush(stdout);
}

int main(int argc, char** argv) {
    const u64 n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 57227416ull;
    const int shift = 16;
    std::vector<u64> hk(n);
    std::vector<u32> hv(n);
    std::mt19937_64 rng(1);
    for (u64 i = 0; i < n; ++i) { hk[i] = rng() >> 1; hv[i] = u32(i); }
    std::vector<u64> want_k(n);
    std::vector<u32> want_v(n);
    {
        std::vector<u64> c(257, 0);
        for (u64 i = 0; i < n; ++i) c[((hk[i] >> shift) & 255) + 1]++;
        for (int d = 0; d < 256; ++d) c[d + 1] += c[d];
        for (u64 i = 0; i < n; ++i) { u64 p = c[(hk[i] >> shift) & 255]++; want_k[p] = hk[i]; want_v[p] = hv[i]; }
    }
    u64 *d_k, *d_ko;
    u32 *d_v, *d_vo;
    CUDA_CHECK(cudaMalloc(&d_k, n * 8)); CUDA_CHECK(cudaMalloc(&d_ko, n * 8));
    CUDA_CHECK(cudaMalloc(&d_v, n * 4)); CUDA_CHECK(cudaMalloc(&d_vo, n * 4));
    CUDA_CHECK(cudaMemcpy(d_k, hk.data(), n * 8, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(d_v, hv.data(), n * 4, cudaMemcpyHostToDevice));
    printf("n = %llu pairs (u64 key, u32 value); algorithmic bytes per scatter pass = %.3f GB\n", (unsigned long long)n, double(n) * 24 / 1e9);
#define RUN(T, I, V, M, H) run_variant<T, I, V, M, H>("T" #T " I" #I " VAR" #V " MINB" #M " HV" #H, d_k, d_v, d_ko, d_vo, n, shift, want_k, want_v)
    RUN(256, 16, 0, 2, 0);
    RUN(256, 16, 1, 2, 1);
    RUN(256, 16, 2, 2, 2);
    RUN(256, 16, 3, 2, 1);
    RUN(256, 16, 4, 3, 1);
    RUN(256, 16, 6, 3, 1);
    RUN(256, 16, 7, 3, 1);
    RUN(512, 8, 0, 2, 1);
    RUN(512, 8, 2, 2, 1);
    RUN(512, 8, 4, 2, 1);
    RUN(512, 8, 6, 2, 1);
    RUN(512, 8, 7, 2, 1);
    RUN(256, 8, 4, 6, 1);
    RUN(256, 8, 6, 6, 1);
    RUN(256, 8, 7, 6, 1);
    RUN(1024, 4, 6, 1, 1);
    RUN(1024, 4, 4, 1, 1);
    RUN(512, 16, 6, 1, 1);
    RUN(384, 12, 6, 2, 1);
    return 0;
}
