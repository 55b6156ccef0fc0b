// tools/scatter_bench.cu — developer microbenchmark: cost of the inverse-permutation scatter rank[SA[i]] = f(i)
// that ends every re-ranking step of the suffix-array build (not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/scatter_bench tools/scatter_bench.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

typedef uint32_t u32;
typedef uint64_t u64;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void scatter_plain(const u32* __restrict__ perm, u32* __restrict__ out, u64 n) {
    const u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[perm[i]] = u32(i);
}
__global__ void scatter_cs(const u32* __restrict__ perm, u32* __restrict__ out, u64 n) {
    const u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) __stcs(out + perm[i], u32(i));
}
__global__ void scatter_cg(const u32* __restrict__ perm, u32* __restrict__ out, u64 n) {
    const u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) __stcg(out + perm[i], u32(i));
}
// only the targets inside [lo, hi): K sweeps keep each sweep's targets inside an L2-sized slice
__global__ void scatter_range(const u32* __restrict__ perm, u32* __restrict__ out, u64 n, u32 lo, u32 hi) {
    const u64 stride = u64(gridDim.x) * blockDim.x * 4;
    for (u64 i = (u64(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            const uint4 p = *reinterpret_cast<const uint4*>(perm + i);
            if (p.x >= lo && p.x < hi) out[p.x] = u32(i);
            if (p.y >= lo && p.y < hi) out[p.y] = u32(i + 1);
            if (p.z >= lo && p.z < hi) out[p.z] = u32(i + 2);
            if (p.w >= lo && p.w < hi) out[p.w] = u32(i + 3);
        } else {
            for (u64 j = i; j < n; ++j) { const u32 p = perm[j]; if (p >= lo && p < hi) out[p] = u32(j); }
        }
    }
}
// streaming loads (evict-first) so that the L2 keeps the lines being written
__global__ void scatter_range_cs(const u32* __restrict__ perm, u32* __restrict__ out, u64 n, u32 lo, u32 hi) {
    const u64 stride = u64(gridDim.x) * blockDim.x * 4;
    for (u64 i = (u64(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            const uint4 p = __ldcs(reinterpret_cast<const uint4*>(perm + i));
            if (p.x >= lo && p.x < hi) out[p.x] = u32(i);
            if (p.y >= lo && p.y < hi) out[p.y] = u32(i + 1);
            if (p.z >= lo && p.z < hi) out[p.z] = u32(i + 2);
            if (p.w >= lo && p.w < hi) out[p.w] = u32(i + 3);
        } else {
            for (u64 j = i; j < n; ++j) { const u32 p = perm[j]; if (p >= lo && p < hi) out[p] = u32(j); }
        }
    }
}
__global__ void gather_plain(const u32* __restrict__ inv, const u32* __restrict__ val, u32* __restrict__ out, u64 n) {
    const u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = val[inv[i]];
}

int main(int argc, char** argv) {
    const u64 n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 57227416ull;
    std::vector<u32> perm(n);
    std::iota(perm.begin(), perm.end(), 0u);
    std::mt19937_64 rng(7);
    std::shuffle(perm.begin(), perm.end(), rng);
    u32 *d_perm, *d_out, *d_val;
    CK(cudaMalloc(&d_perm, n * 4)); CK(cudaMalloc(&d_out, n * 4)); CK(cudaMalloc(&d_val, n * 4));
    CK(cudaMemcpy(d_perm, perm.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_val, 1, n * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned blocks = unsigned((n + 255) / 256);
    auto time = [&](const char* name, auto fn) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaMemset(d_out, 0xFF, n * 4));
            cudaEventRecord(e0);
            fn();
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float t; cudaEventElapsedTime(&t, e0, e1);
            best = std::min(best, t);
        }
        std::vector<u32> h(n);
        CK(cudaMemcpy(h.data(), d_out, n * 4, cudaMemcpyDeviceToHost));
        bool ok = true;
        if (name[0] != 'g') for (u64 i = 0; i < n && ok; i += 997) ok = h[perm[i]] == u32(i);
        printf("%-34s %8.1f us  (%.1f G elem/s)  %s\n", name, best * 1e3, double(n) / best / 1e6, ok ? "OK" : "MISMATCH");
        fflush(stdout);
    };
    printf("n = %llu random 4-byte targets (%.0f MB array)\n", (unsigned long long)n, double(n) * 4 / 1e6);
    time("scatter plain", [&] { scatter_plain<<<blocks, 256>>>(d_perm, d_out, n); });
    time("scatter __stcs", [&] { scatter_cs<<<blocks, 256>>>(d_perm, d_out, n); });
    time("scatter __stcg", [&] { scatter_cg<<<blocks, 256>>>(d_perm, d_out, n); });
    for (int K : {2, 4, 8, 16}) {
        char nm[64];
        snprintf(nm, sizeof nm, "scatter %d range sweeps", K);
        time(nm, [&] {
            for (int k = 0; k < K; ++k) {
                const u64 lo = n * k / K, hi = n * (k + 1) / K;
                scatter_range<<<148 * 16, 256>>>(d_perm, d_out, n, u32(lo), u32(hi));
            }
        });
    }
    for (int K : {4, 8}) {
        char nm[64];
        snprintf(nm, sizeof nm, "scatter %d range sweeps, ld.cs", K);
        time(nm, [&] {
            for (int k = 0; k < K; ++k) {
                const u64 lo = n * k / K, hi = n * (k + 1) / K;
                scatter_range_cs<<<148 * 16, 256>>>(d_perm, d_out, n, u32(lo), u32(hi));
            }
        });
    }
    // pre-partitioned input: elements grouped by target slice (random inside a slice), one in-order scatter
    for (u64 slice_mb : {16, 32, 48, 64, 96}) {
        const u64 slice = slice_mb * 1024 * 1024 / 4;
        std::vector<u32> part(perm);
        std::stable_sort(part.begin(), part.end(), [&](u32 a, u32 b) { return a / slice < b / slice; });
        u32* d_part;
        CK(cudaMalloc(&d_part, n * 4));
        CK(cudaMemcpy(d_part, part.data(), n * 4, cudaMemcpyHostToDevice));
        char nm[64];
        snprintf(nm, sizeof nm, "g in-order, %llu MB slices", (unsigned long long)slice_mb);
        time(nm, [&] { scatter_range<<<148 * 16, 256>>>(d_part, d_out, n, 0u, u32(n)); });
        snprintf(nm, sizeof nm, "g in-order, %llu MB slices, ld.cs", (unsigned long long)slice_mb);
        time(nm, [&] { scatter_range_cs<<<148 * 16, 256>>>(d_part, d_out, n, 0u, u32(n)); });
        snprintf(nm, sizeof nm, "g in-order, %llu MB, 148x4 blocks", (unsigned long long)slice_mb);
        time(nm, [&] { scatter_range_cs<<<148 * 4, 256>>>(d_part, d_out, n, 0u, u32(n)); });
        cudaFree(d_part);
    }
    time("gather plain (out[i]=val[inv[i]])", [&] { gather_plain<<<blocks, 256>>>(d_perm, d_val, d_out, n); });
    return 0;
}
