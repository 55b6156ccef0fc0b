// tools/msd_bench.cu — developer check + microbenchmark of msd_sort.cuh against the LSD path (not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --extended-lambda -lineinfo -o tools/msd_bench tools/msd_bench.cu
//   tools/msd_bench [n] [p0] [mode]      mode: 0 random DNA, 1 + N-runs and planted repeats, 2 low complexity (worst case)
// Checks: output keys non-decreasing, key(vals[i]) == keys[i] (recomputed from the text), vals a permutation of [0, n),
// sorted key array identical to the LSD sort's. Prints the time of each kernel family.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../asgart_b200/csrc/sa_build.cuh"
#include "../asgart_b200/csrc/msd_sort.cuh"

namespace ab200 { thread_local LaunchCounter* g_launch_counter = nullptr; thread_local HostStalls g_host_stalls; thread_local DevicePool* g_device_pool = nullptr; }
using namespace ab200;

__global__ void check_kernel(const u8* text, u64 n, const uint16_t* code, int b, int p0, const u64* keys, const u32* vals, u64 m,
                             unsigned long long* bad, u32* bitmap) {
    const u64 stride = u64(gridDim.x) * blockDim.x;
    const u64 mask = (b * p0 >= 64) ? ~u64(0) : ((u64(1) << (b * p0)) - 1);
    for (u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x; i < m; i += stride) {
        const u64 k = keys[i];
        if (i > 0 && keys[i - 1] > k) atomicAdd(&bad[0], 1ull);
        const u64 v = vals[i];
        if (v >= n) { atomicAdd(&bad[1], 1ull); continue; }
        u64 kk = 0;
        for (int j = 0; j < p0; ++j) kk = (kk << b) | (v + j < n ? u64(code[text[v + j]]) : 0);
        if ((kk & mask) != k) atomicAdd(&bad[1], 1ull);
        const u32 old = atomicOr(&bitmap[v >> 5], 1u << (v & 31));
        if (old & (1u << (v & 31))) atomicAdd(&bad[2], 1ull);
    }
}
__global__ void diff_kernel(const u64* a, const u64* b, u64 m, unsigned long long* bad) {
    const u64 stride = u64(gridDim.x) * blockDim.x;
    for (u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x; i < m; i += stride)
        if (a[i] != b[i]) atomicAdd(&bad[3], 1ull);
}

int main(int argc, char** argv) {
    const u64 n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 50'000'000ull;
    const int p0 = argc > 2 ? atoi(argv[2]) : 18;
    const int mode = argc > 3 ? atoi(argv[3]) : 1;
    const int b = 3;
    std::vector<u8> h(n);
    {
        std::mt19937_64 rng(12345 + mode);
        const char* acgt = "ACGT";
        for (u64 i = 0; i + 1 < n; ++i) h[i] = u8(acgt[rng() & 3]);
        if (mode >= 1) {
            for (int r = 0; r < 40; ++r) {   // planted exact repeats
                const u64 len = 2000 + rng() % 50000, src = rng() % (n - len - 1), dst = rng() % (n - len - 1);
                for (u64 j = 0; j < len; ++j) h[dst + j] = h[src + j];
            }
            const u64 runs[] = {n / 20, 10000, 10000, 5001, 300, 77};
            for (u64 len : runs) { const u64 at = rng() % (n - len - 1); for (u64 j = 0; j < len; ++j) h[at + j] = 'N'; }
            for (int r = 0; r < 2000; ++r) { const u64 at = rng() % (n - 400), len = 1 + rng() % 300; for (u64 j = 0; j < len; ++j) h[at + j] = 'N'; }
            { const u64 at = rng() % (n - 200000); for (u64 j = 0; j < 150000; ++j) h[at + j] = u8("AC"[j & 1]); }   // microsatellite
            { const u64 at = rng() % (n - 200000); for (u64 j = 0; j < 90000; ++j) h[at + j] = 'A'; }
        }
        if (mode == 2) for (u64 i = 0; i + 1 < n; ++i) h[i] = u8((i / 1000) % 7 == 0 ? 'C' : 'A');
        h[n - 1] = '$';
    }
    uint16_t h_code[256] = {};
    {
        bool seen[256] = {};
        for (u64 i = 0; i < n; ++i) seen[h[i]] = true;
        int s = 0;
        for (int c = 0; c < 256; ++c) if (seen[c]) h_code[c] = uint16_t(++s);
        printf("n = %llu, p0 = %d, b = %d, sigma = %d, mode %d\n", (unsigned long long)n, p0, b, s, mode);
    }
    cudaStream_t stream;
    CUDA_CHECK(cudaStreamCreate(&stream));
    DevBuf<u8> d_text(n + 64, stream);
    d_text.zero();
    CUDA_CHECK(cudaMemcpyAsync(d_text.p, h.data(), n, cudaMemcpyHostToDevice, stream));
    DevBuf<uint16_t> d_code(256, stream);
    CUDA_CHECK(cudaMemcpyAsync(d_code.p, h_code, sizeof h_code, cudaMemcpyHostToDevice, stream));
    DevBuf<u64> ka(n + 2, stream), kb(n + 2, stream), kref(n + 2, stream), kref2(n + 2, stream);
    DevBuf<u32> va(n, stream), vb(n, stream), vref(n, stream), vref2(n, stream);
    DevBuf<unsigned long long> bad(4, stream);
    DevBuf<u32> bitmap(n / 32 + 1, stream);

    FamilyTimer t_sc, t_sc0, t_hist, t_loc;
    t_sc.init(stream); t_sc0.init(stream); t_hist.init(stream); t_loc.init(stream);
    MsdStats ms;
    ms.scatter = &t_sc; ms.scatter0 = &t_sc0; ms.hist = &t_hist; ms.local = &t_loc;
    EventTimer tm(stream);
    double best = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
        if (rep == 1) { t_sc.reset(); t_sc0.reset(); t_hist.reset(); t_loc.reset(); }
        tm.start();
        const bool okk = msd_sort_suffixes<u32>(d_text.p, n, d_code.p, b, p0, 0, kMsdBins, n, ka.p, va.p, kb.p, vb.p, nullptr, stream, &ms);
        tm.stop();
        const double t = tm.ms();
        if (!okk) { printf("msd sort declined (table budget)\n"); return 2; }
        if (rep) best = std::min(best, t);
        printf("msd rep %d: %.3f ms (%d levels)\n", rep, t, ms.levels);
    }
    t_sc.drain(); t_sc0.drain(); t_hist.drain(); t_loc.drain();
    printf("per rep: scatter0 %.3f ms (%.0f GB/s)  scatter %.3f ms (%.0f GB/s, %llu launches)  hist %.3f ms (%.0f GB/s)  local %.3f ms (%.0f GB/s)\n",
           t_sc0.total_ms / 3, t_sc0.bytes / (t_sc0.total_ms * 1e6), t_sc.total_ms / 3, t_sc.bytes / (t_sc.total_ms * 1e6),
           (unsigned long long)t_sc.launches / 3, t_hist.total_ms / 3, t_hist.bytes / (t_hist.total_ms * 1e6), t_loc.total_ms / 3,
           t_loc.bytes / (t_loc.total_ms * 1e6));
    printf("msd best %.3f ms = %.2f G suffixes/s\n", best, n / best / 1e6);

    // reference: key generation + LSD sort
    double best_ref = 1e30;
    u64 *k = nullptr, *k2 = nullptr;
    u32 *v = nullptr, *v2 = nullptr;
    for (int rep = 0; rep < 3; ++rep) {
        k = kref.p; k2 = kref2.p; v = vref.p; v2 = vref2.p;
        std::vector<int> shifts;
        for (int s = 0; s < b * p0; s += 8) shifts.push_back(s);
        tm.start();
        init_keys_kernel<u32><<<unsigned(ceil_div(n, u64(kInitTile))), kInitThreads, 0, stream>>>(d_text.p, n, d_code.p, b, p0, k, v);
        radix_sort_pairs<u64, u32>(k, k2, v, v2, n, shifts.data(), int(shifts.size()), stream);
        tm.stop();
        const double t = tm.ms();
        if (rep) best_ref = std::min(best_ref, t);
    }
    printf("lsd best %.3f ms = %.2f G suffixes/s   ->  msd is %.2fx\n", best_ref, n / best_ref / 1e6, best_ref / best);

    bad.zero();
    bitmap.zero();
    check_kernel<<<kNumSMs * 8, 256, 0, stream>>>(d_text.p, n, d_code.p, b, p0, ka.p, va.p, n, bad.p, bitmap.p);
    diff_kernel<<<kNumSMs * 8, 256, 0, stream>>>(ka.p, k, n, bad.p);
    unsigned long long hb[4];
    CUDA_CHECK(cudaMemcpyAsync(hb, bad.p, sizeof hb, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    printf("order violations %llu, key(val) mismatches %llu, duplicate vals %llu, keys differing from LSD %llu  -> %s\n", hb[0], hb[1], hb[2], hb[3],
           (hb[0] | hb[1] | hb[2] | hb[3]) ? "FAIL" : "OK");
    return (hb[0] | hb[1] | hb[2] | hb[3]) ? 1 : 0;
}
