// radix_sort.cuh — hand-written stable LSD radix sort of (key, value) pairs for sm_100a.
//
// One 8-bit digit per pass, three launches per pass and no inter-block spinning:
//   rs_hist_kernel     per-tile digit histogram (shared-memory atomics)                          reads keys
//   device_scan        exclusive scan of the digit-major table hist[digit][tile] -> global offset of every (digit, tile)
//   rs_scatter_kernel  stable in-tile ranking (ballot match + shared atomics), tile reordered in shared memory so that
//                      each digit's run leaves the SM as contiguous, coalesced stores          reads+writes keys, values
// Keys are u32, u64 or U128 (two u64 words; used when suffix indices need more than 32 bits), values u32 or u64.
// HBM traffic per pass and element: 2*sizeof(Key) + sizeof(Key) [histogram re-read] + 2*sizeof(Val).
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace ab200 {

struct U128 {
    u64 hi, lo;
};
__host__ __device__ __forceinline__ bool operator==(const U128& a, const U128& b) { return a.hi == b.hi && a.lo == b.lo; }
__host__ __device__ __forceinline__ bool operator!=(const U128& a, const U128& b) { return !(a == b); }

__device__ __forceinline__ u32 rs_digit(u32 k, int shift) { return (k >> shift) & 255u; }
__device__ __forceinline__ u32 rs_digit(u64 k, int shift) { return u32(k >> shift) & 255u; }
__device__ __forceinline__ u32 rs_digit(const U128& k, int shift) {
    if (shift >= 64) return u32(k.hi >> (shift - 64)) & 255u;
    u64 v = k.lo >> shift;
    if (shift > 56) v |= k.hi << (64 - shift);
    return u32(v) & 255u;
}

// lanes of the warp holding the same 8-bit digit (lanes with valid == false match nobody). One ballot per digit bit:
// __match_any_sync serialises over the distinct values in the warp, and with 256 bins almost all 32 are distinct.
__device__ __forceinline__ unsigned match_digit(u32 d, bool valid) {
    unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const unsigned vote = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? vote : ~vote;
    }
    return peers;
}

constexpr int kRsThreads = 512;
constexpr int kRsWarps = kRsThreads / 32;
// tile = 4096 keys (u64) or 2048 (U128): 16 warps of short ranking chains hide latency better than 8 warps x 16 items
// (tools/rs_bench.cu on B200: 476 us vs 683 us per pass over 57 M pairs)
template <typename KeyT> struct RsItems { static constexpr int value = sizeof(KeyT) > 8 ? 4 : 8; };

// Per-tile digit histogram. Plain shared-memory atomics: on random digits they run at HBM speed (5.5 TB/s measured),
// 2.4x faster than warp-aggregating with ballots and 4.8x faster than __match_any_sync.
template <typename KeyT, int ITEMS>
__global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(const KeyT* __restrict__ kin, u64 n, int shift,
                                                             u32* __restrict__ hist, u64 num_tiles) {
    constexpr int TILE = kRsThreads * ITEMS;
    __shared__ u32 h[256];
    if (threadIdx.x < 256) h[threadIdx.x] = 0;
    __syncthreads();
    const u64 base = u64(blockIdx.x) * TILE;
    const u32 cnt = u32(min(u64(TILE), n - base));
    KeyT keys[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const u32 li = j * kRsThreads + threadIdx.x;
        if (li < cnt) keys[j] = kin[base + li];
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const u32 li = j * kRsThreads + threadIdx.x;
        if (li < cnt) atomicAdd(&h[rs_digit(keys[j], shift)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 256) hist[u64(threadIdx.x) * num_tiles + blockIdx.x] = h[threadIdx.x];
}

template <typename KeyT, typename ValT, typename OffT, int ITEMS>
struct RsSmem {
    static constexpr int TILE = kRsThreads * ITEMS;
    static constexpr size_t bytes = size_t(TILE) * (sizeof(KeyT) + 2 * sizeof(ValT)) + (kRsWarps * 256 + 256) * sizeof(u32) + 256 * sizeof(OffT);
};

// Stable scatter of one tile by one 8-bit digit.
//   values: cp.async'ed into shared memory at kernel start (no registers, overlaps the ranking)
//   ranking: per round of 32 keys, lanes with equal digits are matched with one ballot per digit bit, the first of them
//            bumps the warp's digit counter with a shared atomicAdd; the returned bases are shuffled out afterwards
//   keys and values are reordered in shared memory so each digit's run leaves the SM as contiguous stores
template <typename KeyT, typename ValT, typename OffT, int ITEMS>
__global__ void __launch_bounds__(kRsThreads, 2) rs_scatter_kernel(const KeyT* __restrict__ kin, const ValT* __restrict__ vin,
                                                                   KeyT* __restrict__ kout, ValT* __restrict__ vout, u64 n,
                                                                   int shift, const OffT* __restrict__ offs, u64 num_tiles) {
    constexpr int TILE = kRsThreads * ITEMS;
    extern __shared__ __align__(16) unsigned char rs_smem[];
    KeyT* kstage = reinterpret_cast<KeyT*>(rs_smem);
    ValT* vstage = reinterpret_cast<ValT*>(kstage + TILE);
    ValT* vbuf = vstage + TILE;
    OffT* delta = reinterpret_cast<OffT*>(vbuf + TILE);
    u32* whist = reinterpret_cast<u32*>(delta + 256);  // [kRsWarps][256]
    u32* tile_off = whist + kRsWarps * 256;            // [256]

    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const u64 base = u64(blockIdx.x) * TILE;
    const u32 cnt = u32(min(u64(TILE), n - base));

    if (cnt == TILE) {
        constexpr int PER16 = 16 / sizeof(ValT);
        for (u32 c = tid; c < TILE / PER16; c += kRsThreads) {
            const u32 dst = u32(__cvta_generic_to_shared(vbuf + c * PER16));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(vin + base + c * PER16));
        }
    } else {
        for (u32 i = tid; i < cnt; i += kRsThreads) vbuf[i] = vin[base + i];
    }
    asm volatile("cp.async.commit_group;");
    for (u32 i = tid; i < kRsWarps * 256; i += kRsThreads) whist[i] = 0;

    KeyT keys[ITEMS];
    const u32 wbase = warp * (32 * ITEMS);
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const u32 li = wbase + r * 32 + lane;
        if (li < cnt) keys[r] = kin[base + li];
    }
    __syncthreads();

    unsigned peers_a[ITEMS];
    u32 rnk[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const u32 li = wbase + r * 32 + lane;
        const bool valid = li < cnt;
        const u32 d = valid ? rs_digit(keys[r], shift) : 0u;
        const unsigned peers = match_digit(d, valid) | (valid ? 0u : (1u << lane));
        peers_a[r] = peers;
        rnk[r] = 0;
        if (valid && (peers & lt) == 0) rnk[r] = atomicAdd(&whist[warp * 256 + d], u32(__popc(peers)));
        __syncwarp();  // rounds must reach the counters in order: stability
    }
#pragma unroll
    for (int r = 0; r < ITEMS; ++r)
        rnk[r] = __shfl_sync(0xffffffffu, rnk[r], __ffs(peers_a[r]) - 1) + u32(__popc(peers_a[r] & lt));
    __syncthreads();

    if (tid < 256) {  // exclusive prefix over warps for digit tid
        u32 sum = 0;
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w) {
            const u32 t = whist[w * 256 + tid];
            whist[w * 256 + tid] = sum;
            sum += t;
        }
        tile_off[tid] = sum;
    }
    asm volatile("cp.async.wait_group 0;");
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the 256 digit counts: 8 per lane + warp scan
        u32 c[8], ssum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { c[j] = tile_off[lane * 8 + j]; ssum += c[j]; }
        u32 inc = ssum;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { const u32 o = __shfl_up_sync(0xffffffffu, inc, dd); if (lane >= u32(dd)) inc += o; }
        u32 run = inc - ssum;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const u32 d = lane * 8 + j;
            tile_off[d] = run;
            delta[d] = offs[u64(d) * num_tiles + blockIdx.x] - OffT(run);
            run += c[j];
        }
    }
    __syncthreads();

#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const u32 li = wbase + r * 32 + lane;
        if (li < cnt) {
            const u32 d = rs_digit(keys[r], shift);
            const u32 lp = tile_off[d] + whist[warp * 256 + d] + rnk[r];
            kstage[lp] = keys[r];
            vstage[lp] = vbuf[li];
        }
    }
    __syncthreads();

#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const u32 lp = j * kRsThreads + tid;
        if (lp < cnt) {
            const KeyT k = kstage[lp];
            const OffT gp = delta[rs_digit(k, shift)] + OffT(lp);
            kout[gp] = k;
            vout[gp] = vstage[lp];
        }
    }
}

// Scratch of the radix passes over n elements: the digit-major tile histogram and its scanned offsets.
template <typename KeyT>
struct RadixScratch {
    static constexpr int ITEMS = RsItems<KeyT>::value;
    static constexpr int TILE = kRsThreads * ITEMS;
    u64 n = 0, tiles = 0, table = 0;
    bool wide = false;
    DevBuf<u32> hist, offs32;
    DevBuf<u64> offs64;
    RadixScratch(u64 n_, cudaStream_t stream) : n(n_), tiles(ceil_div(n_, u64(TILE))), table(tiles * 256), wide(n_ >= (u64(1) << 32)) {
        hist.alloc(table, stream);
        offs32.alloc(wide ? 0 : table, stream);
        offs64.alloc(wide ? table : 0, stream);
    }
};

// One stable pass by the digit at bits [shift, shift + 8): (kin, vin) -> (kout, vout).
template <typename KeyT, typename ValT>
void radix_pass(RadixScratch<KeyT>& ws, const KeyT* kin, const ValT* vin, KeyT* kout, ValT* vout, int shift, cudaStream_t stream,
                FamilyTimer* timer = nullptr, FamilyTimer* scatter_timer = nullptr) {
    constexpr int ITEMS = RsItems<KeyT>::value;
    const u64 n = ws.n, tiles = ws.tiles, table = ws.table;
    if (n == 0) return;
    constexpr size_t smem32 = RsSmem<KeyT, ValT, u32, ITEMS>::bytes, smem64 = RsSmem<KeyT, ValT, u64, ITEMS>::bytes;
    static std::atomic<unsigned long long> prepared{0};  // per (KeyT, ValT) instantiation, one bit per device
    unsigned long long dev_bit = 0;
    if (device_needs_prepare(prepared, dev_bit)) {
        CUDA_CHECK(cudaFuncSetAttribute(rs_scatter_kernel<KeyT, ValT, u32, ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem32)));
        CUDA_CHECK(cudaFuncSetAttribute(rs_scatter_kernel<KeyT, ValT, u64, ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem64)));
        device_prepared(prepared, dev_bit);
    }
    if (timer) timer->begin();
    rs_hist_kernel<KeyT, ITEMS><<<unsigned(tiles), kRsThreads, 0, stream>>>(kin, n, shift, ws.hist.p, tiles);
    KERNEL_CHECK();
    const u32* hp = ws.hist.p;
    if (!ws.wide) {
        u32* op = ws.offs32.p;
        device_scan<u32, SumOp>([hp] __device__(u64 i) { return hp[i]; },
                                [op] __device__(u64 i, u32 exc, u32) { op[i] = exc; }, table, (u32*)nullptr, stream);
        if (scatter_timer) scatter_timer->begin();
        rs_scatter_kernel<KeyT, ValT, u32, ITEMS>
            <<<unsigned(tiles), kRsThreads, smem32, stream>>>(kin, vin, kout, vout, n, shift, ws.offs32.p, tiles);
    } else {
        u64* op = ws.offs64.p;
        device_scan<u64, SumOp>([hp] __device__(u64 i) { return u64(hp[i]); },
                                [op] __device__(u64 i, u64 exc, u64) { op[i] = exc; }, table, (u64*)nullptr, stream);
        if (scatter_timer) scatter_timer->begin();
        rs_scatter_kernel<KeyT, ValT, u64, ITEMS>
            <<<unsigned(tiles), kRsThreads, smem64, stream>>>(kin, vin, kout, vout, n, shift, ws.offs64.p, tiles);
    }
    if (scatter_timer) scatter_timer->end(1, n * (2 * sizeof(KeyT) + 2 * sizeof(ValT)));
    KERNEL_CHECK();
    count_launch(2);
    if (timer) timer->end(5, n * (2 * sizeof(KeyT) + 2 * sizeof(ValT)));
}

// Sorts by the listed digit positions (each `shift` selects bits [shift, shift+8)), least significant first.
// Ping-pongs between (keys, vals) and (keys_alt, vals_alt); on return `keys`/`vals` point at the sorted data.
template <typename KeyT, typename ValT>
void radix_sort_pairs(KeyT*& keys, KeyT*& keys_alt, ValT*& vals, ValT*& vals_alt, u64 n, const int* shifts, int n_passes,
                      cudaStream_t stream, FamilyTimer* timer = nullptr, FamilyTimer* scatter_timer = nullptr) {
    if (n == 0 || n_passes == 0) return;
    RadixScratch<KeyT> ws(n, stream);
    for (int p = 0; p < n_passes; ++p) {
        radix_pass<KeyT, ValT>(ws, keys, vals, keys_alt, vals_alt, shifts[p], stream, timer, scatter_timer);
        KeyT* tk = keys; keys = keys_alt; keys_alt = tk;
        ValT* tv = vals; vals = vals_alt; vals_alt = tv;
    }
}

}  // namespace ab200
