// radix_sort.cuh — hand-written stable LSD radix sort of (key, value) pairs for sm_100a.
//
// One 8-bit digit per pass, three launches per pass and no inter-block spinning:
//   rs_hist_kernel     per-tile digit histogram (warp-aggregated shared-memory atomics)        reads keys
//   device_scan        exclusive scan of the digit-major table hist[digit][tile] -> global offset of every (digit, tile)
//   rs_scatter_kernel  stable in-tile ranking with __match_any_sync, tile reordered in shared memory so that each
//                      digit's run leaves the SM as contiguous, coalesced stores               reads+writes keys, values
// Keys are u64 or U128 (two u64 words; used when suffix indices need more than 32 bits), values u32 or u64.
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

__device__ __forceinline__ u32 rs_digit(u64 k, int shift) { return u32(k >> shift) & 255u; }
__device__ __forceinline__ u32 rs_digit(const U128& k, int shift) {
    if (shift >= 64) return u32(k.hi >> (shift - 64)) & 255u;
    u64 v = k.lo >> shift;
    if (shift > 56) v |= k.hi << (64 - shift);
    return u32(v) & 255u;
}

constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
template <typename KeyT> struct RsItems { static constexpr int value = sizeof(KeyT) > 8 ? 8 : 16; };

template <typename KeyT, int ITEMS>
__global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(const KeyT* __restrict__ kin, u64 n, int shift,
                                                             u32* __restrict__ hist, u64 num_tiles) {
    constexpr int TILE = kRsThreads * ITEMS;
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const u64 base = u64(blockIdx.x) * TILE;
    const u32 cnt = u32(min(u64(TILE), n - base));
#pragma unroll 4
    for (int j = 0; j < ITEMS; ++j) {
        u32 li = j * kRsThreads + threadIdx.x;
        bool valid = li < cnt;
        u32 d = valid ? rs_digit(kin[base + li], shift) : 256u;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && (__ffs(peers) - 1) == int(lane_id())) atomicAdd(&h[d], u32(__popc(peers)));
    }
    __syncthreads();
    hist[u64(threadIdx.x) * num_tiles + blockIdx.x] = h[threadIdx.x];
}

template <typename KeyT, typename ValT, typename OffT, int ITEMS>
__global__ void __launch_bounds__(kRsThreads) rs_scatter_kernel(const KeyT* __restrict__ kin, const ValT* __restrict__ vin,
                                                                KeyT* __restrict__ kout, ValT* __restrict__ vout, u64 n,
                                                                int shift, const OffT* __restrict__ offs, u64 num_tiles) {
    constexpr int TILE = kRsThreads * ITEMS;
    static_assert(sizeof(ValT) <= sizeof(KeyT), "value staging reuses the key buffer");
    __shared__ __align__(16) KeyT stage[TILE];
    __shared__ u32 whist[kRsWarps][256];
    __shared__ u32 tile_off[256];
    __shared__ OffT delta[256];
    __shared__ u32 scan_smem[32];

    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const u64 base = u64(blockIdx.x) * TILE;
    const u32 cnt = u32(min(u64(TILE), n - base));
    for (u32 i = tid; i < kRsWarps * 256; i += kRsThreads) (&whist[0][0])[i] = 0;
    __syncthreads();

    KeyT keys[ITEMS];
    u32 rnk[ITEMS];  // rank inside (warp, digit), later the slot inside the tile
    const u32 wbase = warp * (32 * ITEMS);
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const u32 li = wbase + r * 32 + lane;
        const bool valid = li < cnt;
        if (valid) keys[r] = kin[base + li];
        const u32 d = valid ? rs_digit(keys[r], shift) : 256u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        u32 old = 0;
        if (valid && int(lane) == leader) {
            old = whist[warp][d];
            whist[warp][d] = old + u32(__popc(peers));
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rnk[r] = old + u32(__popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();

    {  // thread d: exclusive prefix over warps for digit d, then over digits
        const u32 d = tid;
        u32 sum = 0;
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w) {
            u32 t = whist[w][d];
            whist[w][d] = sum;
            sum += t;
        }
        u32 total;
        u32 excl = block_exclusive_scan<u32, SumOp>(sum, SumOp(), total, scan_smem);
        tile_off[d] = excl;
        delta[d] = offs[u64(d) * num_tiles + blockIdx.x] - OffT(excl);
    }
    __syncthreads();

#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const u32 li = wbase + r * 32 + lane;
        if (li < cnt) {
            const u32 d = rs_digit(keys[r], shift);
            const u32 lp = tile_off[d] + whist[warp][d] + rnk[r];
            rnk[r] = lp;
            stage[lp] = keys[r];
        }
    }
    __syncthreads();

    OffT gp[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const u32 lp = j * kRsThreads + tid;
        if (lp < cnt) {
            const KeyT k = stage[lp];
            gp[j] = delta[rs_digit(k, shift)] + OffT(lp);
            kout[gp[j]] = k;
        }
    }
    __syncthreads();

    ValT* vstage = reinterpret_cast<ValT*>(stage);
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const u32 li = wbase + r * 32 + lane;
        if (li < cnt) vstage[rnk[r]] = vin[base + li];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const u32 lp = j * kRsThreads + tid;
        if (lp < cnt) vout[gp[j]] = vstage[lp];
    }
}

// Sorts by the listed digit positions (each `shift` selects bits [shift, shift+8)), least significant first.
// Ping-pongs between (keys, vals) and (keys_alt, vals_alt); on return `keys`/`vals` point at the sorted data.
template <typename KeyT, typename ValT>
void radix_sort_pairs(KeyT*& keys, KeyT*& keys_alt, ValT*& vals, ValT*& vals_alt, u64 n, const int* shifts, int n_passes,
                      cudaStream_t stream, FamilyTimer* timer = nullptr, FamilyTimer* scatter_timer = nullptr) {
    if (n == 0 || n_passes == 0) return;
    constexpr int ITEMS = RsItems<KeyT>::value;
    constexpr int TILE = kRsThreads * ITEMS;
    const u64 tiles = ceil_div(n, u64(TILE));
    const u64 table = tiles * 256;
    const bool wide = n >= (u64(1) << 32);
    DevBuf<u32> hist(table, stream);
    DevBuf<u32> offs32(wide ? 0 : table, stream);
    DevBuf<u64> offs64(wide ? table : 0, stream);
    for (int p = 0; p < n_passes; ++p) {
        const int shift = shifts[p];
        if (timer) timer->begin();
        rs_hist_kernel<KeyT, ITEMS><<<unsigned(tiles), kRsThreads, 0, stream>>>(keys, n, shift, hist.p, tiles);
        KERNEL_CHECK();
        const u32* hp = hist.p;
        if (!wide) {
            u32* op = offs32.p;
            device_scan<u32, SumOp>([hp] __device__(u64 i) { return hp[i]; },
                                    [op] __device__(u64 i, u32 exc, u32) { op[i] = exc; }, table, (u32*)nullptr, stream);
            if (scatter_timer) scatter_timer->begin();
            rs_scatter_kernel<KeyT, ValT, u32, ITEMS>
                <<<unsigned(tiles), kRsThreads, 0, stream>>>(keys, vals, keys_alt, vals_alt, n, shift, offs32.p, tiles);
        } else {
            u64* op = offs64.p;
            device_scan<u64, SumOp>([hp] __device__(u64 i) { return u64(hp[i]); },
                                    [op] __device__(u64 i, u64 exc, u64) { op[i] = exc; }, table, (u64*)nullptr, stream);
            if (scatter_timer) scatter_timer->begin();
            rs_scatter_kernel<KeyT, ValT, u64, ITEMS>
                <<<unsigned(tiles), kRsThreads, 0, stream>>>(keys, vals, keys_alt, vals_alt, n, shift, offs64.p, tiles);
        }
        if (scatter_timer) scatter_timer->end(1, n * (2 * sizeof(KeyT) + 2 * sizeof(ValT)));
        KERNEL_CHECK();
        count_launch(2);
        if (timer) timer->end(5, n * (2 * sizeof(KeyT) + 2 * sizeof(ValT)));
        KeyT* tk = keys; keys = keys_alt; keys_alt = tk;
        ValT* tv = vals; vals = vals_alt; vals_alt = tv;
    }
}

}  // namespace ab200
