// scatter.cuh — out[idx[i]] = val[i] for an injective idx: the inverse-permutation scatter that turns "rank by sorted
// position" into "rank by text position" in the suffix-array build.
//
// A plain scatter of 4-byte values to random addresses costs a DRAM read-modify-write per element (B200, measured with
// tools/scatter_bench.cu: 30.8 G elem/s for 57 M targets, 22.7 G elem/s for 1 G). When all targets of a launch fall
// inside a slice of `out` that the L2 holds (<= 48 MB), the writes combine in L2 and leave as whole lines: 98 G elem/s
// with four sweeps over 57 M elements. So:
//   out fits one slice          -> plain scatter
//   a few slices (<= kSweepMax) -> one sweep over (idx, val) per slice, each writing only its slice's targets
//   more                        -> one radix partition pass of the pairs by slice number (16 MB slices, up to 256 of them),
//                                  then one in-order scatter from few blocks: 51 G elem/s at 250 M targets with 148 x 4
//                                  blocks, but only 26 G elem/s with 148 x 16 (too many partial lines in flight)
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"

namespace ab200 {

constexpr u64 kScatterSliceBytes = u64(48) << 20;
constexpr int kSweepMax = 6;

template <typename IdxT>
__global__ void __launch_bounds__(256) scatter_slice_kernel(const IdxT* __restrict__ idx, const IdxT* __restrict__ val, u64 n,
                                                            IdxT* __restrict__ out, u64 lo, u64 hi) {
    constexpr int V = 16 / sizeof(IdxT);
    const u64 stride = u64(gridDim.x) * blockDim.x * V;
    for (u64 i = (u64(blockIdx.x) * blockDim.x + threadIdx.x) * V; i < n; i += stride) {
        if (i + V <= n) {
            IdxT p[V], v[V];
            *reinterpret_cast<uint4*>(p) = __ldcs(reinterpret_cast<const uint4*>(idx + i));   // streaming: the L2 is for the written lines
            *reinterpret_cast<uint4*>(v) = __ldcs(reinterpret_cast<const uint4*>(val + i));
#pragma unroll
            for (int j = 0; j < V; ++j)
                if (u64(p[j]) >= lo && u64(p[j]) < hi) out[p[j]] = v[j];
        } else {
            for (u64 j = i; j < n; ++j) { const u64 p = u64(idx[j]); if (p >= lo && p < hi) out[p] = val[j]; }
        }
    }
}

// scratch_idx / scratch_val: n elements each (may be null: the partition path then allocates them)
template <typename IdxT>
void inverse_scatter(const IdxT* idx, const IdxT* val, u64 n, IdxT* out, u64 out_len, IdxT* scratch_idx, IdxT* scratch_val,
                     cudaStream_t stream) {
    if (n == 0) return;
    const u64 slice = kScatterSliceBytes / sizeof(IdxT);
    const u64 K = ceil_div(out_len, slice);
    const unsigned grid = unsigned(std::min<u64>(ceil_div(n, 256 * (16 / sizeof(IdxT))), u64(kNumSMs) * 16));
    if (K <= u64(kSweepMax) || sizeof(IdxT) != 4) {
        const u64 sweeps = sizeof(IdxT) != 4 && K > u64(kSweepMax) ? 1 : K;   // 64-bit indices beyond the sweep range: plain scatter
        for (u64 k = 0; k < sweeps; ++k) {
            const u64 lo = out_len * k / sweeps, hi = out_len * (k + 1) / sweeps;
            scatter_slice_kernel<IdxT><<<grid, 256, 0, stream>>>(idx, val, n, out, lo, hi);
            KERNEL_CHECK();
        }
        count_launch(sweeps);
        return;
    }
    if constexpr (sizeof(IdxT) == 4) {
        // partition by slice number = idx >> shift, at most 256 slices (one radix pass; slices grow beyond 48 MB past 3.2 G targets)
        int shift = 22;
        while ((out_len - 1) >> shift > 255) ++shift;
        DevBuf<u32> own_i, own_v;
        if (!scratch_idx) { own_i.alloc(n, stream); scratch_idx = own_i.p; }
        if (!scratch_val) { own_v.alloc(n, stream); scratch_val = own_v.p; }
        u32 *k = const_cast<u32*>(idx), *ka = scratch_idx, *v = const_cast<u32*>(val), *va = scratch_val;
        radix_sort_pairs<u32, u32>(k, ka, v, va, n, &shift, 1, stream);   // one pass: reads (idx, val), writes the scratch pair
        scatter_slice_kernel<u32><<<kNumSMs * 4, 256, 0, stream>>>(k, v, n, out, 0, out_len);
        KERNEL_CHECK();
        count_launch(1);
    }
}

}  // namespace ab200
