// scatter.cuh — out[idx[i]] = val[i] for an injective idx: the inverse-permutation scatter that turns "rank by sorted
// position" into "rank by text position" in the suffix-array build.
//
// A plain scatter of 4-byte values to random addresses costs a DRAM read-modify-write per element (B200, measured with
// tools/scatter_bench.cu: 30.8 G elem/s for 57 M targets, 22.7 G elem/s for 1 G). When all targets of a launch fall
// inside a slice of `out` that the L2 holds (<= 48 MB), the writes combine in L2 and leave as whole lines: 98 G elem/s
// with four sweeps over 57 M elements. So:
//   out fits one slice          -> plain scatter
//   a few slices (<= kSweepMax) -> one sweep over (idx, val) per slice, each writing only its slice's targets
//   more                        -> plain scatter. Tried and dropped: one radix partition pass of the pairs by slice number
//                                  followed by an in-order scatter (profiles/r1_scatter_bench.log): the in-order scatter
//                                  reaches 26-51 G elem/s only, and with the partition pass the whole is slower than the
//                                  plain scatter at 249 M (15.0 vs 15.5 ms for the re-ranking) and at 3.1 G targets (229 vs 193 ms).
#pragma once
#include "common.cuh"

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

template <typename IdxT>
void inverse_scatter(const IdxT* idx, const IdxT* val, u64 n, IdxT* out, u64 out_len, cudaStream_t stream) {
    if (n == 0) return;
    const u64 slice = kScatterSliceBytes / sizeof(IdxT);
    u64 sweeps = ceil_div(out_len, slice);
    if (sweeps > u64(kSweepMax)) sweeps = 1;
    const unsigned grid = unsigned(std::min<u64>(ceil_div(n, 256 * (16 / sizeof(IdxT))), u64(kNumSMs) * 16));
    for (u64 k = 0; k < sweeps; ++k) {
        const u64 lo = out_len * k / sweeps, hi = out_len * (k + 1) / sweeps;
        scatter_slice_kernel<IdxT><<<grid, 256, 0, stream>>>(idx, val, n, out, lo, hi);
        KERNEL_CHECK();
    }
    count_launch(sweeps);
}

}  // namespace ab200
