// scan.cuh — device-wide scans (reduce-then-scan, three launches, no inter-block spinning).
// Input and output go through functors so flag extraction / scatter can be fused into the scan passes.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace ab200 {

struct SumOp {
    template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return a + b; }
};
struct MaxOp {
    template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return a > b ? a : b; }
};

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

// shuffle of any trivially copyable T whose size is a multiple of 4 bytes (scalars, pair/triple structs)
template <typename T>
__device__ __forceinline__ T shfl_up_any(T v, int d) {
    static_assert(sizeof(T) % 4 == 0, "shfl_up_any: size must be a multiple of 4");
    constexpr int W = sizeof(T) / 4;
    union { T t; u32 w[W]; } a, b;
    a.t = v;
#pragma unroll
    for (int i = 0; i < W; ++i) b.w[i] = __shfl_up_sync(0xffffffffu, a.w[i], d);
    return b.t;
}

template <typename T, typename Op>
__device__ __forceinline__ T warp_inclusive_scan(T v, Op op) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = shfl_up_any(v, d);
        if (lane_id() >= unsigned(d)) v = op(v, o);
    }
    return v;
}

// block-wide exclusive scan of one value per thread (kScanThreads threads); returns exclusive prefix, total via ref.
// Identity is T(0) for both SumOp and MaxOp (unsigned data).
template <typename T, typename Op>
__device__ __forceinline__ T block_exclusive_scan(T v, Op op, T& total, T* warp_smem /* >= 32 */) {
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const unsigned nwarps = blockDim.x >> 5;
    T inc = warp_inclusive_scan(v, op);
    if (lane == 31) warp_smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < nwarps ? warp_smem[lane] : T(0);
        T winc = warp_inclusive_scan(w, op);
        warp_smem[lane] = winc;  // inclusive over warps
    }
    __syncthreads();
    T warp_prefix = warp == 0 ? T(0) : warp_smem[warp - 1];
    total = warp_smem[nwarps - 1];
    T exc = shfl_up_any(inc, 1);
    if (lane == 0) exc = T(0);
    T r = op(warp_prefix, exc);
    __syncthreads();  // warp_smem reusable by the caller afterwards
    return r;
}

template <typename T, typename Op, typename InFn>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(InFn in, u64 n, T* __restrict__ tile_sums, Op op) {
    __shared__ T smem[32];
    const u64 base = u64(blockIdx.x) * kScanTile;
    T acc = T(0);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        u64 i = base + u64(j) * kScanThreads + threadIdx.x;
        if (i < n) acc = op(acc, T(in(i)));
    }
    T total;
    block_exclusive_scan(acc, op, total, smem);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of tile_sums in place; grand total to *total_out (may be null)
template <typename T, typename Op>
__global__ void __launch_bounds__(1024) scan_tile_sums_kernel(T* __restrict__ tile_sums, u64 num_tiles, T* total_out, Op op) {
    __shared__ T smem[32];
    T carry = T(0);
    for (u64 base = 0; base < num_tiles; base += blockDim.x) {
        u64 i = base + threadIdx.x;
        T v = i < num_tiles ? tile_sums[i] : T(0);
        T total;
        T exc = block_exclusive_scan(v, op, total, smem);
        if (i < num_tiles) tile_sums[i] = op(carry, exc);
        carry = op(carry, total);
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// final pass: out(i, exclusive_prefix_i, inclusive_prefix_i)
template <typename T, typename Op, typename InFn, typename OutFn>
__global__ void __launch_bounds__(kScanThreads) scan_final_kernel(InFn in, u64 n, const T* __restrict__ tile_offsets, OutFn out, Op op) {
    __shared__ T smem[32];
    const u64 base = u64(blockIdx.x) * kScanTile + u64(threadIdx.x) * kScanItems;
    T v[kScanItems];
    T acc = T(0);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        u64 i = base + j;
        v[j] = i < n ? T(in(i)) : T(0);
        acc = op(acc, v[j]);
    }
    T total;
    T prefix = block_exclusive_scan(acc, op, total, smem);
    prefix = op(prefix, tile_offsets[blockIdx.x]);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        u64 i = base + j;
        T inc = op(prefix, v[j]);
        if (i < n) out(i, prefix, inc);
        prefix = inc;
    }
}

// Workspace-owning scanner. total (if requested) lands in device memory `d_total`.
template <typename T, typename Op, typename InFn, typename OutFn>
void device_scan(InFn in, OutFn out, u64 n, T* d_total, cudaStream_t stream, Op op = Op()) {
    if (n == 0) {
        if (d_total) CUDA_CHECK(cudaMemsetAsync(d_total, 0, sizeof(T), stream));
        return;
    }
    const u64 tiles = ceil_div(n, u64(kScanTile));
    DevBuf<T> sums(tiles, stream);
    scan_reduce_kernel<T, Op, InFn><<<unsigned(tiles), kScanThreads, 0, stream>>>(in, n, sums.p, op);
    KERNEL_CHECK();
    scan_tile_sums_kernel<T, Op><<<1, 1024, 0, stream>>>(sums.p, tiles, d_total, op);
    KERNEL_CHECK();
    scan_final_kernel<T, Op, InFn, OutFn><<<unsigned(tiles), kScanThreads, 0, stream>>>(in, n, sums.p, out, op);
    KERNEL_CHECK();
    count_launch(3);
}

// plain element-wise pass: f(i) for i in [0, n)
template <typename F>
__global__ void for_each_index_kernel(u64 n, F f) {
    const u64 stride = u64(gridDim.x) * blockDim.x;
    for (u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}
template <typename F>
void for_each_index(u64 n, F f, cudaStream_t stream) {
    if (n == 0) return;
    const unsigned blocks = unsigned(std::min<u64>(ceil_div(n, 256), u64(kNumSMs) * 32));
    for_each_index_kernel<F><<<blocks, 256, 0, stream>>>(n, f);
    KERNEL_CHECK();
    count_launch();
}

// Two-phase form: prepare() computes the tile prefixes and the grand total (so the caller can size the outputs
// after reading the total back), finish() runs the output pass with the same input functor.
template <typename T, typename Op>
struct ScanPlan {
    DevBuf<T> sums;
    u64 tiles = 0, n = 0;
    cudaStream_t stream = nullptr;
    template <typename InFn>
    void prepare(InFn in, u64 n_, T* d_total, cudaStream_t s, Op op = Op()) {
        n = n_; stream = s;
        tiles = ceil_div(n, u64(kScanTile));
        if (n == 0) {
            if (d_total) CUDA_CHECK(cudaMemsetAsync(d_total, 0, sizeof(T), stream));
            return;
        }
        sums.alloc(tiles, stream);
        scan_reduce_kernel<T, Op, InFn><<<unsigned(tiles), kScanThreads, 0, stream>>>(in, n, sums.p, op);
        KERNEL_CHECK();
        scan_tile_sums_kernel<T, Op><<<1, 1024, 0, stream>>>(sums.p, tiles, d_total, op);
        KERNEL_CHECK();
        count_launch(2);
    }
    template <typename InFn, typename OutFn>
    void finish(InFn in, OutFn out, Op op = Op()) {
        if (n == 0) return;
        scan_final_kernel<T, Op, InFn, OutFn><<<unsigned(tiles), kScanThreads, 0, stream>>>(in, n, sums.p, out, op);
        KERNEL_CHECK();
        count_launch(1);
    }
};

}  // namespace ab200
