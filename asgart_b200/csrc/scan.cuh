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

// ---------------------------------------------------------------------------------------------------------------
// Flag scan: the scan every re-ranking step of the suffix-array build needs, specialised so that it runs on ballots.
// A functor returns up to five flag bits per element; the scan delivers, per element,
//   a / b   position of the latest element (inclusive / exclusive of itself) with FS_MARK_A / FS_MARK_B set   (running maximum)
//   c, d, e number of elements with FS_CNT_C / FS_CNT_D / FS_CNT_E set                                        (running sums)
// Elements are laid out warp-contiguous (lane = consecutive element), so the functor's loads coalesce, and the five
// prefixes of a round of 32 elements cost five ballots instead of a 20-byte shuffle scan.
enum : u32 { FS_MARK_A = 1, FS_MARK_B = 2, FS_CNT_C = 4, FS_CNT_D = 8, FS_CNT_E = 16 };

template <typename IdxT>
struct FlagAcc {
    IdxT a, b, c, d, e;
    FlagAcc() = default;
    __host__ __device__ explicit FlagAcc(int) : a(0), b(0), c(0), d(0), e(0) {}
    __host__ __device__ FlagAcc(IdxT a_, IdxT b_, IdxT c_, IdxT d_, IdxT e_) : a(a_), b(b_), c(c_), d(d_), e(e_) {}
};
struct FlagAccOp {
    template <typename T> __device__ __forceinline__ T operator()(const T& x, const T& y) const {
        return T(x.a > y.a ? x.a : y.a, x.b > y.b ? x.b : y.b, x.c + y.c, x.d + y.d, x.e + y.e);
    }
};

constexpr int kFsThreads = 512;
constexpr int kFsWarps = kFsThreads / 32;
constexpr int kFsItems = 8;
constexpr int kFsTile = kFsThreads * kFsItems;

// totals of one warp's chunk of 32 * kFsItems elements starting at `wbase`; bits[r] keeps each lane's flags
template <typename IdxT, typename FlagFn>
__device__ __forceinline__ FlagAcc<IdxT> fs_warp_totals(FlagFn& flags, u64 wbase, u64 n, u32 (&bits)[kFsItems]) {
    FlagAcc<IdxT> acc(0);
#pragma unroll
    for (int r = 0; r < kFsItems; ++r) {
        const u64 c = wbase + u64(r) * 32 + lane_id();
        bits[r] = c < n ? u32(flags(c)) : 0u;
        const unsigned ma = __ballot_sync(0xffffffffu, bits[r] & FS_MARK_A), mb = __ballot_sync(0xffffffffu, bits[r] & FS_MARK_B);
        if (ma) acc.a = IdxT(wbase + u64(r) * 32 + u64(31 - __clz(ma)));
        if (mb) acc.b = IdxT(wbase + u64(r) * 32 + u64(31 - __clz(mb)));
        acc.c += IdxT(__popc(__ballot_sync(0xffffffffu, bits[r] & FS_CNT_C)));
        acc.d += IdxT(__popc(__ballot_sync(0xffffffffu, bits[r] & FS_CNT_D)));
        acc.e += IdxT(__popc(__ballot_sync(0xffffffffu, bits[r] & FS_CNT_E)));
    }
    return acc;
}

template <typename IdxT, typename FlagFn>
__global__ void __launch_bounds__(kFsThreads) fs_reduce_kernel(FlagFn flags, u64 n, FlagAcc<IdxT>* __restrict__ tile_sums) {
    __shared__ FlagAcc<IdxT> wsum[kFsWarps];
    const u32 warp = threadIdx.x >> 5;
    const u64 wbase = u64(blockIdx.x) * kFsTile + u64(warp) * (32 * kFsItems);
    u32 bits[kFsItems];
    const FlagAcc<IdxT> acc = fs_warp_totals<IdxT>(flags, wbase, n, bits);
    if (lane_id() == 0) wsum[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        FlagAcc<IdxT> t(0);
        FlagAccOp op;
#pragma unroll
        for (int w = 0; w < kFsWarps; ++w) t = op(t, wsum[w]);
        tile_sums[blockIdx.x] = t;
    }
}

// out(i, exclusive, inclusive)
template <typename IdxT, typename FlagFn, typename OutFn>
__global__ void __launch_bounds__(kFsThreads) fs_final_kernel(FlagFn flags, u64 n, const FlagAcc<IdxT>* __restrict__ tile_offsets, OutFn out) {
    __shared__ FlagAcc<IdxT> wsum[kFsWarps];
    const u32 warp = threadIdx.x >> 5, lane = lane_id();
    const unsigned lt = lanemask_lt();
    const u64 wbase = u64(blockIdx.x) * kFsTile + u64(warp) * (32 * kFsItems);
    u32 bits[kFsItems];
    const FlagAcc<IdxT> mine = fs_warp_totals<IdxT>(flags, wbase, n, bits);
    if (lane == 0) wsum[warp] = mine;
    __syncthreads();
    FlagAcc<IdxT> carry = tile_offsets[blockIdx.x];
    FlagAccOp op;
    for (u32 w = 0; w < warp; ++w) carry = op(carry, wsum[w]);
#pragma unroll
    for (int r = 0; r < kFsItems; ++r) {
        const u64 rb = wbase + u64(r) * 32;
        const u64 c = rb + lane;
        const unsigned ma = __ballot_sync(0xffffffffu, bits[r] & FS_MARK_A), mb = __ballot_sync(0xffffffffu, bits[r] & FS_MARK_B);
        const unsigned mc = __ballot_sync(0xffffffffu, bits[r] & FS_CNT_C), md = __ballot_sync(0xffffffffu, bits[r] & FS_CNT_D);
        const unsigned me = __ballot_sync(0xffffffffu, bits[r] & FS_CNT_E);
        if (c < n) {
            FlagAcc<IdxT> exc, inc;
            const unsigned la = ma & lt, lb = mb & lt;
            exc.a = la ? IdxT(rb + u64(31 - __clz(la))) : carry.a;
            exc.b = lb ? IdxT(rb + u64(31 - __clz(lb))) : carry.b;
            exc.c = carry.c + IdxT(__popc(mc & lt));
            exc.d = carry.d + IdxT(__popc(md & lt));
            exc.e = carry.e + IdxT(__popc(me & lt));
            inc.a = (bits[r] & FS_MARK_A) ? IdxT(c) : exc.a;
            inc.b = (bits[r] & FS_MARK_B) ? IdxT(c) : exc.b;
            inc.c = exc.c + ((bits[r] & FS_CNT_C) ? 1 : 0);
            inc.d = exc.d + ((bits[r] & FS_CNT_D) ? 1 : 0);
            inc.e = exc.e + ((bits[r] & FS_CNT_E) ? 1 : 0);
            out(c, exc, inc);
        }
        if (ma) carry.a = IdxT(rb + u64(31 - __clz(ma)));
        if (mb) carry.b = IdxT(rb + u64(31 - __clz(mb)));
        carry.c += IdxT(__popc(mc));
        carry.d += IdxT(__popc(md));
        carry.e += IdxT(__popc(me));
    }
}

// prepare() computes the tile prefixes and the grand total (so the caller can size the outputs after reading the
// total back), finish() runs the output pass with the same flag functor.
template <typename IdxT>
struct FlagScanPlan {
    using Acc = FlagAcc<IdxT>;
    DevBuf<Acc> sums;
    u64 tiles = 0, n = 0;
    cudaStream_t stream = nullptr;
    template <typename FlagFn>
    void prepare(FlagFn flags, u64 n_, Acc* d_total, cudaStream_t s) {
        n = n_; stream = s;
        tiles = ceil_div(n, u64(kFsTile));
        if (n == 0) {
            if (d_total) CUDA_CHECK(cudaMemsetAsync(d_total, 0, sizeof(Acc), stream));
            return;
        }
        sums.alloc(tiles, stream);
        fs_reduce_kernel<IdxT, FlagFn><<<unsigned(tiles), kFsThreads, 0, stream>>>(flags, n, sums.p);
        KERNEL_CHECK();
        scan_tile_sums_kernel<Acc, FlagAccOp><<<1, 1024, 0, stream>>>(sums.p, tiles, d_total, FlagAccOp());
        KERNEL_CHECK();
        count_launch(2);
    }
    template <typename FlagFn, typename OutFn>
    void finish(FlagFn flags, OutFn out) {
        if (n == 0) return;
        fs_final_kernel<IdxT, FlagFn, OutFn><<<unsigned(tiles), kFsThreads, 0, stream>>>(flags, n, sums.p, out);
        KERNEL_CHECK();
        count_launch(1);
    }
};

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
