// sa_build.cuh — on-GPU suffix-array construction by prefix doubling (replaces libdivsufsort behind
// src/divsufsort.rs:10 / r_divsufsort, src/bin/asgart.rs:473-479, of the reference).
//
// The suffix array of a text is unique, so any correct construction reproduces divsufsort64's output bit for bit
// (checked against the reference's own library in tests/). Algorithm (Larsson–Sadakane doubling with discarding,
// laid out for a GPU):
//   0. byte histogram -> dense symbol codes 1..sigma (0 = "past the end", which orders a suffix that is a prefix of
//      another one first, as divsufsort does); b bits per symbol, p0 = floor(64 / b) symbols per 64-bit key
//      (DNA + '$': sigma = 6, b = 3, p0 = 21).
//   1. key[i] = first p0 symbols of suffix i; one LSD radix sort of (key, i) over b*p0 bits.
//   2. group heads by key change -> rank[] (= SA index of the group head) and SA[] for all; suffixes in groups of
//      size > 1 are compacted into the work list (G = group head index, I = suffix).
//   3. doubling rounds h = p0, 2p0, ...: key2 = rank[I + h] + 1 (0 past the end)  [random gather],
//      radix sort of the work list by (G, key2), re-rank, write SA for suffixes that became unique, drop them,
//      until the work list is empty.
// Index width is a template parameter: u32 when n < 2^32 - 1 (all BASELINE configs), u64 beyond (composite keys then
// need 128 bits: U128). Both paths are exercised by the tests on small inputs.
#pragma once
#include <vector>

#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace ab200 {

struct SaStats {
    FamilyTimer* sort = nullptr;    // radix passes
    FamilyTimer* scatter = nullptr; // rs_scatter_kernel alone
    FamilyTimer* gather = nullptr;  // rank[I + h] gathers
    FamilyTimer* rank = nullptr;    // re-rank / compaction scans
    u64 rounds = 0;
};

// ---------------------------------------------------------------------------------------------------------------
__global__ void byte_hist_kernel(const u8* __restrict__ t, u64 n, unsigned long long* __restrict__ hist) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;  // 256 threads
    __syncthreads();
    const u64 stride = u64(gridDim.x) * blockDim.x * 16;
    for (u64 i = (u64(blockIdx.x) * blockDim.x + threadIdx.x) * 16; i < n; i += stride) {
        if (i + 16 <= n && (reinterpret_cast<uintptr_t>(t + i) & 15) == 0) {
            uint4 v = *reinterpret_cast<const uint4*>(t + i);
            u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                atomicAdd(&h[w[k] & 255], 1u); atomicAdd(&h[(w[k] >> 8) & 255], 1u);
                atomicAdd(&h[(w[k] >> 16) & 255], 1u); atomicAdd(&h[w[k] >> 24], 1u);
            }
        } else {
            for (u64 j = i; j < n && j < i + 16; ++j) atomicAdd(&h[t[j]], 1u);
        }
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

constexpr int kInitThreads = 256;
constexpr int kInitItems = 16;
constexpr int kInitTile = kInitThreads * kInitItems;

// key[i] = codes of T[i .. i+p0) packed most-significant-first, b bits each; vals[i] = i
template <typename IdxT>
__global__ void __launch_bounds__(kInitThreads) init_keys_kernel(const u8* __restrict__ text, u64 n, const uint16_t* __restrict__ code,
                                                                  int b, int p0, u64* __restrict__ keys, IdxT* __restrict__ vals) {
    __shared__ uint16_t sc[kInitTile + 64];   // codes reach 256 when every byte value occurs
    __shared__ uint16_t scode[256];
    scode[threadIdx.x] = code[threadIdx.x];
    __syncthreads();
    const u64 base = u64(blockIdx.x) * kInitTile;
    for (u32 i = threadIdx.x; i < kInitTile + 64; i += kInitThreads) {
        u64 p = base + i;
        sc[i] = p < n ? scode[text[p]] : uint16_t(0);
    }
    __syncthreads();
    const u32 t0 = threadIdx.x * kInitItems;
    const u64 mask = (b * p0 >= 64) ? ~u64(0) : ((u64(1) << (b * p0)) - 1);
    u64 key = 0;
    for (int j = 0; j < p0; ++j) key = (key << b) | sc[t0 + j];
#pragma unroll
    for (int j = 0; j < kInitItems; ++j) {
        u64 p = base + t0 + j;
        if (p < n) { keys[p] = key; vals[p] = IdxT(p); }
        key = ((key << b) & mask) | sc[t0 + j + p0];
    }
}

// ---------------------------------------------------------------------------------------------------------------
template <typename IdxT>
struct MaxCnt {
    IdxT mx, cnt;
    MaxCnt() = default;
    __host__ __device__ explicit MaxCnt(int) : mx(0), cnt(0) {}
    __host__ __device__ MaxCnt(IdxT m, IdxT c) : mx(m), cnt(c) {}
};
struct MaxCntOp {
    template <typename T> __device__ __forceinline__ T operator()(const T& a, const T& b) const {
        return T(a.mx > b.mx ? a.mx : b.mx, a.cnt + b.cnt);
    }
};
template <typename IdxT>
struct Max2Cnt {
    IdxT cg, cs, cnt;
    Max2Cnt() = default;
    __host__ __device__ explicit Max2Cnt(int) : cg(0), cs(0), cnt(0) {}
    __host__ __device__ Max2Cnt(IdxT a, IdxT b, IdxT c) : cg(a), cs(b), cnt(c) {}
};
struct Max2CntOp {
    template <typename T> __device__ __forceinline__ T operator()(const T& a, const T& b) const {
        return T(a.cg > b.cg ? a.cg : b.cg, a.cs > b.cs ? a.cs : b.cs, a.cnt + b.cnt);
    }
};

// composite (group, key2) keys
template <typename IdxT> struct CompKey;
template <> struct CompKey<u32> {
    using type = u64;
    __device__ static __forceinline__ u64 make(u32 g, u32 k2, int kb) { return (u64(g) << kb) | u64(k2); }
    __device__ static __forceinline__ u32 group(u64 k, int kb) { return u32(k >> kb); }
};
template <> struct CompKey<u64> {
    using type = U128;
    __device__ static __forceinline__ U128 make(u64 g, u64 k2, int) { return U128{g, k2}; }
    __device__ static __forceinline__ u64 group(const U128& k, int) { return k.hi; }
};

template <typename IdxT>
__global__ void gather_rank_kernel(const IdxT* __restrict__ G, const IdxT* __restrict__ I, const IdxT* __restrict__ rank,
                                   u64 U, u64 n, u64 h, int kb, typename CompKey<IdxT>::type* __restrict__ K) {
    const u64 stride = u64(gridDim.x) * blockDim.x;
    for (u64 c = u64(blockIdx.x) * blockDim.x + threadIdx.x; c < U; c += stride) {
        const u64 p = u64(I[c]) + h;
        const IdxT k2 = p < n ? IdxT(rank[p] + 1) : IdxT(0);
        K[c] = CompKey<IdxT>::make(G[c], k2, kb);
    }
}

template <typename IdxT>
void build_suffix_array(const u8* d_text, u64 n, IdxT* d_sa, IdxT* d_rank, cudaStream_t stream, SaStats* st) {
    using KeyT = typename CompKey<IdxT>::type;
    if (n == 0) return;

    // ---- 0. alphabet
    DevBuf<unsigned long long> d_hist(256, stream);
    d_hist.zero();
    {
        int blocks = int(std::min<u64>(ceil_div(n, 256 * 16), u64(kNumSMs) * 8));
        byte_hist_kernel<<<blocks, 256, 0, stream>>>(d_text, n, d_hist.p);
        KERNEL_CHECK();
        count_launch();
    }
    unsigned long long h_hist[256];
    CUDA_CHECK(cudaMemcpyAsync(h_hist, d_hist.p, sizeof h_hist, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    uint16_t h_code[256];
    int sigma = 0;
    for (int c = 0; c < 256; ++c) h_code[c] = h_hist[c] ? uint16_t(++sigma) : uint16_t(0);
    const int b = std::max(1, bit_width_u64(u64(sigma)));  // codes 0..sigma
    const int p0 = 64 / b;
    DevBuf<uint16_t> d_code(256, stream);
    CUDA_CHECK(cudaMemcpyAsync(d_code.p, h_code, sizeof h_code, cudaMemcpyHostToDevice, stream));

    // ---- 1. initial keys + sort
    DevBuf<IdxT> Gbuf, Ibuf;
    u64 U = 0;
    DevBuf<IdxT> d_total_mc_store;  // unused placeholder to keep allocation order simple
    {
        DevBuf<u64> keysA(n, stream), keysB(n, stream);
        DevBuf<IdxT> valsA(n, stream), valsB(n, stream);
        init_keys_kernel<IdxT><<<unsigned(ceil_div(n, u64(kInitTile))), kInitThreads, 0, stream>>>(d_text, n, d_code.p, b, p0,
                                                                                                  keysA.p, valsA.p);
        KERNEL_CHECK();
        count_launch();
        std::vector<int> shifts;
        for (int s = 0; s < b * p0; s += 8) shifts.push_back(s);
        u64 *k = keysA.p, *ka = keysB.p;
        IdxT *v = valsA.p, *va = valsB.p;
        radix_sort_pairs<u64, IdxT>(k, ka, v, va, n, shifts.data(), int(shifts.size()), stream, st ? st->sort : nullptr,
                                    st ? st->scatter : nullptr);

        // ---- 2. heads -> rank, SA, work list
        using MC = MaxCnt<IdxT>;
        DevBuf<MC> d_total(1, stream);
        const u64* kk = k;
        const IdxT* vv = v;
        auto in = [kk, n] __device__(u64 i) {
            const u64 cur = kk[i];
            const bool head = i == 0 || kk[i - 1] != cur;
            const bool tail = i + 1 == n || kk[i + 1] != cur;
            return MC(head ? IdxT(i) : IdxT(0), (head && tail) ? IdxT(0) : IdxT(1));
        };
        if (st && st->rank) st->rank->begin();
        ScanPlan<MC, MaxCntOp> plan;
        plan.prepare(in, n, d_total.p, stream);
        MC h_total;
        CUDA_CHECK(cudaMemcpyAsync(&h_total, d_total.p, sizeof(MC), cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        U = u64(h_total.cnt);
        Gbuf.alloc(U, stream);
        Ibuf.alloc(U, stream);
        IdxT *G = Gbuf.p, *I = Ibuf.p;
        plan.finish(in, [kk, vv, n, d_rank, d_sa, G, I] __device__(u64 i, const MC& exc, const MC& inc) {
            const u64 cur = kk[i];
            const bool head = i == 0 || kk[i - 1] != cur;
            const bool tail = i + 1 == n || kk[i + 1] != cur;
            const IdxT idx = vv[i];
            d_rank[idx] = inc.mx;
            d_sa[i] = idx;
            if (!(head && tail)) { G[exc.cnt] = inc.mx; I[exc.cnt] = idx; }
        });
        if (st && st->rank) st->rank->end(3, 0);
    }

    // ---- 3. doubling rounds
    const int kb = std::max(1, bit_width_u64(n));      // key2 in [0, n]
    const int gb = std::max(1, bit_width_u64(n - 1));  // group head index in [0, n)
    std::vector<int> shifts;
    if (sizeof(IdxT) == 4) {
        for (int s = 0; s < kb + gb; s += 8) shifts.push_back(s);
    } else {
        for (int s = 0; s < kb; s += 8) shifts.push_back(s);
        for (int s = 0; s < gb; s += 8) shifts.push_back(64 + s);
    }
    u64 h = u64(p0);
    while (U > 0) {
        if (st) st->rounds++;
        DevBuf<KeyT> KA(U, stream), KB(U, stream);
        DevBuf<IdxT> IB(U, stream);
        {
            if (st && st->gather) st->gather->begin();
            int blocks = int(std::min<u64>(ceil_div(U, 256), u64(kNumSMs) * 16));
            gather_rank_kernel<IdxT><<<blocks, 256, 0, stream>>>(Gbuf.p, Ibuf.p, d_rank, U, n, h, kb, KA.p);
            KERNEL_CHECK();
            count_launch();
            if (st && st->gather) st->gather->end(1, U * 3 * sizeof(IdxT));
        }
        KeyT *k = KA.p, *ka = KB.p;
        IdxT *v = Ibuf.p, *va = IB.p;
        radix_sort_pairs<KeyT, IdxT>(k, ka, v, va, U, shifts.data(), int(shifts.size()), stream, st ? st->sort : nullptr,
                                     st ? st->scatter : nullptr);

        using M2 = Max2Cnt<IdxT>;
        DevBuf<M2> d_total(1, stream);
        const KeyT* kk = k;
        const IdxT* vv = v;
        const u64 Uc = U;
        auto in = [kk, Uc, kb] __device__(u64 c) {
            const KeyT cur = kk[c];
            bool head_new = true, head_old = true, tail_new = true;
            if (c > 0) {
                const KeyT prev = kk[c - 1];
                head_new = prev != cur;
                head_old = CompKey<IdxT>::group(prev, kb) != CompKey<IdxT>::group(cur, kb);
            }
            if (c + 1 < Uc) tail_new = kk[c + 1] != cur;
            return M2(head_old ? IdxT(c) : IdxT(0), head_new ? IdxT(c) : IdxT(0), (head_new && tail_new) ? IdxT(0) : IdxT(1));
        };
        if (st && st->rank) st->rank->begin();
        ScanPlan<M2, Max2CntOp> plan;
        plan.prepare(in, U, d_total.p, stream);
        M2 h_total;
        CUDA_CHECK(cudaMemcpyAsync(&h_total, d_total.p, sizeof(M2), cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        const u64 U2 = u64(h_total.cnt);
        DevBuf<IdxT> G2(U2, stream), I2(U2, stream);
        IdxT *g2 = G2.p, *i2 = I2.p;
        plan.finish(in, [kk, vv, Uc, kb, d_rank, d_sa, g2, i2] __device__(u64 c, const M2& exc, const M2& inc) {
            const KeyT cur = kk[c];
            const bool head_new = c == 0 || kk[c - 1] != cur;
            const bool tail_new = c + 1 == Uc || kk[c + 1] != cur;
            const IdxT g = CompKey<IdxT>::group(cur, kb);
            const IdxT idx = vv[c];
            const IdxT ng = g + (inc.cs - inc.cg);
            d_rank[idx] = ng;
            if (head_new && tail_new) d_sa[u64(g) + (c - u64(inc.cg))] = idx;
            else { g2[exc.cnt] = ng; i2[exc.cnt] = idx; }
        });
        if (st && st->rank) st->rank->end(3, 0);
        // v may point at Ibuf or IB; both are released when replaced / at scope end (stream-ordered)
        Gbuf = std::move(G2);
        Ibuf = std::move(I2);
        U = U2;
        h <<= 1;
    }
}

}  // namespace ab200
