// search.cuh — text packing, 8-mer LUT, probe search (stage A) kernels.
//   pack_text_kernel      ASCII strand -> 4-bit packed text, or its complemented / reversed / reverse-complemented
//                         image (the needle of src/bin/asgart.rs:206-218 for every chunk at once)      HBM stream
//   lut_build_kernel      SA interval of every 8-mer (Searcher::new, src/searcher.rs:99-143)              HBM gather
//   probe_search_kernel   one lane per probe position: LUT narrow -> literal lock-step equal range -> filtered count
//                         (Searcher::search src/searcher.rs:145-180 + src/automaton.rs:100-117)            HBM gather
//   emit (inside the scan's output pass): surviving matches in SA order + event list
#pragma once
#include "common.cuh"
#include "kmer_core.h"
#include "scan.cuh"

namespace ab200 {

enum PackMode : int { PACK_DIRECT = 0, PACK_COMPLEMENT = 1, PACK_REVERSE = 2, PACK_REVCOMP = 3 };

// One u64 word (16 symbols) per thread. Direct mode packs all n1 bytes (including '$'); the needle modes pack the
// n = n1-1 bases only: needle_image[j] = f(text[j]) (complement) or f(text[n-1-j]) (reversed modes).
// err[0] |= 1 if a byte is outside {A,C,G,N,T} or '$' is misplaced.
__global__ void pack_text_kernel(const u8* __restrict__ text, u64 n1, int mode, u64* __restrict__ packed, u64 n_words,
                                 u32* __restrict__ err) {
    const u64 w = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const u64 n = n1 - 1;
    const u64 limit = mode == PACK_DIRECT ? n1 : n;
    u64 word = 0;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const u64 p = w * 16 + j;
        u32 c = CODE_PAD;
        if (p < limit) {
            const u64 src = (mode & 2) ? (n - 1 - p) : p;
            c = code_of_byte(text[src]);
            if (c == CODE_BAD) bad = true;
            if ((c == CODE_END) != (src == n)) bad = true;
            if (mode & 1) c = complement_code(c);
        }
        word |= u64(c & 15u) << (60 - 4 * j);
    }
    packed[w] = word;
    if (bad) atomicOr(err, 1u);
}

// lo/hi of every 8-mer bucket from the suffix array: boundaries where the first 8 symbols change.
template <typename IdxT>
__global__ void lut_build_kernel(const u64* __restrict__ PT, const IdxT* __restrict__ SA, u64 n1, IdxT* __restrict__ lut_lo,
                                 IdxT* __restrict__ lut_hi) {
    const u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    u32 cur_slot, prev_slot = 0;
    const bool cur_ok = lut_slot(load_window(PT, u64(SA[i])).hi, cur_slot);
    bool prev_ok = false;
    if (i > 0) prev_ok = lut_slot(load_window(PT, u64(SA[i - 1])).hi, prev_slot);
    if (cur_ok && (!prev_ok || prev_slot != cur_slot)) lut_lo[cur_slot] = IdxT(i);
    if (prev_ok && (!cur_ok || prev_slot != cur_slot)) lut_hi[prev_slot] = IdxT(i);
    if (i + 1 == n1 && cur_ok) lut_hi[cur_slot] = IdxT(n1);
}

struct ChunkDev {
    u64 c0, len;        // global start, length (src/bin/asgart.rs:115)
    u64 n_probes;       // loop iterations (src/automaton.rs:96-98)
    u64 probe_base;     // global index of its first probe
    u64 needle_start;   // position of needle[0] in the packed needle image
};

// chunk of global probe g: last c with probe_base[c] <= g  (n_chunks >= 1, chunks with 0 probes are skipped naturally)
__device__ __forceinline__ u32 chunk_of_probe(const ChunkDev* __restrict__ ch, u32 n_chunks, u64 g) {
    u32 lo = 0, hi = n_chunks;
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (ch[mid].probe_base <= g) lo = mid; else hi = mid;
    }
    return lo;
}

template <typename IdxT>
struct ProbeParams {
    const u64* PT;  // packed strand
    const u64* PN;  // packed needle image (== PT when no flag is set)
    const IdxT* SA;
    const IdxT* lut_lo;
    const IdxT* lut_hi;
    const ChunkDev* chunks;
    u32 n_chunks;
    u64 n1;
    u32 k, s;
    u32 reverse;
    u64 max_card;
    u64 p_begin, p_end;  // probe shard [p_begin, p_end), p_begin % 32 == 0
    // per-probe outputs, indexed g - p_begin
    IdxT* out_lo;
    IdxT* out_raw;
    u32* out_surv;
    u32* proc_bits;   // bit (g - p_begin): iteration processed (neither N-skipped nor over the cardinality cap)
    unsigned long long* counters;  // [0] searched [1] skipped_n [2] skipped_card [3] matches [4] algorithmic bytes
};

enum { CTR_SEARCHED = 0, CTR_SKIP_N = 1, CTR_SKIP_CARD = 2, CTR_MATCHES = 3, CTR_ALG_BYTES = 4, CTR_COUNT = 8 };

__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

template <typename IdxT>
__global__ void __launch_bounds__(256) probe_search_kernel(const ProbeParams<IdxT> P) {
    const u64 g = P.p_begin + u64(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool in_range = g < P.p_end;
    bool processed = false, searched = false, skip_n = false, skip_card = false;
    u64 lo = 0, hi = 0, surv = 0, alg = 0;
    if (in_range) {
        const ChunkDev ch = P.chunks[chunk_of_probe(P.chunks, P.n_chunks, g)];
        const u64 i = (g - ch.probe_base + 1) * P.s;
        const u64 q = ch.needle_start + i;
        const int k = int(P.k);
        const Win pw_raw = load_window(P.PN, q);
        if ((pw_raw.hi >> 60) == CODE_N) {
            skip_n = true;  // src/automaton.rs:100-102
        } else {
            searched = true;
            const Win pw0 = mask_window(pw_raw, k < 32 ? k : 32);
            u32 slot = 0;
            u64 lstart = 0, rstart = 0;
            if (lut_slot(pw_raw.hi, slot)) { lstart = u64(P.lut_lo[slot]); rstart = u64(P.lut_hi[slot]); }
            const IdxT* __restrict__ sub = P.SA + lstart;
            const u64 n1 = P.n1;
            const u64* __restrict__ PT = P.PT;
            const u64* __restrict__ PN = P.PN;
            u64 r0, r1;
            equal_range_lockstep(rstart - lstart, [&](u64 ix) -> int {
                const u64 x = u64(sub[ix]);
                if (x + u64(k) > n1) return -1;  // src/searcher.rs:165-166 (Q6)
                return cmp_kmer(PT, x, PN, q, k, pw0);
            }, r0, r1);
            if (r1 < r0) r1 = r0;  // only reachable through Q6; the reference would panic on the slice
            lo = lstart + r0;
            hi = lstart + r1;
            alg = 24ull * (ceil_log2_u64(rstart - lstart + 1) + 1) + 8ull * (hi - lo);
            // filters + cardinality (src/automaton.rs:105-117)
            const bool rev = P.reverse != 0;
            for (u64 j = lo; j < hi; ++j) {
                if (match_survives(u64(P.SA[j]), i, ch.c0, ch.len, rev)) {
                    if (++surv > P.max_card) break;
                }
            }
            if (surv > P.max_card) { skip_card = true; surv = 0; } else processed = true;
        }
        const u64 o = g - P.p_begin;
        P.out_lo[o] = IdxT(lo);
        P.out_raw[o] = IdxT(hi - lo);
        P.out_surv[o] = u32(surv);
    }
    const unsigned bits = __ballot_sync(0xffffffffu, processed);
    const unsigned n_searched = __popc(__ballot_sync(0xffffffffu, searched));
    const unsigned n_skip_n = __popc(__ballot_sync(0xffffffffu, skip_n));
    const unsigned n_skip_card = __popc(__ballot_sync(0xffffffffu, skip_card));
    const u64 w_surv = warp_sum_u64(surv);
    const u64 w_alg = warp_sum_u64(alg);
    if (lane_id() == 0) {
        const u64 word = (P.p_begin + u64(blockIdx.x) * blockDim.x + (threadIdx.x & ~31u) - P.p_begin) >> 5;
        if (P.p_begin + word * 32 < P.p_end) P.proc_bits[word] = bits;
        if (n_searched) atomicAdd(&P.counters[CTR_SEARCHED], (unsigned long long)n_searched);
        if (n_skip_n) atomicAdd(&P.counters[CTR_SKIP_N], (unsigned long long)n_skip_n);
        if (n_skip_card) atomicAdd(&P.counters[CTR_SKIP_CARD], (unsigned long long)n_skip_card);
        if (w_surv) atomicAdd(&P.counters[CTR_MATCHES], (unsigned long long)w_surv);
        if (w_alg) atomicAdd(&P.counters[CTR_ALG_BYTES], (unsigned long long)w_alg);
    }
}

// pair of running sums: match offset and event index
struct Sum2 {
    u64 a, b;
    Sum2() = default;
    __host__ __device__ explicit Sum2(int) : a(0), b(0) {}
    __host__ __device__ Sum2(u64 x, u64 y) : a(x), b(y) {}
};
struct Sum2Op {
    __device__ __forceinline__ Sum2 operator()(const Sum2& x, const Sum2& y) const { return Sum2(x.a + y.a, x.b + y.b); }
};

// one warp per event: stream SA[lo, lo+raw), keep the survivors of the two filters in SA order
template <typename IdxT>
__global__ void __launch_bounds__(256) emit_matches_kernel(const IdxT* __restrict__ SA, const u64* __restrict__ ev_probe,
                                                           const u64* __restrict__ ev_moff, const IdxT* __restrict__ ev_lo,
                                                           const IdxT* __restrict__ ev_raw, u64 n_events,
                                                           const ChunkDev* __restrict__ chunks, u32 n_chunks, u32 s, u32 reverse,
                                                           u64* __restrict__ matches) {
    const u64 e = (u64(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (e >= n_events) return;
    const unsigned lane = lane_id(), lt = lanemask_lt();
    const u64 g = ev_probe[e];
    const ChunkDev ch = chunks[chunk_of_probe(chunks, n_chunks, g)];
    const u64 i = (g - ch.probe_base + 1) * s;
    const u64 b = u64(ev_lo[e]), end = b + u64(ev_raw[e]);
    u64 w = ev_moff[e];
    for (u64 j0 = b; j0 < end; j0 += 32) {
        const u64 j = j0 + lane;
        u64 x = 0;
        bool keep = false;
        if (j < end) { x = u64(SA[j]); keep = match_survives(x, i, ch.c0, ch.len, reverse != 0); }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) matches[w + __popc(m & lt)] = x;
        w += __popc(m);
    }
}

// equal range only (no filters) — test hook behind asgart_b200_ctx_probe_ranges
template <typename IdxT>
__global__ void probe_ranges_kernel(const ProbeParams<IdxT> P, i64* __restrict__ out_lo, i64* __restrict__ out_hi) {
    const u64 g = P.p_begin + u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= P.p_end) return;
    const ChunkDev ch = P.chunks[0];
    const u64 i = (g + 1) * P.s;
    const u64 q = ch.needle_start + i;
    const int k = int(P.k);
    const Win pw_raw = load_window(P.PN, q);
    const Win pw0 = mask_window(pw_raw, k < 32 ? k : 32);
    u32 slot = 0;
    u64 lstart = 0, rstart = 0;
    if (lut_slot(pw_raw.hi, slot)) { lstart = u64(P.lut_lo[slot]); rstart = u64(P.lut_hi[slot]); }
    const IdxT* __restrict__ sub = P.SA + lstart;
    const u64 n1 = P.n1;
    u64 r0, r1;
    equal_range_lockstep(rstart - lstart, [&](u64 ix) -> int {
        const u64 x = u64(sub[ix]);
        if (x + u64(k) > n1) return -1;
        return cmp_kmer(P.PT, x, P.PN, q, k, pw0);
    }, r0, r1);
    if (r1 < r0) r1 = r0;
    out_lo[g] = i64(lstart + r0);
    out_hi[g] = i64(lstart + r1);
}

}  // namespace ab200
