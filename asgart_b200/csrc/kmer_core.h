// kmer_core.h — packed-text arithmetic shared by the CUDA kernels (and compiled for the host by tests/emul/ only,
// to unit-test the logic where no GPU exists; the product never runs it on the CPU).
//
// Packed text: 4-bit order-preserving codes, 16 per u64 word, first symbol in the most significant nibble, so that
// the lexicographic order of k-mers under the reference's byte order  $ < A < C < G < N < T  (ASCII; what
// `dna[x..x+k].cmp(pattern)` sees, src/searcher.rs:147-151,168) is the unsigned order of the packed words.
// A pure 2-bit packing cannot express this 6-symbol alphabet (N sorts between G and T and is a real symbol: with -S
// every soft-masked base is N, src/bin/asgart.rs:294-300).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AB_HD __host__ __device__ __forceinline__
#else
#define AB_HD inline
#endif

namespace ab200 {

enum : uint32_t { CODE_PAD = 0, CODE_END = 1, CODE_A = 2, CODE_C = 3, CODE_G = 4, CODE_N = 5, CODE_T = 6, CODE_BAD = 15 };

AB_HD uint32_t code_of_byte(uint8_t c) {
    switch (c) {
        case '$': return CODE_END;
        case 'A': return CODE_A;
        case 'C': return CODE_C;
        case 'G': return CODE_G;
        case 'N': return CODE_N;
        case 'T': return CODE_T;
        default: return CODE_BAD;
    }
}
// complement on codes (src/utils.rs:1-18): A<->T, C<->G, N->N
AB_HD uint32_t complement_code(uint32_t c) {
    switch (c) {
        case CODE_A: return CODE_T;
        case CODE_T: return CODE_A;
        case CODE_C: return CODE_G;
        case CODE_G: return CODE_C;
        default: return c;
    }
}

// The same two maps without branches, for the packing kernel (a switch per byte diverges five ways inside a warp and made
// packing 3.1 Gbp cost 8 ms, ten times its memory time). A byte picks one of 16 slots, slot(c) = bits 3..1 of c, plus 8 when
// bit 6 is clear: the six strand bytes land in six different slots ('A' 0, 'C' 1, 'T' 2, 'G' 3, 'N' 7, '$' 10). A slot
// holds the byte it expects (0xFF elsewhere — slot(0xFF) = 7 expects 'N', so no byte matches an unused slot) and a nibble
// of kCodeTab holds its code, of kCodeTabComp the code of its complement.
AB_HD constexpr uint32_t code_slot(uint32_t c) { return ((c >> 1) & 7u) | ((((c >> 6) & 1u) ^ 1u) << 3); }
constexpr uint64_t code_tab_build(int what, bool upper) {   // what 0: expected bytes (8 slots per word), 1: codes, 2: complement codes
    const uint8_t by[6] = {'$', 'A', 'C', 'G', 'N', 'T'};
    const uint32_t cd[6] = {CODE_END, CODE_A, CODE_C, CODE_G, CODE_N, CODE_T};
    const uint32_t cc[6] = {CODE_END, CODE_T, CODE_G, CODE_C, CODE_N, CODE_A};
    uint64_t r = what == 0 ? ~uint64_t(0) : 0;
    for (int i = 0; i < 6; ++i) {
        const uint32_t s = code_slot(by[i]);
        if (what == 0) {
            if ((s >= 8) == upper) r = (r & ~(uint64_t(0xFF) << ((s & 7u) * 8))) | (uint64_t(by[i]) << ((s & 7u) * 8));
        } else {
            r |= uint64_t(what == 1 ? cd[i] : cc[i]) << (s * 4);
        }
    }
    return r;
}
constexpr uint64_t kByteTabLo = code_tab_build(0, false), kByteTabHi = code_tab_build(0, true);
constexpr uint64_t kCodeTab = code_tab_build(1, false), kCodeTabComp = code_tab_build(2, false);
// code of strand byte c under `tab` (kCodeTab / kCodeTabComp), CODE_BAD for any other byte
AB_HD uint32_t code_of_byte_tab(uint32_t c, uint64_t tab) {
    const uint32_t s = code_slot(c);
    const uint32_t expect = uint32_t(((s & 8u) ? kByteTabHi : kByteTabLo) >> ((s & 7u) * 8)) & 0xFFu;
    return expect == c ? uint32_t(tab >> (s * 4)) & 15u : uint32_t(CODE_BAD);
}

// ---- four bytes per instruction: the packing kernel's fast path -------------------------------------------------------
// PRMT (byte permute) is a 8-entry byte table look-up for four bytes at once. slot3(c) = bits 3..1 of c tells A, C, T, G, N
// apart (0, 1, 2, 3, 7); one PRMT fetches the byte each slot expects, one the code; when the four expected bytes equal the
// four input bytes they are all in {A,C,G,N,T}. Anything else in a thread's 16 bytes — '$', a byte to reject, the ragged end
// of the text — sends that thread down the byte-by-byte path above, so the fast path never has to decide what is an error.
AB_HD uint32_t ab_prmt(uint32_t x, uint32_t y, uint32_t sel) {   // selectors 0..7 only (no sign-replicate mode)
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, y, sel);
#else
    const uint64_t t = (uint64_t(y) << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= uint32_t((t >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
#endif
}
AB_HD uint32_t ab_funnel_r(uint32_t lo, uint32_t hi, uint32_t shift) {   // low word of (hi:lo) >> (shift & 31)
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, shift);
#else
    return uint32_t(((uint64_t(hi) << 32) | lo) >> (shift & 31u));
#endif
}
constexpr uint32_t kSlotByteLo = 0x47544341u, kSlotByteHi = 0x4EFFFFFFu;   // slot 0 'A', 1 'C', 2 'T', 3 'G', 7 'N'; 0xFF has slot 7
constexpr uint32_t kSlotCodeLo = CODE_A | CODE_C << 8 | CODE_T << 16 | CODE_G << 24, kSlotCodeHi = CODE_N << 24;
constexpr uint32_t kSlotCompLo = CODE_T | CODE_G << 8 | CODE_A << 16 | CODE_C << 24;
// x: four symbols, the FIRST in the most significant byte. false: not all of them in {A,C,G,N,T}; else out16 = their four
// codes (complemented if comp), first symbol in the top nibble.
AB_HD bool pack4_fast(uint32_t x, bool comp, uint32_t& out16) {
    const uint32_t slots = (x >> 1) & 0x07070707u;
    const uint32_t sel = ab_prmt(slots | (slots >> 4), 0u, 0x4420u);          // nibble i = slot of byte i
    const uint32_t codes = ab_prmt(comp ? kSlotCompLo : kSlotCodeLo, kSlotCodeHi, sel);
    out16 = ab_prmt(codes | (codes >> 4), 0u, 0x4420u);                        // byte i -> nibble i
    return ab_prmt(kSlotByteLo, kSlotByteHi, sel) == x;
}
// One thread's 16 symbols. W[0..4]: the five aligned 32-bit words (little endian, as loaded) that contain its 16 source
// bytes, the lowest of which sits at byte m (0..3) of W[0]. Forwards the image follows the bytes, reversed it runs from the
// last byte down. false: take the byte-by-byte path.
AB_HD bool pack16_fast(const uint32_t* W, uint32_t m, bool reversed, bool comp, uint64_t& word) {
    uint32_t q[4];
    bool ok = true;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int g = 0; g < 4; ++g) {
        const int mg = reversed ? 3 - g : g;                                   // memory group that holds image symbols 4g .. 4g+3
        uint32_t x = ab_funnel_r(W[mg], W[mg + 1], 8u * m);
        if (!reversed) x = ab_prmt(x, 0u, 0x0123u);                            // first symbol to the top byte
        ok = pack4_fast(x, comp, q[g]) && ok;
    }
    word = (uint64_t(q[0]) << 48) | (uint64_t(q[1]) << 32) | (uint64_t(q[2]) << 16) | uint64_t(q[3]);
    return ok;
}

struct Win {
    uint64_t hi, lo;  // 32 symbols, first symbol in the top nibble of hi
};

// 32 symbols starting at `pos` (the packed array is padded with >= 3 zero words)
AB_HD Win load_window(const uint64_t* __restrict__ P, uint64_t pos) {
    const uint64_t w = pos >> 4;
    const unsigned o = unsigned(pos & 15u) * 4u;
    const uint64_t a = P[w], b = P[w + 1];
    Win r;
    if (o == 0) { r.hi = a; r.lo = b; }
    else {
        const uint64_t c = P[w + 2];
        r.hi = (a << o) | (b >> (64u - o));
        r.lo = (b << o) | (c >> (64u - o));
    }
    return r;
}
// keep the first k symbols (1 <= k <= 32)
AB_HD Win mask_window(Win w, int k) {
    if (k <= 16) { w.hi &= (k == 16) ? ~uint64_t(0) : ~(~uint64_t(0) >> (4 * k)); w.lo = 0; }
    else if (k < 32) { w.lo &= ~(~uint64_t(0) >> (4 * (k - 16))); }
    return w;
}
AB_HD int cmp_window(const Win& a, const Win& b) {
    if (a.hi != b.hi) return a.hi < b.hi ? -1 : 1;
    if (a.lo != b.lo) return a.lo < b.lo ? -1 : 1;
    return 0;
}

// text k-mer at x (packed PT) against needle k-mer at q (packed PN); pw0 = the needle's first window, already masked
AB_HD int cmp_kmer(const uint64_t* __restrict__ PT, uint64_t x, const uint64_t* __restrict__ PN, uint64_t q, int k, const Win& pw0) {
    const int k0 = k < 32 ? k : 32;
    int c = cmp_window(mask_window(load_window(PT, x), k0), pw0);
    if (c != 0 || k <= 32) return c;
    for (int off = 32; off < k; off += 32) {
        const int kk = (k - off) < 32 ? (k - off) : 32;
        c = cmp_window(mask_window(load_window(PT, x + off), kk), mask_window(load_window(PN, q + off), kk));
        if (c != 0) return c;
    }
    return 0;
}

// 8-mer LUT slot: the first 8 symbols read as a base-5 number, digits A=0,C=1,G=2,N=3,T=4 (lexicographic, so slots are
// in SA order). Returns false when any of the 8 symbols is '$' or padding (such suffixes are in no bucket, like the
// reference's 5^8 enumeration, src/searcher.rs:99-117).
constexpr uint32_t kLutSize = 390625;  // 5^8
AB_HD bool lut_slot(uint64_t win_hi, uint32_t& slot) {
    uint32_t s = 0;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t nib = uint32_t(win_hi >> (60 - 4 * j)) & 15u;
        ok = ok && (nib >= CODE_A && nib <= CODE_T);
        s = s * 5u + (nib - CODE_A);
    }
    slot = s;
    return ok;
}

// deep table slot: the first `depth` symbols (depth <= 16) read as a base-4 number, digits A,C,G,T = 0..3 (lexicographic).
// Returns false when any of them is N, '$' or padding: such probes take the reference's own 8-mer bucket.
AB_HD bool deep_slot(uint64_t win_hi, int depth, uint32_t& slot) {
    uint32_t s = 0;
    bool ok = true;
    for (int j = 0; j < depth; ++j) {
        const uint32_t nib = uint32_t(win_hi >> (60 - 4 * j)) & 15u;
        ok = ok && (nib == CODE_A || nib == CODE_C || nib == CODE_G || nib == CODE_T);
        s = s * 4u + (nib == CODE_T ? 3u : nib - CODE_A);
    }
    slot = s;
    return ok;
}

// superslice::Ext::equal_range_by restated (third-party crate "1.0", call site src/searcher.rs:164): lower and upper
// bound in lock step over `len` slots, `size -= half`, one final probe each. f(ix) -> -1/0/+1.
// Kept literal because the reference's comparator is non-monotone near the end of the strand (quirk Q6,
// src/searcher.rs:165-166), where the result depends on the exact probe sequence.
template <typename F>
AB_HD void equal_range_lockstep(uint64_t len, F f, uint64_t& r0, uint64_t& r1) {
    if (len == 0) { r0 = r1 = 0; return; }
    uint64_t size = len, b0 = 0, b1 = 0;
    while (size > 1) {
        const uint64_t half = size >> 1;
        const uint64_t m0 = b0 + half, m1 = b1 + half;
        const int c0 = f(m0);
        const int c1 = (m1 == m0) ? c0 : f(m1);
        if (c0 < 0) b0 = m0;
        if (c1 <= 0) b1 = m1;
        size -= half;
    }
    const int c0 = f(b0);
    const int c1 = (b1 == b0) ? c0 : f(b1);
    r0 = b0 + (c0 < 0 ? 1u : 0u);
    r1 = b1 + (c1 <= 0 ? 1u : 0u);
}

// the two match filters of src/automaton.rs:106-113 (i is needle-local: quirk Q1)
AB_HD bool match_survives(uint64_t m_start, uint64_t i, uint64_t c0, uint64_t len, bool reverse) {
    if (m_start == i) return false;
    return reverse ? (m_start >= c0 + len - i) : (m_start > i + c0);
}

// number of loop iterations of src/automaton.rs:96-98 for a chunk of length len (0 when the reference returns early,
// :92-94, or when len < k + s, where the reference's usize subtraction would underflow)
AB_HD uint64_t probes_in_chunk(uint64_t len, uint64_t k, uint64_t s, uint64_t min_len) {
    if (len < min_len || len < k + s || s == 0) return 0;
    const uint64_t limit = len - k - s;
    return (limit + s - 1) / s;
}

AB_HD uint32_t ceil_log2_u64(uint64_t v) {  // ceil(log2(v)), v >= 1
    uint32_t l = 0;
    while ((uint64_t(1) << l) < v && l < 63) ++l;
    return l;
}


// ---- the reference's 8-mer bucket search, probe for probe (row N4, --trim; libdivsufsort/lib/utils.c:282-349 sa_search as
// called by Searcher::new, src/searcher.rs:118-128) ------------------------------------------------------------------
// With --trim the array is not sorted for the text it is compared with near the trim end, so the result is
// defined by WHICH suffixes the reference looks at, not by an order: every descent below visits the same middles and
// skips the same already-matched symbols as the reference does. One descent = repeated halving of SA[lo, lo+len):
// the middle suffix is compared with P from the first symbol that the two sides are not yet known to share; going right
// keeps (middle, end), going left keeps [lo, middle). Three descents make a search: to the first suffix that equals P
// (STOP_AT_EQUAL), then a lower bound over what was left of it and an upper bound over what was right of it.
enum LiteralMode { STOP_AT_EQUAL = 0, LOWER_BOUND = 1, UPPER_BOUND = 2 };

// sign of (suffix at `suf`) - P over the 8 symbols of P, reading from symbol `known` on; `known` becomes the number of
// leading symbols found equal. A suffix that ends inside P is smaller.
AB_HD int literal_cmp8(const uint8_t* T, int64_t tsize, const uint8_t* P, int64_t suf, int64_t& known) {
    int64_t m = known;
    int diff = 0;
    for (; m < 8 && suf + m < tsize; ++m) {
        diff = int(T[suf + m]) - int(P[m]);
        if (diff) break;
    }
    known = m;
    return diff ? diff : (m < 8 ? -1 : 0);
}

// returns true when MODE == STOP_AT_EQUAL met an equal suffix: [lo, lo+len) is then the interval that was being halved,
// its middle lo + len/2 is that suffix and known_mid the symbols matched there (8)
template <int MODE, typename IdxT>
AB_HD bool literal_descent(const uint8_t* T, int64_t tsize, const uint8_t* P, const IdxT* SA, int64_t& lo, int64_t& len,
                           int64_t& known_lo, int64_t& known_hi, int64_t& known_mid) {
    while (len > 0) {
        const int64_t mid = len >> 1;
        int64_t m = known_lo < known_hi ? known_lo : known_hi;
        const int r = literal_cmp8(T, tsize, P, int64_t(SA[lo + mid]), m);
        if (MODE == STOP_AT_EQUAL && r == 0) { known_mid = m; return true; }
        const bool right = (MODE == UPPER_BOUND) ? r <= 0 : r < 0;
        if (right) { lo += mid + 1; len -= mid + 1; known_lo = m; }
        else { len = mid; known_hi = m; }
    }
    return false;
}

// (first, count) as sa_search reports them: count suffixes from SA[first] on start with P; count == 0: first = where the
// first descent ended (the reference's insertion point)
template <typename IdxT>
AB_HD void literal_bucket(const uint8_t* T, int64_t tsize, const uint8_t* P, const IdxT* SA, int64_t sa_size, int64_t& first,
                          int64_t& count) {
    int64_t lo = 0, len = sa_size, k_lo = 0, k_hi = 0, k_mid = 0;
    if (!literal_descent<STOP_AT_EQUAL, IdxT>(T, tsize, P, SA, lo, len, k_lo, k_hi, k_mid)) { first = lo; count = 0; return; }
    const int64_t mid = len >> 1;
    int64_t a = lo, a_len = mid, a_lo = k_lo, a_hi = k_mid, unused = 0;
    literal_descent<LOWER_BOUND, IdxT>(T, tsize, P, SA, a, a_len, a_lo, a_hi, unused);
    int64_t b = lo + mid + 1, b_len = len - mid - 1, b_lo = k_mid, b_hi = k_hi;
    literal_descent<UPPER_BOUND, IdxT>(T, tsize, P, SA, b, b_len, b_lo, b_hi, unused);
    first = a;
    count = b - a;
}

}  // namespace ab200
