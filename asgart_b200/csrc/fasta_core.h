// fasta_core.h — the arithmetic of the GPU-side FASTA ingest that needs no CUDA: byte classification four bytes per
// instruction, the keep / base masks of one thread's 32 bytes, the per-thread counts, and the chunk derivation from the
// long N-runs. Shared by the kernels of fasta_ingest.cuh and (compiled for the host by tests/emul/ only) by the CPU tests,
// which run the three passes one "thread" at a time against the oracle's prepare_data; the product never runs it on the CPU.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/asgart_b200.h"

#if !defined(AB_HD)
#if defined(__CUDACC__)
#define AB_HD __host__ __device__ __forceinline__
#else
#define AB_HD inline
#endif
#endif

namespace ab200 {

#if defined(__CUDA_ARCH__)
AB_HD int fc_clz(uint32_t x) { return __clz(int(x)); }
AB_HD int fc_ffs(uint32_t x) { return __ffs(int(x)); }
AB_HD int fc_popc(uint32_t x) { return __popc(x); }
#else
AB_HD int fc_clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
AB_HD int fc_ffs(uint32_t x) { return __builtin_ffs(int(x)); }
AB_HD int fc_popc(uint32_t x) { return __builtin_popcount(x); }
#endif

constexpr uint64_t kLongNRun = 5000;   // src/bin/asgart.rs:326

AB_HD bool fa_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }   // ASCII isspace
AB_HD bool fa_blank(uint8_t c) { return fa_space(c) && c != '\n'; }

// src/bin/asgart.rs:291-301
AB_HD uint8_t fa_normalise(uint8_t c, bool skip_masked) {
    if (!skip_masked && c >= 'a' && c <= 'z') c = uint8_t(c - 32);
    return (c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N') ? c : uint8_t('N');
}

constexpr int kFiThreads = 256;
constexpr int kFiBytes = 32;
constexpr int kFiWarps = kFiThreads / 32;
constexpr uint64_t kFiTile = uint64_t(kFiThreads) * kFiBytes;
enum : uint32_t { FI_NONE = 0, FI_CLR = 1, FI_SET = 2 };

// ---- byte classification, four bytes per instruction (SWAR; exact for every byte value, no carries between bytes) ----
// flags: 0x80 in every byte of x that is zero
AB_HD uint32_t sw_zero(uint32_t x) { return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }
AB_HD uint32_t sw_eq(uint32_t w, uint32_t k4) { return sw_zero(w ^ k4); }
// flags: 0x80 in every byte of w below 0x21 (control characters and the blank: every white-space byte is among them)
AB_HD uint32_t sw_below_21(uint32_t w) { return ~(((w & 0x7F7F7F7Fu) + 0x5F5F5F5Fu) | w) & 0x80808080u; }
// the four flag bits of a word as bits 0..3 (byte 0 = bit 0)
AB_HD uint32_t sw_movemask(uint32_t f) { return ((f >> 7) * 0x01020408u) >> 24; }

struct FiThread {
    uint32_t w[kFiBytes / 4];   // the thread's 32 bytes, little endian: byte j = w[j / 4] >> 8 * (j % 4)
    int valid;                  // bytes of this thread inside the file
    uint64_t base;
    uint32_t vmask;             // bit j: byte j is inside the file
    uint32_t nl, space, heads;  // bit j: byte j is '\n' / white space / a '>' that opens a line
    uint32_t h_last, r_first;   // this thread's last header event / first rest event
};

// masks and events of a thread whose words, valid and vmask are set; prev_nl: the byte before it is '\n' (or the file starts)
AB_HD void fi_classify(FiThread& t, uint32_t prev_nl) {
    uint32_t nl = 0, gt = 0, low = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < kFiBytes / 4; ++q) {
        nl |= sw_movemask(sw_eq(t.w[q], 0x0A0A0A0Au)) << (4 * q);
        gt |= sw_movemask(sw_eq(t.w[q], 0x3E3E3E3Eu)) << (4 * q);
        low |= sw_movemask(sw_below_21(t.w[q])) << (4 * q);
    }
    nl &= t.vmask; gt &= t.vmask; low &= t.vmask;
    uint32_t space = nl;
    if (low != nl) {                       // some other control character or blank: the remaining white-space values
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < kFiBytes / 4; ++q) {
            const uint32_t x = t.w[q];
            const uint32_t f = sw_eq(x, 0x20202020u) | sw_eq(x, 0x09090909u) | sw_eq(x, 0x0B0B0B0Bu) | sw_eq(x, 0x0C0C0C0Cu) | sw_eq(x, 0x0D0D0D0Du);
            space |= sw_movemask(f) << (4 * q);
        }
        space &= t.vmask;
    }
    t.nl = nl;
    t.space = space;
    t.heads = gt & ((nl << 1) | prev_nl);
    const uint32_t ev = t.heads | nl;            // header machine: '>' at a line start sets, '\n' clears; the last event counts
    t.h_last = ev ? (((t.heads >> (31 - fc_clz(ev))) & 1u) ? uint32_t(FI_SET) : uint32_t(FI_CLR)) : uint32_t(FI_NONE);
    const uint32_t rv = nl | (~space & t.vmask);  // rest machine: '\n' sets, a non-white byte clears; the first event counts
    t.r_first = rv ? (((nl >> (fc_ffs(rv) - 1)) & 1u) ? uint32_t(FI_SET) : uint32_t(FI_CLR)) : uint32_t(FI_NONE);
}

// keep mask of this thread's bytes: not on a header line, not '\n', not trailing white space
AB_HD uint32_t fi_keep_mask(const FiThread& t, uint32_t h_in, uint32_t r_in) {
    if (t.valid == 0) return 0;
    uint32_t on_header;
    if (t.heads == 0) {                     // no header starts here: the entering state holds up to the first '\n'
        on_header = h_in ? (t.nl ? ((1u << (fc_ffs(t.nl) - 1)) - 1u) : 0xffffffffu) : 0u;
    } else {
        on_header = 0;
        uint32_t h = h_in;
        for (int j = 0; j < t.valid; ++j) {
            if ((t.nl >> j) & 1u) h = 0;
            else if ((t.heads >> j) & 1u) h = 1;
            on_header |= h << j;
        }
    }
    const uint32_t blank = t.space & ~t.nl;
    uint32_t trailing = 0;
    if (blank) {                            // a blank is trailing iff its successor is '\n', a trailing blank, or (last byte) r_in
        const uint32_t inject = r_in << (t.valid - 1);
        for (;;) {
            const uint32_t next = blank & (((t.nl | trailing) >> 1) | inject);
            if (next == trailing) break;
            trailing = next;
        }
    }
    return ~(on_header | t.nl | trailing) & t.vmask;
}

// flags of the bytes that normalise to a base (A, C, G, T; src/bin/asgart.rs:291-301) and the normalised words
AB_HD uint32_t fi_base_mask(const FiThread& t, bool skip_masked, uint32_t* norm /* kFiBytes / 4, may be null */) {
    uint32_t base = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < kFiBytes / 4; ++q) {
        // upper-casing: clearing bit 5 maps a,c,g,t onto A,C,G,T and no other byte value onto them
        const uint32_t u = skip_masked ? t.w[q] : (t.w[q] & 0xDFDFDFDFu);
        const uint32_t f = sw_eq(u, 0x41414141u) | sw_eq(u, 0x43434343u) | sw_eq(u, 0x47474747u) | sw_eq(u, 0x54545454u);
        base |= sw_movemask(f) << (4 * q);
        if (norm) { const uint32_t bm = (f >> 7) * 0xFFu; norm[q] = (u & bm) | (0x4E4E4E4Eu & ~bm); }
    }
    return base & t.vmask;
}

// kept | header starts << 32 of this thread, and the number of its kept bytes up to and including the last base
AB_HD void fi_thread_counts(const FiThread& t, uint32_t keep, uint32_t base, uint64_t& packed, uint32_t& upto_base) {
    packed = uint64_t(fc_popc(keep)) | (uint64_t(fc_popc(t.heads)) << 32);
    const uint32_t kb = keep & base;
    upto_base = kb ? uint32_t(fc_popc(keep & (0xffffffffu >> fc_clz(kb)))) : 0u;
}

// chunks_to_process of one file (src/bin/asgart.rs:317-366 applied per fragment, :381-387) from its records and its maximal
// N-runs longer than kLongNRun (ascending, fragment borders ignored): inside each fragment the maximal regions between such
// runs; a run that crosses a fragment border counts on each side with the part that lies there. rec_pos = strand position
// (file-local) of every record's first base, kept = bases of the file, off = strand position of the file's first base.
inline void chunks_from_runs(const std::vector<uint64_t>& rec_pos, uint64_t kept, const std::vector<uint64_t>& run_start,
                             const std::vector<uint64_t>& run_len, uint64_t off, std::vector<asgart_b200_chunk>& chunks,
                             std::vector<uint64_t>& frag_pos, std::vector<uint64_t>& frag_len) {
    size_t ri = 0;
    const size_t R = rec_pos.size();
    for (size_t r = 0; r < R; ++r) {
        const uint64_t fs = rec_pos[r], fe = r + 1 < R ? rec_pos[r + 1] : kept;
        frag_pos.push_back(off + fs);
        frag_len.push_back(fe - fs);
        const size_t first = chunks.size();
        uint64_t cur = fs;
        while (ri < run_start.size() && run_start[ri] + run_len[ri] <= fs) ++ri;
        for (size_t j = ri; j < run_start.size() && run_start[j] < fe; ++j) {
            const uint64_t ps = std::max(run_start[j], fs), pe = std::min(run_start[j] + run_len[j], fe);
            if (pe - ps <= kLongNRun) continue;
            if (ps > cur) chunks.push_back({off + cur, ps - cur});
            cur = pe;
        }
        if (fe > cur) chunks.push_back({off + cur, fe - cur});
        if (chunks.size() == first) chunks.push_back({off + fs, fe - fs});
    }
}

// What the bio FASTA reader (crate `bio`, io/fasta.rs; the reference's Cargo.toml takes any version and ships no lock file)
// does besides splitting records, as seen by `for record in reader.records()` at src/bin/asgart.rs:286:
//  * Reader::read returns "Expected > at record start." unless the FIRST byte of a non-empty file is '>' — a leading blank
//    line is a parse error (first_byte_ok);
//  * Records::next ends the iteration at the first EMPTY record — no id, no description, no sequence (Record::is_empty) — and
//    never looks at the rest of the file (stop_at_empty_record). Such a record is a header line with nothing but white space
//    after the '>' whose sequence lines, if any, are all blank.
inline bool first_byte_ok(uint64_t n, uint8_t first) { return n == 0 || first == '>'; }

// blank_header(r): only white space between record r's '>' and the end of that line. Cuts rec_off / rec_pos / kept at the
// first empty record and returns how many records remain.
template <class BlankHeader>
inline size_t stop_at_empty_record(std::vector<uint64_t>& rec_off, std::vector<uint64_t>& rec_pos, uint64_t& kept,
                                   BlankHeader blank_header) {
    const size_t R = rec_pos.size();
    for (size_t r = 0; r < R; ++r) {
        const uint64_t end = r + 1 < R ? rec_pos[r + 1] : kept;
        if (end == rec_pos[r] && blank_header(r)) {
            kept = rec_pos[r];
            rec_off.resize(r);
            rec_pos.resize(r);
            return r;
        }
    }
    return R;
}

}  // namespace ab200
