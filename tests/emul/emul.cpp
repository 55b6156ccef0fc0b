// tests/emul/emul.cpp — TEST-ONLY host build of the device logic headers (kmer_core.h, automaton_core.h).
// There is no GPU in the development container, so the arithmetic the kernels share with these headers (packed
// windows, the comparator, the literal lock-step equal range, the LUT slot, the match filters, the event/segment/death
// reformulation of the automaton) is exercised here on the CPU against the oracle. The glue below restates what the
// kernels in search.cuh / automaton.cuh do around those functions, one "thread" at a time. It is never shipped and the
// product never calls it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../asgart_b200/csrc/automaton_core.h"
#include "../../asgart_b200/csrc/kmer_core.h"
#include "../../include/asgart_b200.h"

using namespace ab200;
using u64 = uint64_t;
using u32 = uint32_t;

namespace {
struct Chunk { u64 c0, len, n_probes, probe_base, needle_start; };

std::vector<u64> pack(const uint8_t* text, u64 n1, int mode) {  // pack_text_kernel, one word per iteration
    const u64 n = n1 - 1, words = (n1 + 15) / 16 + 4, limit = mode == 0 ? n1 : n;
    std::vector<u64> P(words, 0);
    for (u64 w = 0; w < words; ++w) {
        u64 word = 0;
        for (int j = 0; j < 16; ++j) {
            const u64 p = w * 16 + j;
            u32 c = CODE_PAD;
            if (p < limit) {
                const u64 src = (mode & 2) ? (n - 1 - p) : p;
                c = code_of_byte(text[src]);
                if (mode & 1) c = complement_code(c);
            }
            word |= u64(c & 15u) << (60 - 4 * j);
        }
        P[w] = word;
    }
    return P;
}

struct Result { std::vector<int64_t> fam_off{0}; std::vector<u64> fields; };
}  // namespace

extern "C" {

// windows / comparator unit hooks
void emul_window(const uint8_t* text, int64_t n1, int mode, uint64_t pos, int k, uint64_t* hi, uint64_t* lo) {
    auto P = pack(text, u64(n1), mode);
    Win w = mask_window(load_window(P.data(), pos), k);
    *hi = w.hi; *lo = w.lo;
}

int emul_lut(const uint8_t* text, int64_t n1, const int64_t* SA, int64_t* lo, int64_t* hi) {  // lut_build_kernel
    auto P = pack(text, u64(n1), 0);
    for (u32 s = 0; s < kLutSize; ++s) lo[s] = hi[s] = 0;
    for (u64 i = 0; i < u64(n1); ++i) {
        u32 cur = 0, prev = 0;
        const bool cur_ok = lut_slot(load_window(P.data(), u64(SA[i])).hi, cur);
        bool prev_ok = false;
        if (i > 0) prev_ok = lut_slot(load_window(P.data(), u64(SA[i - 1])).hi, prev);
        if (cur_ok && (!prev_ok || prev != cur)) lo[cur] = int64_t(i);
        if (prev_ok && (!cur_ok || prev != cur)) hi[prev] = int64_t(i);
        if (i + 1 == u64(n1) && cur_ok) hi[cur] = n1;
    }
    return 0;
}

void* emul_search(const uint8_t* text, int64_t n1_, const int64_t* SA, const asgart_b200_chunk* chunks, int64_t n_chunks,
                  const asgart_b200_settings* st, uint64_t* counters /* probes, searched, skip_n, skip_card, matches, alg */) {
    const u64 n1 = u64(n1_), n = n1 - 1, k = st->probe_size, s = k / 2;
    const int mode = (st->reverse ? 2 : 0) | (st->complement ? 1 : 0);
    auto PT = pack(text, n1, 0);
    auto PN = mode ? pack(text, n1, mode) : PT;
    std::vector<int64_t> lut_lo(kLutSize), lut_hi(kLutSize);
    emul_lut(text, n1_, SA, lut_lo.data(), lut_hi.data());
    std::vector<Chunk> ch;
    u64 base = 0;
    for (int64_t c = 0; c < n_chunks; ++c) {
        Chunk d{chunks[c].start, chunks[c].length, probes_in_chunk(chunks[c].length, k, s, st->min_duplication_length), base,
                st->reverse ? (n - chunks[c].start - chunks[c].length) : chunks[c].start};
        base += d.n_probes;
        ch.push_back(d);
    }
    const u64 total = base;
    // stage A
    std::vector<uint8_t> processed(total, 0);
    std::vector<u64> ev_probe, ev_moff, matches;
    std::vector<u32> ev_cnt;
    u64 ctr[6] = {total, 0, 0, 0, 0, 0};
    for (u64 g = 0; g < total; ++g) {
        size_t c = 0;
        while (c + 1 < ch.size() && ch[c + 1].probe_base <= g) ++c;
        const Chunk& cd = ch[c];
        const u64 i = (g - cd.probe_base + 1) * s, q = cd.needle_start + i;
        const Win pw_raw = load_window(PN.data(), q);
        if ((pw_raw.hi >> 60) == CODE_N) { ++ctr[2]; continue; }
        ++ctr[1];
        const Win pw0 = mask_window(pw_raw, k < 32 ? int(k) : 32);
        u32 slot = 0;
        u64 lstart = 0, rstart = 0;
        if (lut_slot(pw_raw.hi, slot)) { lstart = u64(lut_lo[slot]); rstart = u64(lut_hi[slot]); }
        u64 r0, r1;
        equal_range_lockstep(rstart - lstart, [&](u64 ix) -> int {
            const u64 x = u64(SA[lstart + ix]);
            if (x + k > n1) return -1;
            return cmp_kmer(PT.data(), x, PN.data(), q, int(k), pw0);
        }, r0, r1);
        if (r1 < r0) r1 = r0;
        ctr[5] += 24ull * (ceil_log2_u64(rstart - lstart + 1) + 1) + 8ull * (r1 - r0);
        std::vector<u64> surv;
        for (u64 j = lstart + r0; j < lstart + r1; ++j)
            if (match_survives(u64(SA[j]), i, cd.c0, cd.len, st->reverse != 0)) surv.push_back(u64(SA[j]));
        if (surv.size() > st->max_cardinality) { ++ctr[3]; continue; }
        processed[g] = 1;
        if (!surv.empty()) {
            ev_probe.push_back(g); ev_cnt.push_back(u32(surv.size())); ev_moff.push_back(matches.size());
            matches.insert(matches.end(), surv.begin(), surv.end());
            ctr[4] += surv.size();
        }
    }
    if (counters) memcpy(counters, ctr, sizeof ctr);
    // stage B bookkeeping (event_info_kernel / chunk_tc_kernel)
    std::vector<u64> pre(total + 1, 0);
    for (u64 g = 0; g < total; ++g) pre[g + 1] = pre[g] + processed[g];
    AutoParams P{};
    P.k = k; P.s = s; P.G = st->max_gap_size; P.min_len = st->min_duplication_length;
    const u64 qq = (P.G + P.s - 1) / P.s;
    P.q_ext = qq > 1 ? qq : 1; P.q_new = qq > 0 ? qq - 1 : 0; P.reverse = st->reverse ? 1 : 0;
    const u64 ne = ev_probe.size();
    std::vector<u64> ev_i(ne), ev_t(ne);
    std::vector<u32> ev_chunk(ne);
    std::vector<u64> seg_first;
    for (u64 e = 0; e < ne; ++e) {
        const u64 g = ev_probe[e];
        size_t c = 0;
        while (c + 1 < ch.size() && ch[c + 1].probe_base <= g) ++c;
        ev_chunk[e] = u32(c);
        ev_i[e] = (g - ch[c].probe_base + 1) * s;
        ev_t[e] = pre[g] - pre[ch[c].probe_base];
        bool head = true;
        if (e > 0 && ev_chunk[e - 1] == c) head = (ev_t[e] - ev_t[e - 1]) > P.q_ext;
        if (head) seg_first.push_back(e);
    }
    seg_first.push_back(ne);
    // automaton, one segment at a time (automaton_kernel)
    Result* R = new Result();
    const u64 nm = matches.size();
    std::vector<int64_t> op(nm);
    std::vector<u64> a_ls(nm), a_le(nm), a_rs(nm), a_re(nm), a_death(nm);
    for (size_t sidx = 0; sidx + 1 < seg_first.size(); ++sidx) {
        const u64 e0 = seg_first[sidx], e1 = seg_first[sidx + 1];
        const Chunk& cd = ch[ev_chunk[e0]];
        const u64 slot0 = ev_moff[e0];
        ArmStore arms{a_ls.data() + slot0, a_le.data() + slot0, a_rs.data() + slot0, a_re.data() + slot0, a_death.data() + slot0};
        const u64 Tc = pre[cd.probe_base + cd.n_probes] - pre[cd.probe_base];
        simulate_segment(e0, e1, ev_i.data(), ev_t.data(), ev_moff.data(), ev_cnt.data(), matches.data(), op.data(), arms, P, cd.c0,
                         cd.len, Tc, [&](const SdOut& sd, bool head) {
                             if (head && !R->fields.empty()) R->fam_off.push_back(int64_t(R->fields.size() / 4));
                             R->fields.push_back(sd.left); R->fields.push_back(sd.right);
                             R->fields.push_back(sd.left_length); R->fields.push_back(sd.right_length);
                         });
    }
    if (!R->fields.empty()) R->fam_off.push_back(int64_t(R->fields.size() / 4));
    return R;
}
int64_t emul_result_n_families(void* h) { return int64_t(static_cast<Result*>(h)->fam_off.size()) - 1; }
int64_t emul_result_n_sds(void* h) { return int64_t(static_cast<Result*>(h)->fields.size() / 4); }
void emul_result_copy(void* h, int64_t* fam_off, uint64_t* fields) {
    Result* R = static_cast<Result*>(h);
    memcpy(fam_off, R->fam_off.data(), R->fam_off.size() * 8);
    if (!R->fields.empty()) memcpy(fields, R->fields.data(), R->fields.size() * 8);
}
void emul_result_free(void* h) { delete static_cast<Result*>(h); }
}
