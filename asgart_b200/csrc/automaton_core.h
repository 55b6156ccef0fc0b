// automaton_core.h — exact, event-driven restatement of the arm automaton (src/automaton.rs:57-204 of the reference)
// for one independent *segment* of a chunk. Shared by the CUDA kernel and tests/emul/ (host build for logic tests).
//
// Reformulation (all exact):
//   * iterations skipped by the N test (:100-102) or the cardinality test (:115-117) touch nothing, so only *processed*
//     iterations count; t = index of an iteration among the processed ones of its chunk.
//   * an arm's `gap`/`active` pair is replaced by the processed index at whose end it turns inactive:
//       created at t  -> death = t + max(0, ceil(G/s) - 1)      (new arms are aged in their own iteration, :166-171)
//       extended at t -> death = t + max(1, ceil(G/s))          (dirty arms are not aged, gap reset to 0, :136-143)
//     an arm is active while matching iteration t iff death >= t.
//   * the family flush (:182-200) can only happen at an iteration without surviving matches, exactly at
//     t = max death of the family; it is applied lazily when the next event arrives (or at segment end, where
//     it is dropped if that iteration does not exist in the chunk: quirk Q3, :201-203).
//   * between two match-bearing iterations ("events") nothing else changes, so the loop runs over events only.
//   * the >200-arm prune (:173-179) only removes arms that are inactive and too short to be reported: unobservable.
#pragma once
#include <stdint.h>

#include "kmer_core.h"

namespace ab200 {

struct ArmStore {  // structure-of-arrays slices owned by the segment (capacity = matches in the segment)
    uint64_t* ls;  // left.start  (needle-local)
    uint64_t* le;  // left.end
    uint64_t* rs;  // right.start (global)
    uint64_t* re;  // right.end
    uint64_t* death;
};

struct AutoParams {
    uint64_t k, s, G;       // probe_size, step, max_gap_size (= gap + k)
    uint64_t min_len;       // min_duplication_length
    uint64_t q_ext, q_new;  // max(1, ceil(G/s)), max(0, ceil(G/s) - 1)
    uint32_t reverse;
};

// src/automaton.rs:206-216
AB_HD int64_t d_ss_core(uint64_t a_start, uint64_t a_end, uint64_t m_start, uint64_t m_end) {
    if ((m_start >= a_start && m_start <= a_end) || (m_end >= a_start && m_end <= a_end)) return 0;
    int64_t d1 = int64_t(a_start) - int64_t(m_end);
    int64_t d2 = int64_t(a_end) - int64_t(m_start);
    if (d1 < 0) d1 = -d1;
    if (d2 < 0) d2 = -d2;
    return d1 < d2 ? d1 : d2;
}

// One output duplicon, already in global coordinates (src/bin/asgart.rs:229-237)
struct SdOut {
    uint64_t left, right, left_length, right_length;
};

// Simulates the events [e0, e1) of one segment of chunk (c0, len) with Tc processed iterations in total.
//   ev_i[e]    needle-local probe position i of event e
//   ev_t[e]    processed index of that iteration inside the chunk
//   ev_moff[e], ev_cnt[e]   its surviving matches matches[moff .. moff+cnt) in SA order
//   op_target  scratch, same indexing as matches
//   arms       scratch, capacity >= total matches of the segment, indexed from 0
//   emit(sd, family_head)   called in output order
template <typename Emit>
AB_HD void simulate_segment(uint64_t e0, uint64_t e1, const uint64_t* __restrict__ ev_i, const uint64_t* __restrict__ ev_t,
                            const uint64_t* __restrict__ ev_moff, const uint32_t* __restrict__ ev_cnt,
                            const uint64_t* __restrict__ matches, int64_t* __restrict__ op_target, ArmStore arms,
                            const AutoParams& P, uint64_t c0, uint64_t len, uint64_t Tc, Emit emit) {
    uint64_t n_arms = 0, fam_start = 0, max_death = 0;

    auto flush = [&]() {
        bool first = true;
        for (uint64_t a = fam_start; a < n_arms; ++a) {
            const uint64_t rl = arms.re[a] - arms.rs[a];
            if (rl >= P.min_len) {  // :185 — right arm only
                const uint64_t ll = arms.le[a] - arms.ls[a];
                SdOut sd;
                sd.left = P.reverse ? (c0 + len - arms.ls[a] - ll) : (arms.ls[a] + c0);  // asgart.rs:229-237
                sd.right = arms.rs[a];
                sd.left_length = ll;
                sd.right_length = rl;
                emit(sd, first);
                first = false;
            }
        }
        fam_start = n_arms;
        max_death = 0;
    };

    for (uint64_t e = e0; e < e1; ++e) {
        const uint64_t t = ev_t[e], i = ev_i[e];
        if (n_arms > fam_start && max_death < t) flush();
        const uint64_t m0 = ev_moff[e], cnt = ev_cnt[e];
        const uint64_t snap = n_arms;
        // try_extend_arms against the snapshot (:122-134, :66-85)
        for (uint64_t r = 0; r < cnt; ++r) {
            const uint64_t ms = matches[m0 + r], me = ms + P.k;
            int64_t target = -1;
            for (uint64_t a = fam_start; a < snap; ++a) {
                if (arms.death[a] < t) continue;  // inactive
                const int64_t tenth = int64_t(0.1 * double(arms.le[a] - arms.ls[a]));  // :69
                const int64_t thr = int64_t(P.G) > tenth ? int64_t(P.G) : tenth;
                if (d_ss_core(arms.rs[a], arms.re[a], ms, me) < thr && me > arms.re[a]) { target = int64_t(a); break; }
            }
            op_target[m0 + r] = target;
        }
        // ExtendArm ops in match order — the last one per arm wins (:136-143)
        for (uint64_t r = 0; r < cnt; ++r) {
            const int64_t a = op_target[m0 + r];
            if (a >= 0) {
                arms.le[a] = i + P.k;
                arms.re[a] = matches[m0 + r] + P.k;
                arms.death[a] = t + P.q_ext;
                if (t + P.q_ext > max_death) max_death = t + P.q_ext;
            }
        }
        // NewArm ops in match order (:145-163)
        for (uint64_t r = 0; r < cnt; ++r) {
            if (op_target[m0 + r] < 0) {
                const uint64_t ms = matches[m0 + r];
                arms.ls[n_arms] = i; arms.le[n_arms] = i + P.k;
                arms.rs[n_arms] = ms; arms.re[n_arms] = ms + P.k;
                arms.death[n_arms] = t + P.q_new;
                if (t + P.q_new > max_death) max_death = t + P.q_new;
                ++n_arms;
            }
        }
    }
    // after the last event: the family is flushed at processed index max_death if the chunk has such an iteration
    if (n_arms > fam_start && max_death + 1 <= Tc) flush();
}

}  // namespace ab200
