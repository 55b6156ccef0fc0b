// oracle/asgart_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY — NOT PART OF THE PRODUCT.
// A plain, sequential CPU restatement of the duplication-search hot path of delehef/asgart
// (reference @ 523b07c). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may load this library, and only as the checker or the timed CPU baseline.
// The product (asgart_b200/) never links or calls it.
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
// It is written fresh: same behaviour, including the output-altering quirks Q1–Q6 listed in
// SURVEY.md §8a, but none of the reference's source text.
//
// PARITY PINNING
//   * Suffix array and the 8-mer LUT are pinned against the reference's real C library
//     (oracle/_ref/libdivsufsort64.so, compiled from /root/reference/libdivsufsort by oracle/Makefile):
//     tests/test_oracle_ref.py compares oracle_suffix_array / oracle_lut with divsufsort64 /
//     sa_searchb64 / sufcheck64.
//   * The Rust parts (searcher.rs, automaton.rs, bin/asgart.rs post-steps, JSON) have NO reference tests,
//     fixtures or golden vectors, and no Rust toolchain exists here to run them: for those rows the
//     oracle is "parity unpinned" — it is checked only against the hand-derived known-answer cases
//     of SURVEY.md §8c (tests/test_oracle_kat.py).
//   * Third-party arithmetic not under /root/reference: superslice (Cargo.toml:31, "1.0", unpinned —
//     no Cargo.lock): equal_range_by is restated from its published algorithm (lock-step branch-free
//     bisection), call site searcher.rs:164. bio (Cargo.toml:13, "*"): FASTA reader semantics restated
//     (id = header up to first whitespace, sequence lines trimmed and concatenated), call site
//     bin/asgart.rs:282-290.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <mutex>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

using usize = size_t;

// ---------------------------------------------------------------------------------------------
// structs.rs:10-11  alphabets; utils.rs:1-18 complement table
// ---------------------------------------------------------------------------------------------
const uint8_t ALPHABET[5] = {'A', 'T', 'G', 'C', 'N'};         // structs.rs:10 (this enumeration order)
const uint8_t ALPHABET_MASKED[5] = {'a', 't', 'g', 'c', 'n'};  // structs.rs:11

bool in_set(const uint8_t* set, uint8_t c) {
    for (int i = 0; i < 5; ++i)
        if (set[i] == c) return true;
    return false;
}

// utils.rs:1-18
uint8_t complement_nucleotide(uint8_t n) {
    switch (n) {
        case 'A': return 'T';
        case 'T': return 'A';
        case 'G': return 'C';
        case 'C': return 'G';
        case 'N': return 'N';
        case 'a': return 't';
        case 't': return 'a';
        case 'g': return 'c';
        case 'c': return 'g';
        case 'n': return 'n';
        default: return 'N';
    }
}

// ---------------------------------------------------------------------------------------------
// structs.rs:36-58 RunSettings, :60-65 Start, :418-429 ProtoSD; automaton.rs:10-41 Segment/Arm
// ---------------------------------------------------------------------------------------------
struct RunSettings {
    usize probe_size;
    uint32_t max_gap_size;  // already gap_size + probe_size (bin/asgart.rs:681)
    usize min_duplication_length;
    usize max_cardinality;
    bool has_trim;
    usize trim_a, trim_b;
    bool reverse, complement, skip_masked;
};

struct Start {
    std::string name;
    usize position, length;
};

struct ProtoSD {
    usize left, right, left_length, right_length;
    float identity;
    bool reversed, complemented;
};
using Family = std::vector<ProtoSD>;

struct Segment {
    usize tag, start, end;
    usize len() const { return end - start; }
};

struct Arm {
    Segment left, right;
    bool active, dirty;
    usize gap;
};

// ---------------------------------------------------------------------------------------------
// libdivsufsort/lib/utils.c:244-255 (_compare) and :282-349 (sa_search) — semantics restated:
// the suffix is compared with P over at most Psize characters; a suffix that ends before a
// difference is found is smaller than P. Returns (first index, count); when count == 0 the index is
// the insertion point (the value `i` holds when the C loop ends, utils.c:347).
// sa_searchb (utils.c:258-280) is sa_search on SA[start_left .. start_right).
// ---------------------------------------------------------------------------------------------
int cmp_suffix_pattern(const uint8_t* T, int64_t Tsize, const uint8_t* P, int64_t Psize, int64_t suf) {
    int64_t i = suf, j = 0;
    while (i < Tsize && j < Psize && T[i] == P[j]) { ++i; ++j; }
    if (j == Psize) return 0;
    if (i >= Tsize) return -1;
    return int(T[i]) - int(P[j]);
}

int64_t sa_searchb_restated(const uint8_t* T, int64_t Tsize, const uint8_t* P, int64_t Psize, const int64_t* SA,
                            int64_t start_left, int64_t start_right, int64_t* idx) {
    int64_t lo = start_left, hi = start_right;  // first suffix >= P
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if (cmp_suffix_pattern(T, Tsize, P, Psize, SA[mid]) < 0) lo = mid + 1; else hi = mid;
    }
    int64_t first = lo;
    hi = start_right;  // first suffix > P
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if (cmp_suffix_pattern(T, Tsize, P, Psize, SA[mid]) <= 0) lo = mid + 1; else hi = mid;
    }
    *idx = first;
    return lo - first;
}

// ---------------------------------------------------------------------------------------------
// The same search, step for step (utils.c:244-255 _compare, :282-349 sa_search): one bisection until a suffix matches P,
// then a lower-bound bisection of the left part and an upper-bound bisection of the right part, every comparison
// skipping the prefix already known to match (`match` counters). Identical to the restatement above whenever SA is
// sorted for T; it differs when it is not — `--trim` hands a suffix array of strand[a..b]+'$' (shifted by a) together
// with the WHOLE strand (bin/asgart.rs:142-155), so the last suffixes before b are compared beyond the
// '$' they were sorted with. Checked against the reference's own sa_searchb64 (tests/test_oracle_ref.py).
// ---------------------------------------------------------------------------------------------
int compare_skip(const uint8_t* T, int64_t Tsize, const uint8_t* P, int64_t Psize, int64_t suf, int64_t* match) {
    int64_t i = suf + *match, j = *match;
    int r = 0;
    while (i < Tsize && j < Psize) {
        r = int(T[i]) - int(P[j]);
        if (r != 0) break;
        ++i; ++j;
    }
    *match = j;
    if (r != 0) return r;
    return j != Psize ? -1 : 0;
}

int64_t sa_search_literal(const uint8_t* T, int64_t Tsize, const uint8_t* P, int64_t Psize, const int64_t* SA,
                          int64_t SAsize, int64_t* idx) {
    *idx = -1;
    if (Tsize == 0 || SAsize == 0) return 0;
    if (Psize == 0) { *idx = 0; return SAsize; }
    int64_t i = 0, j = 0, k = 0, lmatch = 0, rmatch = 0;
    int64_t size = SAsize, half = size >> 1;
    while (size > 0) {
        int64_t match = std::min(lmatch, rmatch);
        const int r = compare_skip(T, Tsize, P, Psize, SA[i + half], &match);
        if (r < 0) {
            i += half + 1;
            half -= (size & 1) ^ 1;
            lmatch = match;
        } else if (r > 0) {
            rmatch = match;
        } else {
            int64_t lsize = half, rsize = size - half - 1;
            j = i; k = i + half + 1;
            int64_t llmatch = lmatch, lrmatch = match;
            for (int64_t h = lsize >> 1; lsize > 0; lsize = h, h >>= 1) {       // first suffix >= P in the left part
                int64_t m = std::min(llmatch, lrmatch);
                if (compare_skip(T, Tsize, P, Psize, SA[j + h], &m) < 0) { j += h + 1; h -= (lsize & 1) ^ 1; llmatch = m; }
                else lrmatch = m;
            }
            int64_t rlmatch = match, rrmatch = rmatch;
            for (int64_t h = rsize >> 1; rsize > 0; rsize = h, h >>= 1) {       // first suffix > P in the right part
                int64_t m = std::min(rlmatch, rrmatch);
                if (compare_skip(T, Tsize, P, Psize, SA[k + h], &m) <= 0) { k += h + 1; h -= (rsize & 1) ^ 1; rlmatch = m; }
                else rrmatch = m;
            }
            break;
        }
        size = half; half >>= 1;
    }
    *idx = (k - j) > 0 ? j : i;
    return k - j;
}

// ---------------------------------------------------------------------------------------------
// superslice::Ext::equal_range_by (third-party crate, version constraint "1.0", source not in the
// reference tree; call site searcher.rs:164). Published algorithm: lower and upper bound searched in
// lock step with the same halving sequence; `size -= half` (not size = half), one final probe each.
// With a monotone comparator this is the ordinary equal range; the literal sequence matters only
// for quirk Q6 (comparator forced to Less for the last k-1 suffixes, searcher.rs:165-166).
// ---------------------------------------------------------------------------------------------
template <class F>
std::pair<usize, usize> equal_range_by(usize len, F f) {  // f(index) -> -1 / 0 / +1
    usize size = len;
    if (size == 0) return {0, 0};
    usize base0 = 0, base1 = 0;
    while (size > 1) {
        usize half = size / 2;
        usize mid0 = base0 + half, mid1 = base1 + half;
        int c0 = f(mid0), c1 = f(mid1);
        if (c0 < 0) base0 = mid0;
        if (c1 <= 0) base1 = mid1;
        size -= half;
    }
    int c0 = f(base0), c1 = f(base1);
    return {base0 + (c0 < 0 ? 1 : 0), base1 + (c1 <= 0 ? 1 : 0)};
}

// ---------------------------------------------------------------------------------------------
// searcher.rs:10-13, :95-97 (indexize), :99-143 (new), :145-180 (search)
// ---------------------------------------------------------------------------------------------
struct Searcher {
    std::unordered_map<uint64_t, std::pair<usize, usize>> cache;
    usize offset;

    static uint64_t indexize(const uint8_t* p) {  // searcher.rs:95-97: transmute [u8;8] -> u64 (little endian host)
        uint64_t v = 0;
        for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
        return v;
    }

    // searcher.rs:99-143: all 5^8 8-mers over ALPHABET, each through sa_searchb64 on the whole SA
    // `literal`: sa_search step for step (needed when sa is not sorted for dna: --trim)
    Searcher(const uint8_t* dna, usize dna_len, const int64_t* sa, usize sa_len, usize off, bool literal = false) : offset(off) {
        cache.reserve(400000);
        uint8_t p[8];
        int d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (;;) {
            for (int j = 0; j < 8; ++j) p[j] = ALPHABET[d[j]];
            int64_t out = 0;
            int64_t count = literal ? sa_search_literal(dna, int64_t(dna_len), p, 8, sa, int64_t(sa_len), &out)
                                    : sa_searchb_restated(dna, int64_t(dna_len), p, 8, sa, 0, int64_t(sa_len), &out);
            cache[indexize(p)] = {usize(out), usize(out) + usize(count)};
            int j = 7;  // odometer: last letter varies fastest, like the nested loops :108-115
            while (j >= 0 && ++d[j] == 5) { d[j] = 0; --j; }
            if (j < 0) break;
        }
    }

    // searcher.rs:145-180
    std::vector<Segment> search(const uint8_t* dna, usize dna_len, const int64_t* sa, const uint8_t* pattern,
                                usize plen) const {
        uint64_t index = indexize(pattern);
        auto it = cache.find(index);
        if (it == cache.end()) {  // searcher.rs:155-161 panics; cannot happen for a normalised strand
            fprintf(stderr, "oracle: Unable to find %llu\n", (unsigned long long)index);
            abort();
        }
        usize lstart = it->second.first, rstart = it->second.second;
        const int64_t* sub = sa + lstart;
        auto range = equal_range_by(rstart - lstart, [&](usize ix) -> int {
            usize x = usize(sub[ix]);
            if (x + plen > dna_len) return -1;  // searcher.rs:165-166 (Q6)
            int c = memcmp(dna + x, pattern, plen);  // a.cmp(b) on equal-length byte slices, :147-151
            return c < 0 ? -1 : (c > 0 ? 1 : 0);
        });
        std::vector<Segment> out;
        out.reserve(range.second > range.first ? range.second - range.first : 0);
        for (usize ix = lstart + range.first; ix < lstart + range.second; ++ix) {
            usize s = offset + usize(sa[ix]);
            out.push_back(Segment{0, s, s + plen});
        }
        return out;
    }
};

// ---------------------------------------------------------------------------------------------
// automaton.rs:206-216 d_ss
// ---------------------------------------------------------------------------------------------
int64_t d_ss(const Segment& a, const Segment& m) {
    if ((m.start >= a.start && m.start <= a.end) || (m.end >= a.start && m.end <= a.end)) return 0;
    int64_t d1 = std::llabs(int64_t(a.start) - int64_t(m.end));
    int64_t d2 = std::llabs(int64_t(a.end) - int64_t(m.start));
    return std::min(d1, d2);
}

struct Op {
    bool extend;
    usize i, l_end, r_end;      // ExtendArm { i: arm index, l_end, r_end }
    usize m_start, m_end;       // NewArm { i: probe position, m_start, m_end }
};

// automaton.rs:66-85
Op try_extend_arms(const std::vector<Arm>& arms, const Segment& m, int64_t e, usize i, usize ps) {
    for (usize j = 0; j < arms.size(); ++j) {
        const Arm& a = arms[j];
        int64_t thr = std::max<int64_t>(e, int64_t(0.1 * double(a.left.len())));  // :69, f64 then `as i64`
        if (a.active && d_ss(a.right, m) < thr && m.end > a.right.end) {
            Op op{}; op.extend = true; op.i = j; op.l_end = i + ps; op.r_end = m.end;
            return op;
        }
    }
    Op op{}; op.extend = false; op.i = i; op.m_start = m.start; op.m_end = m.end;
    return op;
}

struct Counters {
    std::atomic<uint64_t> probes{0}, searched{0}, skipped_n{0}, skipped_card{0}, matches{0};
    std::atomic<uint64_t> alg_bytes{0};  // sum 24*(ceil(log2(B+1))+1) + 8*(hi-lo) over searched probes
};

// automaton.rs:57-204
std::vector<Family> automaton_search_duplications(const uint8_t* needle, usize needle_len, usize needle_offset,
                                                  const uint8_t* strand, usize strand_len, const int64_t* sa,
                                                  const Searcher& searcher, const RunSettings& settings,
                                                  Counters* ctr) {
    std::vector<Arm> arms;
    usize i = 0;
    std::vector<Family> r;
    const usize step_size = settings.probe_size / 2;  // :90

    if (needle_len < settings.min_duplication_length) return r;  // :92-94
    // :96 computes needle.len() - probe_size - step_size in usize; a shorter needle would underflow
    // (panic in a debug build); we treat it as "no probes".
    if (needle_len < settings.probe_size + step_size) return r;
    const usize limit = needle_len - settings.probe_size - step_size;

    uint64_t n_probes = 0, n_searched = 0, n_skip_n = 0, n_skip_card = 0, n_matches = 0, alg = 0;
    while (i < limit) {  // :96
        i += step_size;  // :97
        ++n_probes;
        if (needle[i] == 'N') { ++n_skip_n; continue; }  // :100-102
        std::vector<Segment> found = searcher.search(strand, strand_len, sa, needle + i, settings.probe_size);
        if (ctr) {
            auto it = searcher.cache.find(Searcher::indexize(needle + i));
            uint64_t B = it->second.second - it->second.first;
            uint64_t lg = 0; while ((uint64_t(1) << lg) < B + 1) ++lg;  // ceil(log2(B+1))
            alg += 24 * (lg + 1) + 8 * uint64_t(found.size());
        }
        std::vector<Segment> matches;
        matches.reserve(found.size());
        for (const Segment& m : found) {
            if (m.start == i) continue;  // :106 (needle-local i against a global start: Q1)
            bool keep = !settings.reverse ? (m.start > i + needle_offset)                  // :109
                                          : (m.start >= needle_offset + needle_len - i);  // :111
            if (keep) matches.push_back(m);
        }
        ++n_searched;
        if (matches.size() > settings.max_cardinality) { ++n_skip_card; continue; }  // :115-117
        n_matches += matches.size();

        for (Arm& a : arms) a.dirty = false;  // :120

        std::vector<Op> todo;  // :122-134 (the rayon map preserves order)
        todo.reserve(matches.size());
        for (const Segment& m : matches)
            todo.push_back(try_extend_arms(arms, m, int64_t(settings.max_gap_size), i, settings.probe_size));

        for (const Op& op : todo)  // :136-143
            if (op.extend) {
                arms[op.i].left.end = op.l_end;
                arms[op.i].right.end = op.r_end;
                arms[op.i].dirty = true;
                arms[op.i].gap = 0;
            }
        for (const Op& op : todo)  // :145-163
            if (!op.extend)
                arms.push_back(Arm{Segment{0, op.i, op.i + settings.probe_size}, Segment{0, op.m_start, op.m_end},
                                   true, false, 0});

        for (Arm& a : arms)  // :166-171
            if (!a.dirty) {
                a.gap += step_size;
                if (uint32_t(a.gap) >= settings.max_gap_size) a.active = false;
            }

        if (arms.size() > 200) {  // :173-179
            std::vector<Arm> kept;
            for (const Arm& a : arms)
                if (a.active || a.left.len() >= settings.min_duplication_length ||
                    a.right.len() >= settings.min_duplication_length)
                    kept.push_back(a);
            arms.swap(kept);
        }

        bool all_inactive = true;  // :182
        for (const Arm& a : arms) if (a.active) { all_inactive = false; break; }
        if (!arms.empty() && all_inactive) {
            Family family;
            for (const Arm& a : arms)
                if (a.right.len() >= settings.min_duplication_length)  // :185 (right arm only)
                    family.push_back(ProtoSD{a.left.start, a.right.start, a.left.len(), a.right.len(), 0.f, false,
                                             false});
            if (!family.empty()) r.push_back(std::move(family));
            arms.clear();
        }
    }
    if (ctr) {
        ctr->probes += n_probes; ctr->searched += n_searched; ctr->skipped_n += n_skip_n;
        ctr->skipped_card += n_skip_card; ctr->matches += n_matches; ctr->alg_bytes += alg;
    }
    return r;  // :203 — arms still alive here are dropped (Q3)
}

// ---------------------------------------------------------------------------------------------
// bin/asgart.rs:137-258 SearchDuplications::run, given the suffix array (built by the caller:
// :149 r_divsufsort) — LUT (:151-155), chunk fan-out (:201-240), fix-ups (:229-237), flatten (:241-253)
// ---------------------------------------------------------------------------------------------
std::vector<Family> search_duplications_step(const uint8_t* strand, usize strand_len, const int64_t* sa,
                                             const std::vector<std::pair<usize, usize>>& chunks,
                                             const RunSettings& settings, int threads, double* t_lut,
                                             double* t_search, Counters* ctr, usize sa_len = 0) {
    auto t0 = std::chrono::steady_clock::now();
    // sa_len != 0: --trim, sa is the (shifted) suffix array of a slice of the strand (:142-147)
    Searcher searcher(strand, strand_len, sa, sa_len ? sa_len : strand_len, 0, sa_len != 0);  // :151-155
    auto t1 = std::chrono::steady_clock::now();

    std::vector<std::vector<Family>> results(chunks.size());
    std::atomic<usize> next{0};
    auto worker = [&]() {
        for (;;) {
            usize id = next.fetch_add(1);
            if (id >= chunks.size()) break;
            const auto& chunk = chunks[id];
            std::vector<uint8_t> needle_buf;
            const uint8_t* needle;
            if (!settings.reverse && !settings.complement) {  // :207-208
                needle = strand + chunk.first;
            } else {
                needle_buf.assign(strand + chunk.first, strand + chunk.first + chunk.second);  // :210
                if (settings.complement)                                                       // :211-213
                    for (auto& c : needle_buf) c = complement_nucleotide(c);
                if (settings.reverse) std::reverse(needle_buf.begin(), needle_buf.end());  // :214-216
                needle = needle_buf.data();
            }
            std::vector<Family> fams = automaton_search_duplications(needle, chunk.second, chunk.first, strand,
                                                                     strand_len, sa, searcher, settings, ctr);
            for (Family& f : fams)  // :229-237
                for (ProtoSD& sd : f) {
                    if (!settings.reverse) sd.left += chunk.first;
                    else sd.left = chunk.first + chunk.second - sd.left - sd.left_length;
                }
            results[id] = std::move(fams);
        }
    };
    int nt = std::max(1, std::min<int>(threads, int(chunks.size())));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();

    std::vector<Family> result;  // :241-253
    for (auto& per_chunk : results)
        for (Family& f : per_chunk) {
            for (ProtoSD& sd : f) { sd.reversed = settings.reverse; sd.complemented = settings.complement; }
            result.push_back(std::move(f));
        }
    auto t2 = std::chrono::steady_clock::now();
    if (t_lut) *t_lut = std::chrono::duration<double>(t1 - t0).count();
    if (t_search) *t_search = std::chrono::duration<double>(t2 - t1).count();
    return result;
}

// ---------------------------------------------------------------------------------------------
// Post-steps. structs.rs:454-467 n_content; bin/asgart.rs:81-96 FilterNs; :33-51 ReOrder;
// :481-562 subsegment/overlap/merge/reduce_overlap; :53-65 Sort
// ---------------------------------------------------------------------------------------------
struct RefPanic : std::runtime_error { using std::runtime_error::runtime_error; };

// `n1` = strand.len() ('$' included): an inclusive range that passes it is a slice panic in the reference
float n_content(const ProtoSD& sd, const uint8_t* strand, usize n1) {  // structs.rs:454-467, inclusive ranges, f32
    if (sd.left >= n1 || sd.left_length >= n1 - sd.left || sd.right >= n1 || sd.right_length >= n1 - sd.right)
        throw RefPanic("FilterNs: slice index out of range (structs.rs:455-466)");
    usize cl = 0, cr = 0;
    for (usize p = sd.left; p <= sd.left + sd.left_length; ++p) cl += (strand[p] == 'n' || strand[p] == 'N');
    for (usize p = sd.right; p <= sd.right + sd.right_length; ++p) cr += (strand[p] == 'n' || strand[p] == 'N');
    float l = float(cl) / float(sd.left_length);
    float r = float(cr) / float(sd.right_length);
    return fmaxf(l, r);
}

void step_filter_ns(std::vector<Family>& fams, const uint8_t* strand, usize n1) {  // bin/asgart.rs:81-96
    std::vector<Family> out;
    for (Family& f : fams) {
        Family kept;
        for (const ProtoSD& sd : f)
            if (n_content(sd, strand, n1) <= 0.2f) kept.push_back(sd);
        if (!kept.empty()) out.push_back(std::move(kept));
    }
    fams.swap(out);
}

void step_reorder(std::vector<Family>& fams) {  // bin/asgart.rs:33-51 — positions only (Q4)
    for (Family& f : fams)
        for (ProtoSD& sd : f)
            if (sd.left > sd.right) std::swap(sd.left, sd.right);
}

bool subsegment(usize xs, usize xl, usize ys, usize yl) {  // :482-487
    return xs >= ys && xs + xl <= ys + yl;
}
bool overlap(usize xs, usize xl, usize ys, usize yl) {  // :489-495
    usize xe = xs + xl, ye = ys + yl;
    return (xs >= ys && xs <= ye && xe >= ye) || (ys >= xs && ys <= xe && ye >= xe);
}
ProtoSD merge(const ProtoSD& x, const ProtoSD& y) {  // :497-513 — including the mixed-up lengths (Q5)
    usize new_left = std::min(x.left, y.left);
    usize lsize = std::max(x.left + x.left_length, y.left + y.right_length) - new_left;
    usize new_right = std::min(x.right, y.right);
    usize rsize = std::max(x.right + x.left_length, y.right + y.right_length) - new_right;
    return ProtoSD{new_left, new_right, lsize, rsize, 0.f, x.reversed, x.complemented};
}
Family reduce_once(const Family& result) {  // :516-551
    Family news;
    for (const ProtoSD& x : result) {
        bool absorbed = false;
        for (ProtoSD& y : news) {
            if (subsegment(x.left, x.left_length, y.left, y.left_length) &&
                subsegment(x.right, x.right_length, y.right, y.right_length)) { absorbed = true; break; }
            if (subsegment(y.left, y.left_length, x.left, x.left_length) &&
                subsegment(y.right, y.right_length, x.right, x.right_length)) {
                y.left = x.left; y.right = x.right; y.left_length = x.left_length; y.right_length = x.right_length;
                absorbed = true; break;
            }
            if (overlap(x.left, x.left_length, y.left, y.left_length) &&
                overlap(x.right, x.right_length, y.right, y.right_length)) {
                ProtoSD z = merge(x, y);
                y.left = z.left; y.right = z.right; y.left_length = z.left_length; y.right_length = z.right_length;
                absorbed = true; break;
            }
        }
        if (!absorbed) news.push_back(x);
    }
    return news;
}
Family reduce_overlap(const Family& result) {  // :515-562
    usize old_size = result.size();
    Family news = reduce_once(result);
    usize new_size = news.size();
    while (new_size < old_size) {
        old_size = news.size();
        news = reduce_once(news);
        new_size = news.size();
    }
    return news;
}
void step_reduce_overlap(std::vector<Family>& fams) {  // :67-79
    for (Family& f : fams) f = reduce_overlap(f);
}
// structs.rs:28-34 — complement() of structs.rs: panics on anything outside the TR table (so on '$')
uint8_t complement_strict(uint8_t n) {
    switch (n) {
        case 'A': return 'T'; case 'T': return 'A'; case 'G': return 'C'; case 'C': return 'G'; case 'N': return 'N';
        case 'a': return 't'; case 't': return 'a'; case 'g': return 'c'; case 'c': return 'g'; case 'n': return 'n';
        default: throw RefPanic("Unknown nucleotide");
    }
}
// bio::alignment::distance::levenshtein (crate bio "*", Cargo.toml:13; source not on disk): the unit-cost global edit
// distance. Restated as the textbook two-row dynamic programme.
uint32_t levenshtein_dp(const std::vector<uint8_t>& a, const std::vector<uint8_t>& b) {
    std::vector<uint32_t> prev(b.size() + 1), cur(b.size() + 1);
    for (size_t j = 0; j <= b.size(); ++j) prev[j] = uint32_t(j);
    for (size_t i = 1; i <= a.size(); ++i) {
        cur[0] = uint32_t(i);
        for (size_t j = 1; j <= b.size(); ++j) {
            uint32_t sub = prev[j - 1] + (a[i - 1] != b[j - 1] ? 1u : 0u);
            cur[j] = std::min(sub, std::min(prev[j], cur[j - 1]) + 1u);
        }
        prev.swap(cur);
    }
    return prev[b.size()];
}
// structs.rs:439-452 ProtoSD::levenshtein: inclusive ranges, reverse THEN complement of the right arm, f64 arithmetic
double sd_levenshtein(const ProtoSD& sd, const uint8_t* strand, usize strand_len) {
    if (sd.left + sd.left_length >= strand_len || sd.right + sd.right_length >= strand_len)
        throw RefPanic("range end index out of range for slice");
    std::vector<uint8_t> left_arm(strand + sd.left, strand + sd.left + sd.left_length + 1);
    std::vector<uint8_t> right_arm(strand + sd.right, strand + sd.right + sd.right_length + 1);
    if (sd.reversed) std::reverse(right_arm.begin(), right_arm.end());
    if (sd.complemented)
        for (auto& c : right_arm) c = complement_strict(c);
    double dist = double(levenshtein_dp(left_arm, right_arm));
    return 100.0 * (1.0 - dist / double(std::max(sd.left_length, sd.right_length)));
}
void step_compute_score(std::vector<Family>& fams, const uint8_t* strand, usize strand_len) {  // bin/asgart.rs:98-111
    for (Family& f : fams)
        for (ProtoSD& sd : f) sd.identity = float(sd_levenshtein(sd, strand, strand_len));
}
void step_sort(std::vector<Family>& fams) {  // :53-65 (sort_by is stable)
    for (Family& f : fams)
        std::stable_sort(f.begin(), f.end(), [](const ProtoSD& a, const ProtoSD& b) { return a.left < b.left; });
}

// ---------------------------------------------------------------------------------------------
// bin/asgart.rs:273-471 prepare_data: read_fasta (:278-313), find_chunks_to_process (:317-366),
// concatenation + '$' (:375-430). --trim: validation in oracle_effective_trim, search in oracle_search_trim.
// ---------------------------------------------------------------------------------------------
struct Prepared {
    std::string file_names;
    std::vector<uint8_t> data;  // strand incl. trailing '$'
    std::vector<Start> map;
    std::vector<std::pair<usize, usize>> chunks;
    std::string error;
};

std::string rtrim(const std::string& s) {
    usize e = s.size();
    while (e > 0 && (s[e - 1] == '\n' || s[e - 1] == '\r' || s[e - 1] == ' ' || s[e - 1] == '\t' ||
                     s[e - 1] == '\v' || s[e - 1] == '\f'))
        --e;
    return s.substr(0, e);
}

// :278-313 with the record semantics of the `bio` crate's FASTA reader (io/fasta.rs; Cargo.toml asks for "*" and the
// reference ships no lock file, so this follows the published Reader::read / Records::next):
//  * a record starts at a line whose first byte is '>'; the first line of a non-empty file must be one ("Expected > at
//    record start." -> "Unable to parse", also for a leading blank line);
//  * id = header up to the first white space, the rest is the description; sequence lines are joined after trim_end;
//  * the iterator ENDS at the first empty record (no id, no description, no sequence: Record::is_empty) and the rest of
//    the file is never read.
bool read_fasta(const std::string& filename, bool skip_masked, std::vector<Start>& map, std::vector<uint8_t>& r,
                std::string& err) {
    std::ifstream in(filename, std::ios::binary);
    if (!in) { err = "Unable to read FASTA file `" + filename + "`"; return false; }
    std::string line;
    usize counter = 0;
    bool have = false, no_id_no_desc = false;
    std::string name;
    std::vector<uint8_t> seq;
    auto flush = [&]() {
        if (no_id_no_desc && seq.empty()) return false;   // Records::next: `Ok(()) if record.is_empty() => None`
        if (!skip_masked)  // :291-293
            for (auto& c : seq) if (c >= 'a' && c <= 'z') c = uint8_t(c - 'a' + 'A');
        for (auto& c : seq) {  // :294-301
            if (in_set(ALPHABET_MASKED, c) && skip_masked) c = 'N';
            else if (!in_set(ALPHABET, c)) c = 'N';
        }
        map.push_back(Start{name, counter, seq.size()});  // :303-307
        counter += seq.size();
        r.insert(r.end(), seq.begin(), seq.end());
        seq.clear();
        return true;
    };
    while (std::getline(in, line)) {
        if (!line.empty() && line[0] == '>') {
            if (have && !flush()) return true;
            std::string h = rtrim(line.substr(1));
            usize sp = 0;
            while (sp < h.size() && !isspace((unsigned char)h[sp])) ++sp;
            name = h.substr(0, sp);  // record.id(): header up to the first whitespace
            no_id_no_desc = h.empty();
            have = true;
        } else {
            if (!have) { err = "Unable to parse `" + filename + "`"; return false; }
            std::string t = rtrim(line);
            seq.insert(seq.end(), t.begin(), t.end());
        }
    }
    if (have) flush();
    return true;
}

// :317-366, coordinates relative to the fragment
std::vector<std::pair<usize, usize>> find_chunks_to_process(const uint8_t* strand, usize len) {
    auto count_n = [&](usize start) {
        usize c = 0;
        while (start + c < len && (strand[start + c] == 'n' || strand[start + c] == 'N')) ++c;
        return c;
    };
    const usize threshold = 5000;
    usize start = 0, count = 0, i = 0;
    std::vector<std::pair<usize, usize>> chunks;
    while (i < len) {
        uint8_t n = strand[i];
        if (n == 'n' || n == 'N') {
            usize n_count = count_n(i);
            if (n_count > threshold) {
                if (count > 0) { chunks.push_back({start, count}); count = 0; }
                start = i + n_count;
            } else {
                count += n_count;
            }
            i += n_count;
        } else {
            if (count == 0) { count = 1; start = i; } else { count += 1; }
            i += 1;
        }
    }
    if (count != 0) chunks.push_back({start, count});
    if (chunks.empty()) chunks.push_back({0, len});
    return chunks;
}

Prepared* prepare_data(const std::vector<std::string>& files, bool skip_masked) {
    Prepared* p = new Prepared();
    usize offset = 0;
    for (const std::string& f : files) {
        std::vector<Start> map;
        std::vector<uint8_t> new_strand;
        if (!read_fasta(f, skip_masked, map, new_strand, p->error)) return p;
        for (const Start& chr : map)  // :381-387
            for (auto c : find_chunks_to_process(new_strand.data() + chr.position, chr.length))
                p->chunks.push_back({chr.position + offset + c.first, c.second});
        for (Start& s : map) { s.position += offset; p->map.push_back(s); }  // :388-391
        offset += new_strand.size();
        p->data.insert(p->data.end(), new_strand.begin(), new_strand.end());
    }
    p->data.push_back('$');  // :430
    for (usize i = 0; i < files.size(); ++i) p->file_names += (i ? ", " : "") + files[i];  // :466
    return p;
}

// ---------------------------------------------------------------------------------------------
// bin/asgart.rs:770-821 (ProtoSD -> SD, chr lookup structs.rs:86-90) + exporters.rs:12-25
// (serde_json::to_string_pretty + '\n'; field order = declaration order, structs.rs:36-58,60-72,93-98,471-493)
// ---------------------------------------------------------------------------------------------
void json_escape(std::string& o, const std::string& s) {
    o += '"';
    for (unsigned char c : s) {
        switch (c) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\n': o += "\\n"; break;
            case '\r': o += "\\r"; break;
            case '\t': o += "\\t"; break;
            case '\b': o += "\\b"; break;
            case '\f': o += "\\f"; break;
            default:
                if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
                else o += char(c);
        }
    }
    o += '"';
}

std::string f32_json(float v) {  // serde_json prints f32 via ryu: shortest round-trip, always with ".0" for integers
    if (v == 0.f) return std::signbit(v) ? "-0.0" : "0.0";
    char b[64];
    for (int prec = 1; prec <= 9; ++prec) {
        snprintf(b, sizeof b, "%.*g", prec, double(v));
        if (strtof(b, nullptr) == v) break;
    }
    std::string s(b);
    if (s.find('.') == std::string::npos && s.find('e') == std::string::npos && s.find("inf") == std::string::npos &&
        s.find("nan") == std::string::npos)
        s += ".0";
    return s;
}

const Start* find_chr_by_pos(const std::vector<Start>& map, usize pos) {  // structs.rs:86-90
    for (const Start& c : map)
        if (pos >= c.position && pos < c.position + c.length) return &c;
    return nullptr;
}

std::string to_json(const Prepared& prep, const RunSettings& st, const std::vector<Family>& fams) {
    std::string o;
    auto num = [&](usize v) { o += std::to_string(v); };
    usize total = 0;
    for (const Start& s : prep.map) total += s.length;  // :772
    o += "{\n  \"strand\": {\n    \"name\": ";
    json_escape(o, prep.file_names);
    o += ",\n    \"length\": "; num(total);
    o += ",\n    \"map\": [";
    for (usize i = 0; i < prep.map.size(); ++i) {
        o += i ? ",\n" : "\n";
        o += "      {\n        \"name\": "; json_escape(o, prep.map[i].name);
        o += ",\n        \"position\": "; num(prep.map[i].position);
        o += ",\n        \"length\": "; num(prep.map[i].length);
        o += "\n      }";
    }
    o += prep.map.empty() ? "]" : "\n    ]";
    o += "\n  },\n  \"settings\": {\n    \"probe_size\": "; num(st.probe_size);
    o += ",\n    \"max_gap_size\": "; num(st.max_gap_size);
    o += ",\n    \"min_duplication_length\": "; num(st.min_duplication_length);
    o += ",\n    \"max_cardinality\": "; num(st.max_cardinality);
    o += ",\n    \"trim\": ";
    if (st.has_trim) { o += "[\n      "; num(st.trim_a); o += ",\n      "; num(st.trim_b); o += "\n    ]"; }
    else o += "null";
    o += ",\n    \"skip_masked\": "; o += st.skip_masked ? "true" : "false";
    o += "\n  },\n  \"families\": [";
    for (usize f = 0; f < fams.size(); ++f) {
        o += f ? ",\n" : "\n";
        o += "    [";
        for (usize j = 0; j < fams[f].size(); ++j) {
            const ProtoSD& sd = fams[f][j];
            const Start* cl = find_chr_by_pos(prep.map, sd.left);
            const Start* cr = find_chr_by_pos(prep.map, sd.right);
            o += j ? ",\n" : "\n";
            o += "      {\n        \"chr_left\": "; json_escape(o, cl ? cl->name : "unknown");
            o += ",\n        \"chr_right\": "; json_escape(o, cr ? cr->name : "unknown");
            o += ",\n        \"global_left_position\": "; num(sd.left);
            o += ",\n        \"global_right_position\": "; num(sd.right);
            o += ",\n        \"chr_left_position\": "; num(sd.left - (cl ? cl->position : 0));
            o += ",\n        \"chr_right_position\": "; num(sd.right - (cr ? cr->position : 0));
            o += ",\n        \"left_length\": "; num(sd.left_length);
            o += ",\n        \"right_length\": "; num(sd.right_length);
            o += ",\n        \"left_seq\": null,\n        \"right_seq\": null";
            o += ",\n        \"identity\": "; o += f32_json(sd.identity);
            o += ",\n        \"reversed\": "; o += sd.reversed ? "true" : "false";
            o += ",\n        \"complemented\": "; o += sd.complemented ? "true" : "false";
            o += "\n      }";
        }
        o += fams[f].empty() ? "]" : "\n    ]";
    }
    o += fams.empty() ? "]" : "\n  ]";
    o += "\n}\n";  // exporters.rs:15-17 writeln!
    return o;
}

// ---------------------------------------------------------------------------------------------
// Suffix array — the reference uses libdivsufsort (divsufsort.c:331-358). The suffix array of a text
// is unique, so the oracle restates it as a plain prefix-doubling sort (Manber–Myers with std::sort);
// tests pin it to the real divsufsort64 from oracle/_ref. O(n log^2 n): for test-sized inputs only.
// A suffix that is a proper prefix of another sorts first (what divsufsort yields on texts without
// a terminator).
// ---------------------------------------------------------------------------------------------
int suffix_array_doubling(const uint8_t* T, int64_t* SA, int64_t n) {
    if (!T || !SA || n < 0) return -1;
    if (n == 0) return 0;
    std::vector<int64_t> rank(n), tmp(n);
    for (int64_t i = 0; i < n; ++i) { SA[i] = i; rank[i] = T[i]; }
    for (int64_t h = 1;; h <<= 1) {
        auto key2 = [&](int64_t i) { return i + h < n ? rank[i + h] : int64_t(-1); };
        auto less = [&](int64_t a, int64_t b) {
            if (rank[a] != rank[b]) return rank[a] < rank[b];
            return key2(a) < key2(b);
        };
        std::sort(SA, SA + n, less);
        tmp[SA[0]] = 0;
        for (int64_t i = 1; i < n; ++i) tmp[SA[i]] = tmp[SA[i - 1]] + (less(SA[i - 1], SA[i]) ? 1 : 0);
        rank = tmp;
        if (rank[SA[n - 1]] == n - 1) break;
    }
    return 0;
}

struct Result {
    std::vector<Family> fams;
};

std::vector<Family> from_arrays(const int64_t* fam_offsets, int64_t n_fam, const uint64_t* fields,
                                const float* identity, const uint8_t* flags) {
    std::vector<Family> fams(n_fam);
    for (int64_t f = 0; f < n_fam; ++f)
        for (int64_t j = fam_offsets[f]; j < fam_offsets[f + 1]; ++j)
            fams[f].push_back(ProtoSD{usize(fields[4 * j]), usize(fields[4 * j + 1]), usize(fields[4 * j + 2]),
                                      usize(fields[4 * j + 3]), identity ? identity[j] : 0.f,
                                      flags ? flags[2 * j] != 0 : false, flags ? flags[2 * j + 1] != 0 : false});
    return fams;
}

}  // namespace

// =============================================================================================
// C ABI (ctypes) — test infrastructure
// =============================================================================================
extern "C" {

struct oracle_settings {
    uint64_t probe_size;
    uint32_t max_gap_size;
    uint32_t reverse, complement, skip_masked;
    uint64_t min_duplication_length;
    uint64_t max_cardinality;
    uint32_t has_trim;
    uint32_t compute_score;   // --compute-score (structs.rs:57); occupies former padding
    uint64_t trim_a, trim_b;
};

static RunSettings to_settings(const oracle_settings* s) {
    RunSettings r{};
    r.probe_size = s->probe_size; r.max_gap_size = s->max_gap_size;
    r.min_duplication_length = s->min_duplication_length; r.max_cardinality = s->max_cardinality;
    r.reverse = s->reverse != 0; r.complement = s->complement != 0; r.skip_masked = s->skip_masked != 0;
    r.has_trim = s->has_trim != 0; r.trim_a = s->trim_a; r.trim_b = s->trim_b;
    return r;
}

int oracle_suffix_array(const uint8_t* T, int64_t* SA, int64_t n) { return suffix_array_doubling(T, SA, n); }

// Searcher::new restated: entry e (odometer order over ALPHABET=[A,T,G,C,N], last letter fastest) gets
// keys[e] = LE u64 of the 8-mer, lo[e], hi[e]
int oracle_lut(const uint8_t* T, int64_t n1, const int64_t* SA, uint64_t* keys, int64_t* lo, int64_t* hi) {
    uint8_t p[8];
    int d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t e = 0;
    for (;;) {
        for (int j = 0; j < 8; ++j) p[j] = ALPHABET[d[j]];
        int64_t out = 0;
        int64_t count = sa_searchb_restated(T, n1, p, 8, SA, 0, n1, &out);
        keys[e] = Searcher::indexize(p); lo[e] = out; hi[e] = out + count; ++e;
        int j = 7;
        while (j >= 0 && ++d[j] == 5) { d[j] = 0; --j; }
        if (j < 0) break;
    }
    return 0;
}

// sa_search64-shaped entry for pinning the restated search against the reference C function
int64_t oracle_sa_searchb(const uint8_t* T, int64_t Tsize, const uint8_t* P, int64_t Psize, const int64_t* SA,
                          int64_t SAsize, int64_t* idx, int64_t l, int64_t r) {
    (void)SAsize;
    return sa_searchb_restated(T, Tsize, P, Psize, SA, l, r, idx);
}

// Searcher::search for one pattern (LUT rebuilt per handle): returns the SA-ordered match starts
struct oracle_searcher { Searcher* s; };
void* oracle_searcher_new(const uint8_t* T, int64_t n1, const int64_t* SA) {
    return new Searcher(T, usize(n1), SA, usize(n1), 0);
}
void oracle_searcher_free(void* h) { delete static_cast<Searcher*>(h); }
int64_t oracle_searcher_search(void* h, const uint8_t* T, int64_t n1, const int64_t* SA, const uint8_t* pattern,
                               int64_t plen, int64_t* out, int64_t cap) {
    auto v = static_cast<Searcher*>(h)->search(T, usize(n1), SA, pattern, usize(plen));
    for (usize i = 0; i < v.size() && int64_t(i) < cap; ++i) out[i] = int64_t(v[i].start);
    return int64_t(v.size());
}

// SearchDuplications::run (given SA) [+ post-steps selected by post_mask: 1 FilterNs, 2 ReOrder,
// 4 ReduceOverlap, 8 Sort — applied in the reference's order, bin/asgart.rs:738-747]
// phase_seconds[0..3) = LUT, search+automaton, post-steps; counters[0..6) = probes, searched, skipped_n,
// skipped_card, matches, algorithmic bytes
void* oracle_search(const uint8_t* T, int64_t n1, const int64_t* SA, const uint64_t* chunks, int64_t n_chunks,
                    const oracle_settings* s, int post_mask, int threads, double* phase_seconds,
                    uint64_t* counters) {
    RunSettings st = to_settings(s);
    std::vector<std::pair<usize, usize>> ch;
    for (int64_t i = 0; i < n_chunks; ++i) ch.push_back({usize(chunks[2 * i]), usize(chunks[2 * i + 1])});
    Counters ctr;
    double tl = 0, ts = 0;
    Result* r = new Result();
    r->fams = search_duplications_step(T, usize(n1), SA, ch, st, threads, &tl, &ts, counters ? &ctr : nullptr);
    auto t0 = std::chrono::steady_clock::now();
    try { if (post_mask & 1) step_filter_ns(r->fams, T, usize(n1)); } catch (const RefPanic&) { delete r; return nullptr; }
    if (post_mask & 2) step_reorder(r->fams);
    if (post_mask & 4) step_reduce_overlap(r->fams);
    if (post_mask & 16) {   // ComputeScore sits between ReduceOverlap and Sort (bin/asgart.rs:744-747)
        try { step_compute_score(r->fams, T, usize(n1)); } catch (const RefPanic&) { delete r; return nullptr; }
    }
    if (post_mask & 8) step_sort(r->fams);
    auto t1 = std::chrono::steady_clock::now();
    if (phase_seconds) {
        phase_seconds[0] = tl; phase_seconds[1] = ts;
        phase_seconds[2] = std::chrono::duration<double>(t1 - t0).count();
    }
    if (counters) {
        counters[0] = ctr.probes; counters[1] = ctr.searched; counters[2] = ctr.skipped_n;
        counters[3] = ctr.skipped_card; counters[4] = ctr.matches; counters[5] = ctr.alg_bytes;
    }
    return r;
}

// the same with --trim: SA = suffix array of strand[a..b]+'$' shifted by a (sa_len = b - a + 1 entries), built by the caller
void* oracle_search_trim(const uint8_t* T, int64_t n1, const int64_t* SA, int64_t sa_len, const uint64_t* chunks,
                         int64_t n_chunks, const oracle_settings* s, int post_mask, int threads) {
    RunSettings st = to_settings(s);
    std::vector<std::pair<usize, usize>> ch;
    for (int64_t i = 0; i < n_chunks; ++i) ch.push_back({usize(chunks[2 * i]), usize(chunks[2 * i + 1])});
    Result* r = new Result();
    r->fams = search_duplications_step(T, usize(n1), SA, ch, st, threads, nullptr, nullptr, nullptr, usize(sa_len));
    try { if (post_mask & 1) step_filter_ns(r->fams, T, usize(n1)); } catch (const RefPanic&) { delete r; return nullptr; }
    if (post_mask & 2) step_reorder(r->fams);
    if (post_mask & 4) step_reduce_overlap(r->fams);
    if (post_mask & 16) {
        try { step_compute_score(r->fams, T, usize(n1)); } catch (const RefPanic&) { delete r; return nullptr; }
    }
    if (post_mask & 8) step_sort(r->fams);
    return r;
}

int64_t oracle_sa_search_literal(const uint8_t* T, int64_t Tsize, const uint8_t* P, int64_t Psize, const int64_t* SA,
                                 int64_t SAsize, int64_t* idx) {
    return sa_search_literal(T, Tsize, P, Psize, SA, SAsize, idx);
}

// prepare_data's validation of --trim (bin/asgart.rs:432-463; strand_len includes the '$'): returns 0 when trimming is
// skipped, else 1 with the effective (start, stop)
int oracle_effective_trim(uint64_t shift, uint64_t stop, uint64_t strand_len, uint64_t* out) {
    if (stop >= strand_len) stop = strand_len - 1;
    if (stop <= shift) return 0;
    if (shift >= strand_len) return 0;
    out[0] = shift; out[1] = stop;
    return 1;
}

void* oracle_result_from_arrays(const int64_t* fam_offsets, int64_t n_fam, const uint64_t* fields,
                                const float* identity, const uint8_t* flags) {
    Result* r = new Result();
    r->fams = from_arrays(fam_offsets, n_fam, fields, identity, flags);
    return r;
}
// returns -1 where the reference panics (ComputeScore on an arm that ends on '$' under -C, or past the strand)
int oracle_result_post(void* h, const uint8_t* T, int64_t n1, int post_mask) {
    Result* r = static_cast<Result*>(h);
    try { if (post_mask & 1) step_filter_ns(r->fams, T, usize(n1)); } catch (const RefPanic&) { return -1; }
    if (post_mask & 2) step_reorder(r->fams);
    if (post_mask & 4) step_reduce_overlap(r->fams);
    if (post_mask & 16) {
        try { step_compute_score(r->fams, T, usize(n1)); } catch (const RefPanic&) { return -1; }
    }
    if (post_mask & 8) step_sort(r->fams);
    return 0;
}
// plain edit distance of two byte strings (for pinning the GPU kernel on arbitrary pairs)
uint32_t oracle_levenshtein(const uint8_t* a, int64_t na, const uint8_t* b, int64_t nb) {
    return levenshtein_dp(std::vector<uint8_t>(a, a + na), std::vector<uint8_t>(b, b + nb));
}
int64_t oracle_result_n_families(void* h) { return int64_t(static_cast<Result*>(h)->fams.size()); }
int64_t oracle_result_n_sds(void* h) {
    int64_t n = 0;
    for (auto& f : static_cast<Result*>(h)->fams) n += int64_t(f.size());
    return n;
}
void oracle_result_copy(void* h, int64_t* fam_offsets, uint64_t* fields, float* identity, uint8_t* flags) {
    Result* r = static_cast<Result*>(h);
    int64_t j = 0;
    fam_offsets[0] = 0;
    for (usize f = 0; f < r->fams.size(); ++f) {
        for (const ProtoSD& sd : r->fams[f]) {
            fields[4 * j] = sd.left; fields[4 * j + 1] = sd.right;
            fields[4 * j + 2] = sd.left_length; fields[4 * j + 3] = sd.right_length;
            identity[j] = sd.identity; flags[2 * j] = sd.reversed; flags[2 * j + 1] = sd.complemented;
            ++j;
        }
        fam_offsets[f + 1] = j;
    }
}
void oracle_result_free(void* h) { delete static_cast<Result*>(h); }

// prepare_data: `files` = '\n'-separated paths
void* oracle_prepare(const char* files, int skip_masked) {
    std::vector<std::string> fl;
    std::stringstream ss(files);
    std::string f;
    while (std::getline(ss, f, '\n')) if (!f.empty()) fl.push_back(f);
    return prepare_data(fl, skip_masked != 0);
}
const char* oracle_prepared_error(void* h) {
    Prepared* p = static_cast<Prepared*>(h);
    return p->error.empty() ? nullptr : p->error.c_str();
}
int64_t oracle_prepared_strand_len(void* h) { return int64_t(static_cast<Prepared*>(h)->data.size()); }
const uint8_t* oracle_prepared_strand(void* h) { return static_cast<Prepared*>(h)->data.data(); }
int64_t oracle_prepared_n_chunks(void* h) { return int64_t(static_cast<Prepared*>(h)->chunks.size()); }
void oracle_prepared_chunks(void* h, uint64_t* out) {
    Prepared* p = static_cast<Prepared*>(h);
    for (usize i = 0; i < p->chunks.size(); ++i) { out[2 * i] = p->chunks[i].first; out[2 * i + 1] = p->chunks[i].second; }
}
int64_t oracle_prepared_n_fragments(void* h) { return int64_t(static_cast<Prepared*>(h)->map.size()); }
const char* oracle_prepared_fragment(void* h, int64_t i, uint64_t* position, uint64_t* length) {
    Prepared* p = static_cast<Prepared*>(h);
    *position = p->map[i].position; *length = p->map[i].length;
    return p->map[i].name.c_str();
}
// prepared data from memory (tests that skip the FASTA file): one or more fragments
void* oracle_prepare_from_memory(const char* file_names, const uint8_t* strand_no_dollar, int64_t n,
                                 const char* frag_names /* '\n'-separated */, const uint64_t* frag_pos,
                                 const uint64_t* frag_len, int64_t n_frag) {
    Prepared* p = new Prepared();
    p->file_names = file_names;
    p->data.assign(strand_no_dollar, strand_no_dollar + n);
    std::stringstream ss(frag_names);
    std::string nm;
    for (int64_t i = 0; i < n_frag; ++i) {
        std::getline(ss, nm, '\n');
        p->map.push_back(Start{nm, usize(frag_pos[i]), usize(frag_len[i])});
        for (auto c : find_chunks_to_process(p->data.data() + frag_pos[i], usize(frag_len[i])))
            p->chunks.push_back({usize(frag_pos[i]) + c.first, c.second});
    }
    p->data.push_back('$');
    return p;
}
void oracle_prepared_free(void* h) { delete static_cast<Prepared*>(h); }

// JSON (exporters.rs:12-25): returns a malloc'd NUL-terminated string; free with oracle_free_string
char* oracle_to_json(void* prepared, const oracle_settings* s, void* result) {
    std::string js = to_json(*static_cast<Prepared*>(prepared), to_settings(s), static_cast<Result*>(result)->fams);
    char* out = static_cast<char*>(malloc(js.size() + 1));
    memcpy(out, js.c_str(), js.size() + 1);
    return out;
}
void oracle_free_string(char* s) { free(s); }

}  // extern "C"
