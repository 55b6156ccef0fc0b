"""Device-logic headers (kmer_core.h / automaton_core.h) compiled for the host, against the oracle.
Covers the packed-window comparator, the LUT slots, the literal lock-step equal range, the match filters and the
event / segment / death-time reformulation of the automaton — everything that does not need CUDA to be checked."""
import numpy as np
import pytest

import oracle
from asgart_b200 import _lib as ablib
from tests import cases, emul_harness, kat


def _settings_pair(kw):
    so = oracle.make_settings(**kw)
    sc = ablib.Settings(so.probe_size, so.max_gap_size, so.reverse, so.complement, so.skip_masked,
                        so.min_duplication_length, so.max_cardinality, 0, 0, 0)
    return so, sc


def _strip(fams):
    return [[sd[:4] for sd in f] for f in fams]


@pytest.mark.parametrize("case", kat.cases(), ids=lambda c: c[0])
def test_emul_kat(case):
    name, text, chunks, kw, expected = case
    strand = np.concatenate([text, np.frombuffer(b"$", dtype=np.uint8)])
    sa = oracle.best_suffix_array(strand)
    so, sc = _settings_pair(kw)
    want = oracle.search(strand, sa, chunks, so, 0, threads=1)
    got, ctr = emul_harness.search(strand, sa, chunks, sc)
    assert got == _strip(want.families.as_lists())
    assert ctr == want.counters


@pytest.mark.parametrize("seed", [11, 12, 13])
@pytest.mark.parametrize("label,kw", cases.settings_grid(), ids=[s[0] for s in cases.settings_grid()])
def test_emul_stress(seed, label, kw):
    text = cases.stress_text(seed)
    prep = oracle.Prepared.from_memory(text, [("a", 0, 25000), ("b", 25000, len(text) - 25000)])
    strand = prep.strand
    sa = oracle.best_suffix_array(strand)
    so, sc = _settings_pair(kw)
    want = oracle.search(strand, sa, prep.chunks, so, 0, threads=2)
    got, ctr = emul_harness.search(strand, sa, prep.chunks, sc)
    assert got == _strip(want.families.as_lists())
    assert ctr == want.counters
    assert ctr["matches"] > 0


def test_emul_random_settings_sweep():
    """Random texts (size, duplication density, N-runs, tandem repeats, one or two fragments) under random settings: probe
    sizes around the 16 / 32 / 64-symbol window borders, gaps from 0 to 1000, cardinality caps from 1 up, all four
    orientations. (An offline run of 8000 such cases found no difference; this keeps 60 of them in the suite.)"""
    for seed in range(3000, 3015):
        rng = np.random.default_rng(seed)
        text = cases.stress_text(seed, n=int(rng.integers(20000, 60000)), n_dups=int(rng.integers(1, 30)), with_n=bool(rng.integers(0, 2)),
                                 tandem=bool(rng.integers(0, 2)))
        cut = int(rng.integers(1, len(text) - 1))
        frags = [("a", 0, cut), ("b", cut, len(text) - cut)] if rng.random() < 0.7 else [("a", 0, len(text))]
        prep = oracle.Prepared.from_memory(text, frags)
        sa = oracle.best_suffix_array(prep.strand)
        for _ in range(4):
            kw = dict(probe_size=int(rng.choice([8, 9, 10, 12, 15, 16, 17, 20, 24, 31, 32, 33, 40, 48, 64, 65])),
                      gap_size=int(rng.choice([0, 1, 5, 10, 30, 100, 250, 1000])), min_length=int(rng.choice([50, 100, 300, 1000, 2000])),
                      max_cardinality=int(rng.choice([1, 2, 8, 50, 500, 100000])), reverse=bool(rng.integers(0, 2)),
                      complement=bool(rng.integers(0, 2)))
            so, sc = _settings_pair(kw)
            want = oracle.search(prep.strand, sa, prep.chunks, so, 0, threads=2)
            got, ctr = emul_harness.search(prep.strand, sa, prep.chunks, sc)
            assert got == _strip(want.families.as_lists()), (seed, kw)
            assert ctr == want.counters, (seed, kw)


def test_branch_free_byte_codes_equal_the_switches():
    """pack_text_kernel's table-driven byte -> 4-bit code map (kmer_core.h code_of_byte_tab), direct and complemented, on all
    256 byte values; the packing emulation below also runs it next to the plain switch on every byte it packs."""
    assert emul_harness.lib().emul_code_tab_mismatches() == 0


def test_packing_fast_path_equals_the_byte_by_byte_definition():
    """pack_text_kernel (search.cuh) packs 16 bases with byte permutes, four per instruction (kmer_core.h pack16_fast), and
    falls back to one byte at a time for anything else. Both against the plain definition: every length 17..96 (all the
    alignments of the reversed modes and ragged ends), N-runs, '$' and foreign bytes in every position of a word."""
    rng = np.random.default_rng(31)
    fast_total = 0
    for n in list(range(17, 97)) + [1000, 4096, 4097, 4111, 8192 + 5, 70001]:
        t = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=n, p=[0.24, 0.24, 0.24, 0.24, 0.04])
        t[-1] = ord("$")
        fast, bad_fast, bad_tab = emul_harness.pack_check(t)
        assert bad_fast == 0 and bad_tab == 0, n
        fast_total += fast
    assert fast_total > 10000
    base = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=161)
    base[-1] = ord("$")
    for pos in range(0, 160):
        for byte in (ord("$"), ord("a"), 0, 0xFF, ord("B"), ord("U"), 0x4F, 0x40, ord("N")):
            t = base.copy()
            t[pos] = byte
            fast, bad_fast, bad_tab = emul_harness.pack_check(t)
            assert bad_fast == 0 and bad_tab == 0, (pos, byte)


def test_emul_lut_matches_oracle():
    text = cases.stress_text(5, n=30000)
    strand = np.concatenate([text, np.frombuffer(b"$", dtype=np.uint8)])
    sa = oracle.best_suffix_array(strand)
    keys, lo, hi = oracle.lut(strand, sa)
    L = emul_harness.lib()
    glo = np.zeros(ablib.LUT_SIZE, dtype=np.int64)
    ghi = np.zeros(ablib.LUT_SIZE, dtype=np.int64)
    L.emul_lut(strand.ctypes.data, len(strand), sa.ctypes.data, glo.ctypes.data, ghi.ctypes.data)
    # oracle entries are in the reference's enumeration order with LE-u64 keys; ours are base-5 slots (A,C,G,N,T)
    digit = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("N"): 3, ord("T"): 4}
    n_nonempty = 0
    for k, a, b in zip(keys, lo, hi):
        bs = int(k).to_bytes(8, "little")
        slot = 0
        for ch in bs:
            slot = slot * 5 + digit[ch]
        if b > a:
            n_nonempty += 1
            assert (glo[slot], ghi[slot]) == (a, b)
        else:
            assert glo[slot] == ghi[slot]
    assert n_nonempty > 1000


def test_q6_forced_less_near_strand_end():
    """Quirk Q6: suffixes within k-1 of the end compare Less regardless of content (searcher.rs:165-166). Build a text
    whose tail shares 8-mers with many probes so the literal bisection matters, and require equality with the oracle."""
    rng = np.random.default_rng(3)
    unit = kat.rand_dna(rng, 11)
    body = np.tile(unit, 400)                      # 4400 bp of an 11-periodic repeat: every probe shares 8-mers with the tail
    text = np.concatenate([kat.rand_dna(rng, 3000), body])
    strand = np.concatenate([text, np.frombuffer(b"$", dtype=np.uint8)])
    sa = oracle.best_suffix_array(strand)
    for kw in (dict(probe_size=20, gap_size=100, min_length=200, max_cardinality=100000),
               dict(probe_size=20, gap_size=100, min_length=200, max_cardinality=100000, reverse=True, complement=True)):
        so, sc = _settings_pair(kw)
        want = oracle.search(strand, sa, [(0, len(text))], so, 0)
        got, ctr = emul_harness.search(strand, sa, [(0, len(text))], sc)
        assert got == _strip(want.families.as_lists())
        assert ctr == want.counters


def test_device_literal_bucket_equals_reference_sa_search_on_trimmed_indexes():
    """literal_bucket (kmer_core.h; what lut_literal_kernel runs per 8-mer) against the reference's own sa_searchb64
    (oracle/_ref) where present, else the oracle's literal restatement that is pinned to it — on trimmed indexes, where the
    array is not sorted for the text near the cut and only the exact probe sequence gives the reference's answer (Q9)."""
    import ctypes as C
    t = oracle.as_strand(cases.stress_text(17, n=30000, n_dups=12))
    t = np.concatenate([t, np.frombuffer(b"$", dtype=np.uint8)])
    L = emul_harness.lib()
    R = oracle.ref()
    rng = np.random.default_rng(2)
    letters = np.frombuffer(b"ATGCN", dtype=np.uint8)
    checked = nonempty = 0
    for (a, b) in [(0, len(t) - 1), (1000, 20000), (12345, 12399), (29000, len(t) - 1), (7, 8), (5000, 5009)]:
        sa = np.ascontiguousarray(oracle.trimmed_suffix_array(t, (a, b)), dtype=np.int64)
        pats = [bytes(t[i:i + 8]) for i in range(max(a, b - 60), min(b + 9, len(t) - 8))]
        pats += [bytes(t[i:i + 8]) for i in rng.integers(a, max(a + 1, b - 8), 300)]
        pats += [bytes(rng.choice(letters, size=8)) for _ in range(1200)]
        for p in pats:
            if b"$" in p or len(p) != 8:
                continue
            pa = np.frombuffer(p, dtype=np.uint8)
            first, count = C.c_int64(), C.c_int64()
            L.emul_literal_bucket(t.ctypes.data, len(t), pa.ctypes.data, sa.ctypes.data, len(sa), C.byref(first), C.byref(count))
            if R is not None:
                out = C.c_int64()
                cnt = R.sa_searchb64(t.ctypes.data, len(t), pa.ctypes.data, 8, sa.ctypes.data, len(sa), C.byref(out), 0, len(sa))
                want = (out.value, cnt)
            else:
                want = oracle.sa_search_literal(t, p, sa)
            assert (first.value, count.value) == want, (a, b, p)
            checked += 1
            nonempty += count.value > 0
    assert checked > 5000 and nonempty > 500
