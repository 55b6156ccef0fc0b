"""ComputeScore (--compute-score; src/bin/asgart.rs:98-111, src/structs.rs:439-452) — the oracle's restatement, checked
without a GPU: the edit-distance DP against an independent numpy DP, the arm extraction rules (inclusive ranges, reverse
then complement), the f64 -> f32 identity formula and the two inputs the reference panics on."""
import numpy as np
import pytest

import oracle
from tests import kat


def _np_lev(a: bytes, b: bytes) -> int:
    a = np.frombuffer(a, dtype=np.uint8); b = np.frombuffer(b, dtype=np.uint8)
    prev = np.arange(len(b) + 1, dtype=np.int64)
    for i in range(1, len(a) + 1):
        sub = prev[:-1] + (b != a[i - 1])
        cur = np.minimum(prev[1:] + 1, sub)
        cur = np.concatenate([[i], cur])
        # left-to-right dependency cur[j] = min(cur[j], cur[j-1] + 1) as a running minimum of (cur[j] - j)
        cur = np.minimum.accumulate(cur - np.arange(len(b) + 1)) + np.arange(len(b) + 1)
        prev = cur
    return int(prev[-1])


def test_levenshtein_known_answers():
    assert oracle.levenshtein(b"kitten", b"sitting") == 3
    assert oracle.levenshtein(b"", b"ACGT") == 4
    assert oracle.levenshtein(b"ACGT", b"") == 4
    assert oracle.levenshtein(b"ACGT", b"ACGT") == 0
    assert oracle.levenshtein(b"AAAA", b"TTTT") == 4
    assert oracle.levenshtein(b"ACGTACGT", b"CGTACGTA") == 2


@pytest.mark.parametrize("seed", range(6))
def test_levenshtein_matches_independent_dp(seed):
    rng = np.random.default_rng(seed)
    a = kat.rand_dna(rng, int(rng.integers(1, 400)))
    b = a.copy()
    m = rng.random(len(b)) < 0.1
    b[m] = kat.rand_dna(rng, int(m.sum()))
    if seed % 2:
        p = int(rng.integers(0, len(b)))
        b = np.concatenate([b[:p], kat.rand_dna(rng, 7), b[p:]])
    assert oracle.levenshtein(a.tobytes(), b.tobytes()) == _np_lev(a.tobytes(), b.tobytes())


def _fam(rows, reverse=False, complement=False):
    off = np.array([0, len(rows)], dtype=np.int64)
    fields = np.array(rows, dtype=np.uint64).reshape(-1, 4)
    ident = np.zeros(len(rows), dtype=np.float32)
    flags = np.tile(np.array([[int(reverse), int(complement)]], dtype=np.uint8), (len(rows), 1))
    return oracle.Families(off, fields, ident, flags)


def test_identity_formula_and_arm_rules():
    rng = np.random.default_rng(5)
    t = kat.rand_dna(rng, 4000)
    S = t[500:800].copy()                         # 300 bp
    t[2000:2300] = S
    t[2050] = ord("A") if t[2050] != ord("A") else ord("C")     # one substitution
    strand = np.concatenate([t, np.frombuffer(b"$", dtype=np.uint8)])
    # direct: arms are strand[p ..= p + len] (len + 1 bytes, structs.rs:441-442)
    got = oracle.post_steps(_fam([(500, 2000, 299, 299)]), strand, oracle.POST_COMPUTE_SCORE)
    d = _np_lev(strand[500:800].tobytes(), strand[2000:2300].tobytes())
    assert d == 1
    assert got.identity[0] == np.float32(100.0 * (1.0 - d / 299.0))
    # max(left_length, right_length) is the divisor
    got = oracle.post_steps(_fam([(500, 2000, 299, 199)]), strand, oracle.POST_COMPUTE_SCORE)
    d = _np_lev(strand[500:800].tobytes(), strand[2000:2200].tobytes())
    assert got.identity[0] == np.float32(100.0 * (1.0 - d / 299.0))
    # reversed + complemented: right arm reversed, then complemented
    t2 = t.copy()
    t2[3000:3300] = kat.revcomp(S)
    strand2 = np.concatenate([t2, np.frombuffer(b"$", dtype=np.uint8)])
    got = oracle.post_steps(_fam([(500, 3000, 299, 299)], True, True), strand2, oracle.POST_COMPUTE_SCORE)
    assert got.identity[0] == np.float32(100.0)
    got = oracle.post_steps(_fam([(500, 3000, 299, 299)], True, False), strand2, oracle.POST_COMPUTE_SCORE)
    want = _np_lev(strand2[500:800].tobytes(), strand2[3000:3300][::-1].tobytes())
    assert got.identity[0] == np.float32(100.0 * (1.0 - want / 299.0)) and got.identity[0] < 60


def test_score_sits_between_reduce_overlap_and_sort():
    rng = np.random.default_rng(9)
    t = kat.rand_dna(rng, 6000)
    t[3000:3400] = t[1000:1400]
    strand = np.concatenate([t, np.frombuffer(b"$", dtype=np.uint8)])
    fam = _fam([(1200, 3200, 199, 199), (1000, 3000, 250, 250)])
    got = oracle.post_steps(fam, strand, oracle.POST_ALL | oracle.POST_COMPUTE_SCORE)
    # ReduceOverlap merges the two, the merged duplicon is scored, Sort keeps it
    assert got.as_lists() == [[(1000, 3000, 399, 399, False, False)]]
    assert got.identity[0] == np.float32(100.0)


def test_inputs_the_reference_panics_on():
    strand = np.frombuffer(b"ACGTACGTACGTACGTACGT$", dtype=np.uint8)
    n = len(strand) - 1
    # an arm whose inclusive range ends on '$' is fine without -C ...
    ok = oracle.post_steps(_fam([(0, n - 8, 8, 8)]), strand, oracle.POST_COMPUTE_SCORE)
    assert ok.identity[0] == np.float32(100.0 * (1.0 - 1 / 8.0))
    # ... but complement() knows no '$' (structs.rs:28-34)
    with pytest.raises(oracle.RefPanic):
        oracle.post_steps(_fam([(0, n - 8, 8, 8)], False, True), strand, oracle.POST_COMPUTE_SCORE)
    # and a range past the strand is a slice panic
    with pytest.raises(oracle.RefPanic):
        oracle.post_steps(_fam([(0, n - 8, 8, 9)]), strand, oracle.POST_COMPUTE_SCORE)


def test_filter_ns_panics_past_the_strand_like_the_reference():
    """strand[p ..= p + len] in n_content (src/structs.rs:455-466) is a slice panic when the inclusive range passes '$'."""
    strand = np.frombuffer(b"ACGTACGTACGTACGTACGT$", dtype=np.uint8)
    n = len(strand) - 1
    assert oracle.post_steps(_fam([(0, n - 8, 8, 8)]), strand, oracle.POST_FILTER_NS).as_lists() == [[(0, n - 8, 8, 8, False, False)]]
    with pytest.raises(oracle.RefPanic):
        oracle.post_steps(_fam([(0, n - 8, 8, 9)]), strand, oracle.POST_FILTER_NS)
    with pytest.raises(oracle.RefPanic):
        oracle.post_steps(_fam([(10 ** 12, 0, 8, 8)]), strand, oracle.POST_FILTER_NS)
