"""ComputeScore on the GPU (levenshtein_kernel through the C ABI, ASGART_B200_POST_COMPUTE_SCORE) against the oracle's
restatement of ProtoSD::levenshtein (src/structs.rs:439-452). Distances are integers and the identity is one f64
expression cast to f32, so the bar is bit-exact (the north star allows 1e-6)."""
import numpy as np
import pytest

import asgart_b200 as ab
from asgart_b200 import _lib
import oracle
from tests import cases, kat

pytestmark = pytest.mark.gpu


def _ofam(fam: ab.Families) -> oracle.Families:
    s = fam.sds
    fields = np.stack([s["left"], s["right"], s["left_length"], s["right_length"]], axis=1).astype(np.uint64)
    flags = np.stack([s["reversed"], s["complemented"]], axis=1).astype(np.uint8)
    return oracle.Families(fam.fam_offsets.astype(np.int64), fields, s["identity"].astype(np.float32), flags)


def _mutated_copy(rng, seg, rate, indels):
    seg = seg.copy()
    m = rng.random(len(seg)) < rate
    seg[m] = kat.rand_dna(rng, int(m.sum()))
    for _ in range(indels):
        p = int(rng.integers(1, len(seg) - 1))
        if rng.random() < 0.5:
            seg = np.concatenate([seg[:p], kat.rand_dna(rng, int(rng.integers(1, 4))), seg[p:]])
        else:
            seg = np.concatenate([seg[:p], seg[p + int(rng.integers(1, 4)):]])
    return seg


@pytest.mark.parametrize("flags", [(False, False), (True, False), (False, True), (True, True)])
def test_score_matches_oracle_over_strip_and_word_boundaries(flags):
    """Arm lengths around the 64-row word and 2048-row strip boundaries, pairs with more strips than warps in a block
    (17 408 rows = 9 strips), unequal arms, a pair of unrelated arms (distance near the maximum)."""
    rev, comp = flags
    rng = np.random.default_rng(40 + 2 * rev + comp)
    lens = [1, 2, 31, 62, 63, 64, 65, 127, 128, 500, 2046, 2047, 2048, 2049, 4095, 4096, 4100, 6000, 17408, 20001]
    parts, rows, pos = [], [], 0

    def put(seg):
        nonlocal pos
        parts.append(seg)
        start = pos
        pos += len(seg)
        return start

    put(kat.rand_dna(rng, 100))
    for L in lens:
        src = kat.rand_dna(rng, L + 1)
        dst = _mutated_copy(rng, src, 0.02, 0 if L < 100 else 3)
        if rev:
            dst = dst[::-1].copy()
        if comp:
            dst = kat.complement(dst)
        a = put(src); put(kat.rand_dna(rng, 7)); b = put(dst); put(kat.rand_dna(rng, 5))
        rows.append((a, b, L, len(dst) - 1))
    a = put(kat.rand_dna(rng, 3000)); b = put(kat.rand_dna(rng, 2500)); put(kat.rand_dna(rng, 50))
    rows.append((a, b, 2999, 2400))                                   # unrelated arms
    rows.append((rows[5][0], rows[5][1], rows[5][2], 5))              # very unequal lengths
    strand = np.concatenate(parts + [np.frombuffer(b"$", dtype=np.uint8)])
    fam = ab.families_from_lists([rows[:7], rows[7:]], reverse=rev, complement=comp)
    want = oracle.post_steps(_ofam(fam), strand, oracle.POST_COMPUTE_SCORE)
    with ab.Context(0) as ctx:
        ctx.load_strand(strand)
        got = ctx.post_steps(fam, ab.POST_COMPUTE_SCORE)
        assert got.as_lists() == want.as_lists()
        assert np.array_equal(got.sds["identity"].view(np.uint32), want.identity.view(np.uint32)), \
            (got.sds["identity"], want.identity)
        s = ctx.stats()
        assert s["score_pairs"] == len(rows) and s["score_cells"] == sum((r[2] + 1) * (r[3] + 1) for r in rows)


def test_score_with_n_and_terminator_symbols():
    rng = np.random.default_rng(77)
    t = kat.rand_dna(rng, 9000)
    t[1000:1040] = ord("N"); t[5020:5030] = ord("N")
    t[5000:6500] = _mutated_copy(rng, t[1000:2500], 0.01, 0)
    t[8000:9000] = t[3000:4000]                      # right arm's inclusive range ends on '$'
    strand = np.concatenate([t, np.frombuffer(b"$", dtype=np.uint8)])
    fam = ab.families_from_lists([[(1000, 5000, 1499, 1499), (3000, 8000, 1000, 1000)]])
    want = oracle.post_steps(_ofam(fam), strand, oracle.POST_COMPUTE_SCORE)
    with ab.Context(0) as ctx:
        ctx.load_strand(strand)
        got = ctx.post_steps(fam, ab.POST_COMPUTE_SCORE)
        assert np.array_equal(got.sds["identity"].view(np.uint32), want.identity.view(np.uint32))
        # the two inputs the reference panics on are errors here as well, with the reason in the message
        with pytest.raises(ab.AsgartB200Error, match="terminator"):
            ctx.post_steps(ab.families_from_lists([[(3000, 8000, 1000, 1000)]], complement=True), ab.POST_COMPUTE_SCORE)
        with pytest.raises(ab.AsgartB200Error, match="past the strand"):
            ctx.post_steps(ab.families_from_lists([[(3000, 8000, 1000, 1001)]]), ab.POST_COMPUTE_SCORE)
        assert ctx.post_steps(ab.families_from_lists([]), ab.POST_COMPUTE_SCORE).n_families == 0
        # FilterNs reads the same inclusive ranges (src/structs.rs:455-466): past the strand is the reference's slice panic,
        # not an out-of-bounds device read — and the context stays usable afterwards
        for bad in ((3000, 8000, 1000, 1001), (10 ** 15, 8000, 10, 10), (3000, 2 ** 63, 10, 2 ** 63)):
            with pytest.raises(oracle.RefPanic):
                oracle.post_steps(_ofam(ab.families_from_lists([[bad]])), strand, oracle.POST_FILTER_NS)
            with pytest.raises(ab.AsgartB200Error, match="past the strand") as ei:
                ctx.post_steps(ab.families_from_lists([[bad]]), ab.POST_FILTER_NS)
            assert ei.value.code == _lib.EPANIC
        assert ctx.post_steps(fam, ab.POST_FILTER_NS).as_lists() == oracle.post_steps(_ofam(fam), strand, oracle.POST_FILTER_NS).as_lists()


@pytest.mark.parametrize("seed", [11, 13])
def test_search_with_compute_score_equals_oracle(seed):
    """The whole pipeline with --compute-score: SearchDuplications, FilterNs, ReOrder, ReduceOverlap, ComputeScore, Sort."""
    text = cases.stress_text(seed)
    prep = ab.Prepared.from_memory(text, [("a", 0, 25000), ("b", 25000, len(text) - 25000)])
    strand = np.array(prep.strand)
    sa = oracle.best_suffix_array(strand)
    with ab.Context(0) as ctx:
        ctx.load_strand(strand)
        ctx.build_index()
        n_scored = 0
        for kw in (dict(), dict(reverse=True, complement=True), dict(reverse=True), dict(complement=True)):
            st = ab.RunSettings(min_duplication_length=400, compute_score=True, **kw)
            got = ctx.search(prep.chunks, st)
            so = oracle.make_settings(min_length=400, reverse=st.reverse, complement=st.complement)
            want = oracle.search(strand, sa, prep.chunks, so, oracle.POST_ALL | oracle.POST_COMPUTE_SCORE, threads=2).families
            assert got.as_lists() == want.as_lists(), kw
            assert np.array_equal(got.sds["identity"].view(np.uint32), want.identity.view(np.uint32)), kw
            n_scored += len(got.sds)
            plain = ctx.search(prep.chunks, ab.RunSettings(min_duplication_length=400, **kw))
            assert plain.as_lists() == got.as_lists() and not plain.sds["identity"].any()
        assert n_scored > 0


def test_json_carries_the_identity(tmp_path):
    g, fr = ab.synth_genome(2, scale_n=800_000)
    fa = tmp_path / "score.fa"
    with open(fa, "wb") as f:
        for name, pos, ln in fr:
            f.write(b">" + name.encode() + b"\n")
            seq = g[pos:pos + ln]
            for i in range(0, ln, 60):
                f.write(seq[i:i + 60].tobytes() + b"\n")
    best = 0.0
    for kw in (dict(), dict(reverse=True, complement=True)):
        st = ab.RunSettings(compute_score=True, **kw)
        got = ab.search_duplications([str(fa)], st)
        prep = oracle.Prepared.from_files([str(fa)], False)
        so = oracle.make_settings(**kw)
        sa = oracle.best_suffix_array(prep.strand)
        fam = oracle.search(prep.strand, sa, prep.chunks, so, oracle.POST_ALL | oracle.POST_COMPUTE_SCORE, threads=2).families
        assert got == prep.to_json(so, fam)
        if len(fam.identity):
            best = max(best, float(fam.identity.max()))
    assert best > 90
