"""--trim: the index covers strand[start..stop] only (src/bin/asgart.rs:142-147) while the LUT and every
comparison read the whole strand (quirk Q9). The oracle restates this with the reference's own suffix-array code for the
slice and sa_search step for step (pinned against the reference's sa_searchb64 in tests/test_oracle_ref.py)."""
import json
import os
import subprocess

import numpy as np
import pytest

import asgart_b200 as ab
import oracle
from tests import cases

pytestmark = pytest.mark.gpu


def _osettings(st: ab.RunSettings):
    return oracle.make_settings(probe_size=st.probe_size, gap_size=st.gap_size, min_length=st.min_duplication_length,
                                max_cardinality=st.max_cardinality, reverse=st.reverse, complement=st.complement,
                                skip_masked=st.skip_masked, trim=st.trim)


def test_effective_trim_matches_oracle():
    for trim, n1 in [((5, 100), 50), ((5, 49), 50), ((5, 5), 50), ((9, 3), 50), ((49, 1000), 50), ((60, 1000), 50), ((0, 1), 50)]:
        assert ab.api.effective_trim(trim, n1) == oracle.effective_trim(trim, n1), (trim, n1)


@pytest.mark.parametrize("bits", [32, 64])
def test_trimmed_index_and_lut_equal_reference_bisection(bits):
    """SA of the slice (shifted) and all 5^8 LUT entries, including the buckets the inconsistent order near `stop` shifts."""
    text = cases.stress_text(17, n=30000, n_dups=12)
    strand = np.concatenate([np.frombuffer(text, dtype=np.uint8), np.frombuffer(b"$", dtype=np.uint8)])
    letters = "ACGNT"
    with ab.Context(0) as ctx:
        ctx.set_index_bits(bits)
        ctx.load_strand(strand)
        for trim in [(1000, 20000), (12345, 12399), (29000, 10 ** 9), (7, 8), (0, 30000)]:
            eff = oracle.effective_trim(trim, len(strand))
            ctx.build_index(trim=trim)
            sa = oracle.trimmed_suffix_array(strand, eff)
            assert np.array_equal(ctx.download_sa(), sa), trim
            lo, hi = ctx.download_lut()
            rng = np.random.default_rng(3)
            slots = set(int(x) for x in rng.integers(0, 5 ** 8, 600))
            for x in range(max(eff[0], eff[1] - 40), min(eff[1] + 9, len(strand) - 9)):      # 8-mers around the cut
                slots.add(sum(letters.index(chr(c)) * 5 ** (7 - j) for j, c in enumerate(strand[x:x + 8])))
            slots |= set(int(x) for x in np.nonzero(hi > lo)[0][:400])
            for slot in slots:
                p = bytes(ord(letters[(slot // 5 ** (7 - j)) % 5]) for j in range(8))
                first, cnt = oracle.sa_search_literal(strand, p, sa)
                assert (int(lo[slot]), int(hi[slot])) == (first, first + cnt), (trim, p)
        assert ctx.check_sa() == 0               # (0, n): the slice is the whole strand
        ctx.build_index(trim=(1000, 20000))
        with pytest.raises(ab.AsgartB200Error):
            ctx.check_sa()                       # not a suffix array of the strand
        ctx.build_index(trim=(9, 3))             # the reference skips such a trim: the full index
        assert ctx.check_sa() == 0


@pytest.mark.parametrize("seed", [11, 12])
def test_trimmed_search_equals_oracle(seed):
    text = cases.stress_text(seed)
    prep = ab.Prepared.from_memory(text, [("a", 0, 25000), ("b", 25000, len(text) - 25000)])
    strand = np.array(prep.strand)
    n = len(strand) - 1
    with ab.Context(0) as ctx:
        ctx.load_strand(strand)
        trims = [(n // 3, 2 * n // 3), (0, n // 2), (n // 2, 10 ** 12), (n // 5, n // 5 + 3000), (17, n - 13)]
        for trim in trims:
            eff = oracle.effective_trim(trim, len(strand))
            ctx.build_index(trim=trim)
            for label, kw in cases.settings_grid()[:5]:
                kw = dict(kw)
                kw["min_duplication_length"] = kw.pop("min_length")
                st = ab.RunSettings(trim=trim, **kw)
                for mask in (0, ab.POST_ALL):
                    got = ctx.search(prep.chunks, st, mask).as_lists()
                    want = oracle.search_trim(strand, eff, prep.chunks, _osettings(st), mask, threads=2).as_lists()
                    assert got == want, (trim, label, mask)
        # the slice really restricts the result: right arms lie inside it, and those outside are gone
        half = (n // 2, n)
        st = ab.RunSettings(min_duplication_length=200, trim=half)
        ctx.build_index(trim=half)
        fams = ctx.search(prep.chunks, st, ab.POST_ALL).as_lists()
        ctx.build_index()
        full = ctx.search(prep.chunks, st, ab.POST_ALL).as_lists()
        assert all(half[0] <= sd[1] < half[1] for f in fams for sd in f)
        outside = sum(1 for f in full for sd in f if sd[1] < half[0])
        assert outside > 0 and sum(map(len, fams)) < sum(map(len, full))


def test_trim_through_run_files_and_cli(tmp_path):
    g, fr = ab.synth_genome(2, scale_n=1_500_000)
    fa = tmp_path / "y.fa"
    s = g.tobytes().decode()
    fa.write_text(">synthY\n" + "\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + "\n")
    for trim in [(200_000, 900_000), (1_000_000, 5_000_000), (7, 3)]:
        st = ab.RunSettings(reverse=True, complement=True, trim=trim)
        js = ab.search_duplications([str(fa)], st)
        assert js == oracle.run_files([str(fa)], _osettings(st), threads=4)
        assert json.loads(js)["settings"]["trim"] == list(trim)
    cli = os.path.join(os.path.dirname(ab.__file__), "asgart-b200")
    out = tmp_path / "t.json"
    r = subprocess.run([cli, "-RC", "--trim", "200000", "900000", "--out", str(out), str(fa)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    st = ab.RunSettings(reverse=True, complement=True, trim=(200_000, 900_000))
    assert out.read_text() == oracle.run_files([str(fa)], _osettings(st), threads=4)
    name = ab.out_filename([str(fa)], st)
    assert name.endswith("_RC_200000-900000.json")
