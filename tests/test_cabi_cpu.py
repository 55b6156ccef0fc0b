"""CPU-side checks of the C ABI and the host mirror: the library loads, exports every symbol include/asgart_b200.h
declares, refuses to compute without a device (no fallback), and its host side (prepare_data, JSON, file naming, synthetic
genomes) agrees with the oracle."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import asgart_b200 as ab
import oracle
from asgart_b200 import _lib
from tests import cases, kat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "asgart_b200.h")).read()
    declared = set(re.findall(r"\b(asgart_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS)
    L = _lib.load()
    for s in declared:
        assert hasattr(L, s), s
    assert b"asgart_b200" in L.asgart_b200_version()


def test_struct_layouts():
    assert C.sizeof(_lib.ProtoSD) == 40 and C.sizeof(_lib.Chunk) == 16
    assert C.sizeof(_lib.Settings) == 64
    assert C.sizeof(oracle.Settings) == C.sizeof(_lib.Settings)


@pytest.mark.skipif(ab.device_count() > 0, reason="checks the no-device behaviour")
def test_no_cpu_fallback():
    with pytest.raises(ab.AsgartB200Error) as e:
        ab.Context(0)
    assert e.value.code == _lib.ENODEVICE
    t = np.frombuffer(b"ACGTACGT$", dtype=np.uint8)
    sa = np.zeros(len(t), dtype=np.int64)
    assert _lib.load().asgart_b200_divsufsort64(t.ctypes.data, sa.ctypes.data, len(t)) == _lib.ENODEVICE
    assert _lib.load().asgart_b200_divsufsort64(None, sa.ctypes.data, 4) == _lib.EINVAL  # divsufsort.c:337 contract


def _write_fasta(path, recs, width=60, crlf=False):
    nl = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as f:
        for name, seq in recs:
            f.write(">" + name + nl)
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width] + nl)


@pytest.mark.parametrize("skip_masked", [False, True])
def test_prepare_data_matches_oracle(tmp_path, skip_masked):
    rng = np.random.default_rng(1)
    a = cases.stress_text(21, n=30000).tobytes().decode()
    b = kat.rand_dna(rng, 9000).tobytes().decode()
    b = b[:2000] + b[2000:4000].lower() + "RYKM-*x" + b[4000:]            # soft-masked stretch + IUPAC/garbage -> N
    c = "N" * 6000                                                       # all-N fragment -> one chunk
    f1, f2 = str(tmp_path / "one.fa"), str(tmp_path / "two.fasta")
    _write_fasta(f1, [("chrA desc here", a), ("chrB", b)], crlf=True)
    _write_fasta(f2, [("chrC\tx", c), ("chrD", b[:50])], width=70)
    po = oracle.Prepared.from_files([f1, f2], skip_masked)
    pp = ab.prepare_data([f1, f2], skip_masked)
    assert np.array_equal(po.strand, pp.strand)
    assert po.chunks == pp.chunks
    assert po.map == pp.map
    assert [m[0] for m in pp.map] == ["chrA", "chrB", "chrC", "chrD"]
    assert pp.strand[-1] == ord("$")
    # in-memory path + vectorised normalise agree with the file path
    raw = np.frombuffer((a + b + c + b[:50]).encode(), dtype=np.uint8)
    pm = ab.Prepared.from_memory(ab.normalise(raw, skip_masked), pp.map)
    assert np.array_equal(pm.strand, pp.strand) and pm.chunks == pp.chunks


def test_missing_file_errors():
    with pytest.raises(IOError):
        ab.prepare_data(["/nonexistent/x.fa"])
    with pytest.raises(IOError):
        oracle.Prepared.from_files(["/nonexistent/x.fa"])


def test_json_matches_oracle_bytes():
    rng = np.random.default_rng(2)
    text = kat.rand_dna(rng, 5000)
    frags = [("chr \"q\"\\1", 0, 3000), ("chrZ", 3000, 2000)]
    po = oracle.Prepared.from_memory(text, frags, "a.fa, b.fa")
    pp = ab.Prepared.from_memory(text, frags, "a.fa, b.fa")
    fams = [[(10, 3100, 1000, 1001), (20, 4999, 5, 1)], [(5000, 7000, 1, 1)], [(0, 0, 0, 0)]]
    for kw in (dict(), dict(reverse=True, complement=True, skip_masked=True)):
        st = ab.RunSettings(**kw)
        so = oracle.make_settings(reverse=st.reverse, complement=st.complement, skip_masked=st.skip_masked)
        fa = ab.families_from_lists(fams, st.reverse, st.complement)
        fo = oracle.Families(fa.fam_offsets.astype(np.int64),
                             np.array([[s["left"], s["right"], s["left_length"], s["right_length"]] for s in fa.sds], dtype=np.uint64),
                             np.zeros(len(fa.sds), np.float32),
                             np.array([[s["reversed"], s["complemented"]] for s in fa.sds], dtype=np.uint8))
        js = pp.to_json(st, fa)
        assert js == po.to_json(so, fo)
        d = json.loads(js)
        assert d["families"][1][0]["chr_left"] == "unknown" and d["families"][1][0]["chr_left_position"] == 5000
    empty = ab.families_from_lists([])
    assert pp.to_json(ab.RunSettings(), empty) == po.to_json(
        oracle.make_settings(), oracle.Families(np.zeros(1, np.int64), np.zeros((0, 4), np.uint64), np.zeros(0, np.float32), np.zeros((0, 2), np.uint8)))
    st = ab.RunSettings(trim=(5, 10))
    assert '"trim": [\n      5,\n      10\n    ],' in pp.to_json(st, empty)


def test_out_filename_rule():
    S = ab.RunSettings
    assert ab.out_filename(["/data/chrY.fa"], S()) == "chrY.json"
    assert ab.out_filename(["/data/chrY.fa", "x/chr1.v2.fasta"], S(reverse=True, complement=True), prefix="run_") == "run_chrY-chr1.v2_RC.json"
    assert ab.out_filename(["a.fa"], S(complement=True)) == "a_C.json"
    assert ab.out_filename(["a.fa"], S(reverse=True, trim=(3, 9))) == "a_R_3-9.json"
    assert ab.out_filename(["a.fa"], S(), out="out/result.txt") == "out/result.json"
    assert ab.out_filename(["a.fa"], S(), out="res") == "res.json"


def test_synth_deterministic_and_shaped():
    g1, fr = ab.synth_genome(2, scale_n=600_000)
    g2, _ = ab.synth_genome(2, scale_n=600_000, threads=3)
    assert np.array_equal(g1, g2) and fr == [("synthY", 0, 600_000)]
    assert set(np.unique(g1)) <= set(b"ACGTacgtN")
    low = (g1 >= ord("a")).mean()
    assert 0.08 < low < 0.22                       # ~15 % soft-masked
    assert (g1[:5001] == ord("N")).all() and (g1[-5001:] == ord("N")).all()   # telomeric N-runs
    g4, fr4 = ab.synth_genome(4, scale_n=3_000_000)
    assert len(fr4) == 24 and fr4[0][0] == "chr1" and fr4[-1][0] == "chrY" and sum(f[2] for f in fr4) == len(g4)
    assert ab._lib.load().asgart_b200_synth_length(4, 0, 0) == 3_088_269_832
    assert ab._lib.load().asgart_b200_synth_length(2, 0, 0) == 57_227_415
    # the base formula of DESIGN.md: base[i] = "ACGT"[splitmix64(seed * 0x9E3779B97F4A7C15 + i) >> 62] (config 0 = no extras)
    g0, _ = ab.synth_genome(0, scale_n=1000, seed=7, n_pairs=0)
    M = (1 << 64) - 1
    def sm(x):
        z = (x + 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)
    want = bytes(b"ACGT"[sm((7 * 0x9E3779B97F4A7C15 + i) & M) >> 62] for i in range(1000))
    assert g0.tobytes() == want


def test_synth_planted_pairs_found_by_oracle():
    """The generator's planted duplications are what the reference algorithm reports (direct run on a C1-shaped input)."""
    g, fr = ab.synth_genome(1, scale_n=1_000_000)
    prep = ab.Prepared.from_memory(ab.normalise(g, False), fr)
    sa = oracle.best_suffix_array(np.array(prep.strand))
    out = oracle.search(np.array(prep.strand), sa, prep.chunks, oracle.make_settings(), oracle.POST_ALL, threads=2)
    assert len(out.families.as_lists()) >= 2


def test_effective_trim_matches_the_oracle():
    """prepare_data's --trim validation (src/bin/asgart.rs:432-463) is host code: C ABI vs the oracle's restatement."""
    import oracle
    from asgart_b200.api import effective_trim
    cases = [((5, 100), 50), ((5, 49), 50), ((5, 5), 50), ((9, 3), 50), ((49, 1000), 50), ((60, 1000), 50), ((0, 1), 50),
             ((0, 0), 1), ((0, 5), 1), ((3, 2 ** 40), 2 ** 33)]
    for trim, n1 in cases:
        assert effective_trim(trim, n1) == oracle.effective_trim(trim, n1), (trim, n1)
    assert effective_trim((5, 100), 50) == (5, 49) and effective_trim((9, 3), 50) is None
