"""GPU-side FASTA ingest (fasta_ingest.cuh) against the oracle's restatement of prepare_data
(src/bin/asgart.rs:273-471): strand bytes, fragment map and chunks_to_process must be identical, for files and for
in-memory buffers, and the context must be left exactly as load_strand leaves it."""
import json
import os

import numpy as np
import pytest

import asgart_b200 as ab
import oracle
from tests.test_gpu_parity import _osettings

pytestmark = pytest.mark.gpu


from tests.fasta_cases import fasta as _fasta, line_and_record_blobs, n_run_records, rand_seq as _rand_seq, unparsable_blobs  # noqa: E402


def _check(tmp_path, blobs, skip_masked, tag="f"):
    """ingest from paths and from memory; compare with the oracle's prepare_data on the same files"""
    paths = []
    for i, b in enumerate(blobs):
        p = tmp_path / f"{tag}{i}.fa"
        p.write_bytes(b)
        paths.append(str(p))
    want = oracle.Prepared.from_files(paths, skip_masked)
    with ab.Context(0) as ctx:
        for files, names in ((paths, None), (blobs, paths)):
            got = ctx.ingest(files, skip_masked, names=names)
            assert got.strand is None and got.n1 == len(want.strand)
            assert np.array_equal(ctx.download_strand(), want.strand)
            assert got.map == want.map
            assert got.chunks == want.chunks
            s = ctx.stats()
            assert s["ingest_records"] >= len(want.map)
    return want


@pytest.mark.parametrize("skip_masked", [False, True])
def test_ingest_line_and_record_shapes(tmp_path, skip_masked):
    blobs = line_and_record_blobs()
    for i, blob in enumerate(blobs):
        _check(tmp_path, [blob], skip_masked, tag=f"s{i}_")
    _check(tmp_path, blobs[:4], skip_masked, tag="multi")                            # several files: running offset


def test_ingest_n_runs_and_chunks(tmp_path):
    rng = np.random.default_rng(9)
    s = lambda n: _rand_seq(rng, n)   # noqa: E731
    recs = n_run_records()
    for width in (60, 4096, 10 ** 9):
        want = _check(tmp_path, [_fasta(recs, width=width)], False, tag=f"n{width}_")
        assert len(want.chunks) > len(recs)
    _check(tmp_path, [_fasta(recs[:5]), _fasta(recs[5:])], True, tag="two")
    # masked stretches become N-runs under -S and split chunks there
    masked = [("m", s(4000) + s(5200).lower() + s(3000) + s(4999).lower() + s(100))]
    w0 = _check(tmp_path, [_fasta(masked)], False, tag="m0")
    w1 = _check(tmp_path, [_fasta(masked)], True, tag="m1")
    assert len(w0.chunks) == 1 and len(w1.chunks) == 2


def test_ingest_errors(tmp_path):
    bad = tmp_path / "bad.fa"
    bad.write_bytes(b"ACGT\n>late header\nACGT\n")
    with ab.Context(0) as ctx:
        with pytest.raises(ab.AsgartB200Error, match="Unable to parse"):
            ctx.ingest([str(bad)])
        with pytest.raises(ab.AsgartB200Error, match="Unable to parse"):
            ctx.ingest([b"\r\n>x\nAC\n"])
        for i, blob in enumerate(unparsable_blobs()):     # anything but '>' as the first byte, a leading blank line included
            u = tmp_path / f"u{i}.fa"
            u.write_bytes(blob)
            with pytest.raises(ab.AsgartB200Error, match="Unable to parse"):
                ctx.ingest([str(u)])
            with pytest.raises(ab.AsgartB200Error, match="Unable to parse"):
                ctx.ingest([blob])
            with pytest.raises(IOError, match="Unable to parse"):
                oracle.Prepared.from_files([str(u)])
        with pytest.raises(ab.AsgartB200Error, match="Unable to read FASTA file"):
            ctx.ingest([str(tmp_path / "missing.fa")])
        with pytest.raises(ab.AsgartB200Error):
            ctx.build_index()                                                         # no strand after a failed ingest
        with pytest.raises(IOError):
            oracle.Prepared.from_files([str(bad)])
        ok = ctx.ingest([b">x\nACGTACGTAC\n"])
        assert ok.map == [("x", 0, 10)] and ok.chunks == [(0, 10)]
        assert ctx.download_strand().tobytes() == b"ACGTACGTAC$"


def test_ingest_empty_record_ends_the_file(tmp_path):
    """The bio reader's iterator stops at the first record without id, description and sequence (fasta_core.h)."""
    big = b">" + b"N" * 6000 + b"\nAC\n>\n>b\n" + b"N" * 7000 + b"\n"
    want = _check(tmp_path, [b">a\nAC\n>\n>b\nGG\n", big, b">\n>x\nAC\n", b">c\nGG\n"], False, tag="er")
    assert [m[0] for m in want.map] == ["a", "N" * 6000, "c"] and want.strand.tobytes() == b"ACACGG$"


def test_ingest_then_search_equals_host_prepared_path(tmp_path):
    """Synthetic C2-shaped genome (soft-masked, N-runs) written as a 60-column FASTA plus a second small file: the
    ingested context gives the same index, families and JSON as the host-prepared one and as the oracle."""
    g, fr = ab.synth_genome(2, scale_n=3_000_000)
    text = g.tobytes().decode()
    fa = tmp_path / "synthY.fa"
    fa.write_bytes(_fasta([("synthY desc", text)]))
    rng = np.random.default_rng(3)
    fb = tmp_path / "extra.fa"
    fb.write_bytes(_fasta([("e1", _rand_seq(rng, 70000)), ("e2", text[100000:160000])], width=80))
    files = [str(fa), str(fb)]
    for kw in (dict(reverse=True, complement=True, skip_masked=True), dict()):
        st = ab.RunSettings(**kw)
        want_prep = oracle.Prepared.from_files(files, st.skip_masked)
        sa = oracle.best_suffix_array(want_prep.strand)
        want = oracle.search(want_prep.strand, sa, want_prep.chunks, _osettings(st), oracle.POST_ALL, threads=4)
        with ab.Context(0) as ctx:
            prep = ctx.ingest(files, st.skip_masked)
            assert prep.map == want_prep.map and prep.chunks == want_prep.chunks
            ctx.build_index()
            assert np.array_equal(ctx.download_sa(), sa)
            got = ctx.search(prep.chunks, st, ab.POST_ALL)
            assert got.as_lists() == want.families.as_lists()
            assert got.n_families > 0
            assert prep.to_json(st, got) == want_prep.to_json(_osettings(st), want.families)
            s = ctx.stats()
            assert s["ingest_bytes"] == os.path.getsize(fa) + os.path.getsize(fb)
        js = ab.search_duplications(files, st)                                        # run_files goes through the ingest
        assert js == oracle.run_files(files, _osettings(st), threads=4)
        assert json.loads(js)["strand"]["name"] == ", ".join(files)


def test_c5_cross_genome_two_files(tmp_path):
    """BASELINE configs[4] scaled down: `asgart A.fa B.fa -k 32 -g 200 -RC` on two 10-fragment genomes that share planted
    segments (3 Mbp each here). Two files -> running offset, 20 fragments, probe_size 32 (two-word window compares),
    step 16, max_gap 232; families identical to the oracle's and some duplicons pair a fragment of A with one of B."""
    files = []
    for part, tag in ((0, "A"), (1, "B")):
        g, fr = ab.synth_genome(5, part=part, scale_n=3_000_000)
        text = g.tobytes().decode()
        p = tmp_path / f"genome{tag}.fa"
        p.write_bytes(_fasta([(nm + " synthetic", text[pos:pos + ln]) for nm, pos, ln in fr]))
        files.append(str(p))
    st = ab.RunSettings(probe_size=32, gap_size=200, reverse=True, complement=True)
    want_prep = oracle.Prepared.from_files(files, False)
    sa = oracle.best_suffix_array(want_prep.strand)
    want = oracle.search(want_prep.strand, sa, want_prep.chunks, _osettings(st), oracle.POST_ALL, threads=4)
    with ab.Context(0) as ctx:
        prep = ctx.ingest(files, False)
        assert len(prep.map) == 20 and prep.map == want_prep.map and prep.chunks == want_prep.chunks
        ctx.build_index()
        assert ctx.check_sa() == 0
        got = ctx.search(prep.chunks, st, ab.POST_ALL)
        assert got.as_lists() == want.families.as_lists()
        half = prep.map[10][1]                                                     # first base of genome B
        cross = sum(1 for fam in got.as_lists() for sd in fam if sd[0] < half <= sd[1])
        assert cross > 0
        js = prep.to_json(st, got)
        assert js == want_prep.to_json(_osettings(st), want.families)
        assert json.loads(js)["settings"]["max_gap_size"] == 232
        # the direct pass over the same index (what asgart-slice would merge with the -RC run)
        st_d = ab.RunSettings(probe_size=32, gap_size=200)
        got_d = ctx.search(prep.chunks, st_d, ab.POST_ALL)
        want_d = oracle.search(want_prep.strand, sa, want_prep.chunks, _osettings(st_d), oracle.POST_ALL, threads=4)
        assert got_d.as_lists() == want_d.families.as_lists()


def test_passes_over_one_index_combine_like_asgart_slice(tmp_path):
    """Direct + -RC passes over one index = RunResult::from_files (src/structs.rs:114-141) of the two runs' JSON files:
    strand and settings of the first, families concatenated in input order. Also through the CLI (--with-direct)."""
    import subprocess
    g, fr = ab.synth_genome(2, scale_n=1_500_000)
    fa = tmp_path / "y.fa"
    fa.write_bytes(_fasta([("synthY", g.tobytes().decode())]))
    files = [str(fa)]
    rc, direct = ab.RunSettings(reverse=True, complement=True, skip_masked=True), ab.RunSettings(skip_masked=True)
    runs = [json.loads(oracle.run_files(files, _osettings(st), threads=4)) for st in (rc, direct)]
    want = {"strand": runs[0]["strand"], "settings": runs[0]["settings"], "families": runs[0]["families"] + runs[1]["families"]}
    got = json.loads(ab.search_duplications_passes(files, [rc, direct]))
    assert got == want
    assert len(runs[0]["families"]) > 0 and len(runs[1]["families"]) > 0
    assert {sd["reversed"] for f in got["families"] for sd in f} == {True, False}
    with pytest.raises(ab.AsgartB200Error, match="skip_masked"):
        ab.search_duplications_passes(files, [rc, ab.RunSettings()])
    cli = os.path.join(os.path.dirname(ab.__file__), "asgart-b200")
    out = tmp_path / "both.json"
    r = subprocess.run([cli, "-RCS", "--with-direct", "--out", str(out), str(fa)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert json.loads(out.read_text()) == want
