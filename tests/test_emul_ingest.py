"""The arithmetic of the GPU-side FASTA ingest (asgart_b200/csrc/fasta_core.h: four-bytes-per-instruction classification,
keep / base masks, per-thread counts, N-run reporting, chunk derivation) compiled for the host and run one "thread" at a
time against the oracle's restatement of prepare_data (src/bin/asgart.rs:273-471). The warp / tile plumbing of the kernels
is covered by tests/test_gpu_ingest.py on the same inputs."""
import numpy as np
import pytest

import asgart_b200 as ab
import oracle
from tests import emul_harness
from tests.fasta_cases import fasta, line_and_record_blobs, n_run_records, rand_seq, unparsable_blobs


def _check(tmp_path, blob, skip_masked, tag):
    p = tmp_path / f"{tag}.fa"
    p.write_bytes(blob)
    want = oracle.Prepared.from_files([str(p)], skip_masked)
    strand, frags, chunks = emul_harness.ingest(blob, skip_masked)
    assert np.array_equal(strand, want.strand[:-1]), tag
    assert frags == want.map, tag
    assert chunks == want.chunks, tag
    return want


@pytest.mark.parametrize("skip_masked", [False, True])
def test_emul_ingest_line_and_record_shapes(tmp_path, skip_masked):
    for i, blob in enumerate(line_and_record_blobs()):
        _check(tmp_path, blob, skip_masked, f"s{i}")


def test_emul_ingest_unparsable_files(tmp_path):
    for i, blob in enumerate(unparsable_blobs()):
        p = tmp_path / f"u{i}.fa"
        p.write_bytes(blob)
        with pytest.raises(IOError, match="Unable to parse"):
            oracle.Prepared.from_files([str(p)])
        with pytest.raises(IOError, match="Unable to parse"):
            emul_harness.ingest(blob, False)


def test_emul_ingest_empty_record_ends_the_file(tmp_path):
    want = _check(tmp_path, b">a\nAC\n>\n>b\nGG\n", False, "e0")
    assert want.map == [("a", 0, 2)] and want.strand.tobytes() == b"AC$"
    want = _check(tmp_path, b">a\nAC\n> d\n>b\nGG\n", False, "e1")
    assert want.map == [("a", 0, 2), ("", 2, 0), ("b", 2, 2)]
    want = _check(tmp_path, b">\n>b\nGG\n", False, "e2")
    assert want.map == [] and want.strand.tobytes() == b"$"


@pytest.mark.parametrize("skip_masked", [False, True])
def test_host_prepare_files_equals_oracle(tmp_path, skip_masked):
    """asgart_b200_prepare_files (host.cpp: the host-side read_fasta + find_chunks_to_process, no GPU involved) on the same
    files, several at once included, and the same refusals."""
    paths = []
    for i, blob in enumerate(line_and_record_blobs() + [fasta(n_run_records())]):
        p = tmp_path / f"h{i}.fa"
        p.write_bytes(blob)
        paths.append(str(p))
    for files in [[f] for f in paths] + [paths[:4], paths]:
        want = oracle.Prepared.from_files(files, skip_masked)
        got = ab.Prepared.from_files(files, skip_masked)
        assert np.array_equal(got.strand, want.strand) and got.map == want.map and got.chunks == want.chunks, files
    for i, blob in enumerate(unparsable_blobs()):
        p = tmp_path / f"hu{i}.fa"
        p.write_bytes(blob)
        with pytest.raises(IOError, match="Unable to parse"):
            ab.Prepared.from_files([paths[0], str(p)], skip_masked)


def test_emul_ingest_n_runs_and_chunks(tmp_path):
    recs = n_run_records()
    for width in (60, 31, 32, 33, 4096, 10 ** 9):
        want = _check(tmp_path, fasta(recs, width=width), False, f"n{width}")
        assert len(want.chunks) > len(recs)
    _check(tmp_path, fasta(recs, eol="\r\n"), True, "crlf")
    rng = np.random.default_rng(9)
    s = lambda n: rand_seq(rng, n)   # noqa: E731
    masked = fasta([("m", s(4000) + s(5200).lower() + s(3000) + s(4999).lower() + s(100))])
    assert len(_check(tmp_path, masked, False, "m0").chunks) == 1
    assert len(_check(tmp_path, masked, True, "m1").chunks) == 2


def test_emul_ingest_random_files(tmp_path):
    """Random mixes of bases, lower case, blanks, line ends and header starts at every alignment inside the 32-byte threads."""
    rng = np.random.default_rng(21)
    alphabet = np.frombuffer(b"ACGTacgtNn \t\r\n\n\n>xX-", dtype=np.uint8)
    weights = np.array([8, 8, 8, 8, 2, 2, 2, 2, 3, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1], dtype=float)
    for i in range(60):
        n = int(rng.integers(1, 3000))
        body = rng.choice(alphabet, size=n, p=weights / weights.sum()).tobytes()
        _check(tmp_path, b">r" + str(i).encode() + b" d\n" + body, bool(i & 1), f"r{i}")


def test_random_header_heavy_files_three_ways(tmp_path):
    """Short random files dense in '>' and blanks (empty records, blank headers, refusals): the oracle, the host-side
    prepare_files and the emulated device passes must agree on strand, map, chunks and on which files are refused."""
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"ACGTacgtNn \t\r\n\n\n>xX-", dtype=np.uint8)
    weights = np.array([8, 8, 8, 8, 2, 2, 2, 2, 3, 1, 1, 1, 1, 2, 2, 6, 1, 1, 1, 1], dtype=float)
    refused = stopped_early = 0
    for i in range(400):
        body = rng.choice(alphabet, size=int(rng.integers(0, 400)), p=weights / weights.sum()).tobytes()
        blob = (b">" if rng.random() < 0.9 else b"") + body
        p = tmp_path / "r.fa"
        p.write_bytes(blob)
        sm = bool(i & 1)
        try:
            want = oracle.Prepared.from_files([str(p)], sm)
        except IOError:
            refused += 1
            with pytest.raises(IOError):
                emul_harness.ingest(blob, sm)
            with pytest.raises(IOError):
                ab.Prepared.from_files([str(p)], sm)
            continue
        strand, frags, chunks = emul_harness.ingest(blob, sm)
        assert np.array_equal(strand, want.strand[:-1]) and frags == want.map and chunks == want.chunks, blob
        host = ab.Prepared.from_files([str(p)], sm)
        assert np.array_equal(host.strand, want.strand) and host.map == want.map and host.chunks == want.chunks, blob
        stopped_early += len(want.map) < blob.count(b"\n>") + 1
    assert refused > 10 and stopped_early > 3
