"""FASTA inputs shared by the ingest tests (GPU: tests/test_gpu_ingest.py, CPU emulation: tests/test_emul_ingest.py)."""
import numpy as np


def fasta(records, width=60, eol="\n", final_eol=True) -> bytes:
    out = []
    for name, seq in records:
        out.append(">" + name)
        for i in range(0, len(seq), width):
            out.append(seq[i:i + width])
    s = eol.join(out)
    return (s + (eol if final_eol else "")).encode()


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(list(alphabet), size=n))


def line_and_record_blobs():
    """Line / record shapes: CRLF, missing final newline, one-line records, interior and trailing blanks, empty records,
    '>' inside a line, every byte value (bytes >= 0x80, control characters), blanks before random newlines."""
    rng = np.random.default_rng(7)
    a = rand_seq(rng, 1000, "ACGTacgtNnRYKM-*xX")
    b = rand_seq(rng, 7321, "ACGTacgt")
    blobs = [
        fasta([("chr1 some description", a), ("chr2", b)]),
        fasta([("chr1\tdesc", a), ("chr2", b)], eol="\r\n"),
        fasta([("x", a), ("y", b)], final_eol=False),
        fasta([("x", a)], width=10 ** 9),
        b">interior empty lines\nACGT\nAC GT  \t\nGG\r\n\n\nTT \n>e1\n>e2 d\n\n>last\nNNNN>AC\n ACGT\nA",
        b">only header",
        b">only header\n",
        b">\nACGT\n",
        b"",
        b">a\n   \n \t \n>b\n\x0b\x0cAC\x0b\x0c\n",
    ]
    # the bio reader's iterator ends at the first EMPTY record (no id, no description, no sequence); records that have any
    # of the three go on
    blobs += [
        b">a\nAC\n>\n>b\nGG\n",                         # only `a`
        b">\n>b\nGG\n",                                 # nothing at all
        b">a\nACGT\n>  \t\r\n\n  \n \t\n>b\nGG",          # blank header, blank lines: still empty -> only `a`
        b">a\nAC\n> d\n>b\nGG\n",                       # a description is enough: a, "", b
        b">a\nAC\n>\nT\n>b\nGG\n",                      # a sequence is enough
        b">a\nAC\n>",                                   # empty record at the very end
        b">a\n>b\n>\n>c\nAC\n",                         # empty NAMED records stay: a, b
        b">" + b"N" * 6000 + b"\nAC\n>\n>b\n" + b"N" * 7000 + b"\n",
    ]
    blobs.append(b">bin\n" + rng.integers(0, 256, size=20000, dtype=np.uint8).tobytes().replace(b"\n>", b"\n?"))   # ids must stay UTF-8
    blobs.append(b">runs\n" + b"".join(bytes([v]) * 37 for v in range(256)) + b"\n" + bytes(range(256)) * 3)
    blobs.append(b">ws\n" + rng.choice(np.frombuffer(b"AC \t\r\n\x0b\x0c>", dtype=np.uint8), size=30000).tobytes())
    return blobs


def unparsable_blobs():
    """Files the bio reader refuses ("Expected > at record start."): anything but '>' as the first byte."""
    return [b"\n\n>lead empty lines\nACGT\n", b"\n\n", b"\n", b" >x\nAC\n", b"\r\n>x\nAC\n", b"ACGT\n>late header\nACGT\n", b"A"]


def n_run_records():
    """N-runs around the 5000 threshold, at fragment borders, across them, whole fragments of N, an empty fragment."""
    rng = np.random.default_rng(8)
    s = lambda n: rand_seq(rng, n)   # noqa: E731
    N = lambda n: "N" * n            # noqa: E731
    return [
        ("short_and_edge", s(300) + N(5000) + s(700) + N(5001) + s(900) + N(12) + s(10)),
        ("leading_short", N(100) + s(2000) + N(6000) + N(1) + s(50) + N(40)),
        ("leading_long", N(7000) + s(3000) + N(9000)),
        ("ends_with_3000", s(500) + N(3000)),
        ("starts_with_3000", N(3000) + s(500)),
        ("ends_with_6000", s(500) + N(6000)),
        ("starts_with_6000", N(6000) + s(500) + "n" * 5500 + s(100)),
        ("all_n_long", N(20000)),
        ("all_n_short", N(30)),
        ("empty", ""),
        ("tail", s(12345)),
    ]
