"""Small run of the whole hot path for compute-sanitizer (memcheck / racecheck / synccheck), one GPU:
    compute-sanitizer --tool memcheck  python tools/sanitize.py
smoke() (SA build + LUT + search + automaton + post-steps on a 200 kbp two-fragment strand, checked against the oracle),
the u64 index path, --trim, ComputeScore, the GPU FASTA ingest and a 2-member sharded build on the same device (host
threads as members: the peer stores / loads and the staged exchanges of the sharded build run for real)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
import asgart_b200 as ab  # noqa: E402
from tests import cases  # noqa: E402

g.smoke()
text = cases.stress_text(5, n=120_000, n_dups=20)
prep = ab.Prepared.from_memory(text, [("a", 0, 50_000), ("b", 50_000, len(text) - 50_000)], "san.fa")
strand = np.array(prep.strand)
with ab.Context(0) as ctx:
    ctx.load_strand(strand)
    ctx.build_index()
    ref = ctx.download_sa()
    fam = ctx.search(prep.chunks, ab.RunSettings(min_duplication_length=300, reverse=True, complement=True, compute_score=True))
    ctx.set_index_bits(64)
    ctx.build_index()
    assert np.array_equal(ctx.download_sa(), ref)
    assert ctx.check_sa() == 0
    ctx.set_index_bits(0)
    ctx.build_index(trim=(1000, 90_000))
    ctx.search(prep.chunks, ab.RunSettings(min_duplication_length=300, trim=(1000, 90_000)))
    fa = b">a x\n" + bytes(text[:50_000]) + b"\n>b\n" + bytes(text[50_000:]) + b"\n"
    got = ctx.ingest([fa], False, names=["san.fa"])
    assert got.chunks == prep.chunks
ctxs = [ab.Context(0) for _ in range(2)]
for c in ctxs:
    c.load_strand(strand)
ab.build_index_group(ctxs)
for c in ctxs:
    assert np.array_equal(c.download_sa(), ref)
    c.close()
print("sanitize run ok:", fam.n_families, "families")
