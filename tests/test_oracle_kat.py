"""The CPU oracle against the hand-derived known-answer cases and its own invariants."""
import numpy as np
import pytest

import oracle
from tests import kat


@pytest.mark.parametrize("case", kat.cases(), ids=lambda c: c[0])
def test_kat(case):
    name, text, chunks, kw, expected = case
    strand = np.concatenate([text, np.frombuffer(b"$", dtype=np.uint8)])
    sa = oracle.best_suffix_array(strand)
    assert sa[0] == len(strand) - 1
    st = oracle.make_settings(**kw)
    out = oracle.search(strand, sa, chunks, st, oracle.POST_ALL, threads=2)
    got = out.families.as_lists()
    if name == "KAT-3COPIES":
        assert [sorted(f) for f in got] == [sorted(f) for f in expected]
    else:
        assert got == expected


def test_counters_and_alg_bytes():
    name, text, chunks, kw, expected = kat.cases()[0]
    strand = np.concatenate([text, np.frombuffer(b"$", dtype=np.uint8)])
    sa = oracle.best_suffix_array(strand)
    out = oracle.search(strand, sa, chunks, oracle.make_settings(**kw), 0, threads=1)
    c = out.counters
    # probes per chunk = ceil((l - k - s) / s)   (SURVEY §8 notation)
    assert c["probes"] == -(-(20000 - 20 - 10) // 10)
    assert c["skipped_n"] == 0 and c["skipped_card"] == 0
    assert c["searched"] == c["probes"]
    assert c["matches"] >= 299          # one match per aligned probe of the planted copy
    assert c["alg_bytes"] >= 24 * c["searched"]


def test_chunker_rules():
    # N-runs <= 5000 stay inside a chunk; > 5000 split; leading short run belongs to the chunk (asgart.rs:317-366)
    rng = np.random.default_rng(0)
    t = kat.rand_dna(rng, 40000)
    t[0:100] = ord("N")            # short leading run
    t[10000:15001] = ord("N")      # 5001 > threshold -> split
    t[20000:25000] = ord("N")      # exactly 5000 -> kept inside
    p = oracle.Prepared.from_memory(t, [("f", 0, 40000)])
    assert p.chunks == [(0, 10000), (15001, 24999)]
    # all-N fragment -> one chunk covering it (asgart.rs:361-363)
    t2 = np.full(6000, ord("N"), dtype=np.uint8)
    p2 = oracle.Prepared.from_memory(t2, [("f", 0, 6000)])
    assert p2.chunks == [(0, 6000)]
    # chunks never span fragments
    t3 = kat.rand_dna(rng, 3000)
    p3 = oracle.Prepared.from_memory(t3, [("a", 0, 1000), ("b", 1000, 2000)])
    assert p3.chunks == [(0, 1000), (1000, 2000)]
    assert p3.strand[-1] == ord("$")


def test_fasta_and_json(tmp_path):
    rng = np.random.default_rng(11)
    t = kat.rand_dna(rng, 20000)
    s = kat.rand_dna(rng, 3000)
    t[5000:8000] = s
    t[12000:15000] = s
    seq = t.tobytes().decode()
    fa = tmp_path / "toy genome.fa"
    with open(fa, "w") as f:
        f.write(">chrT some description\r\n")
        for i in range(0, 12000, 60):
            f.write(seq[i:i + 60].lower() + "\r\n")   # lower-case: upper-cased when -S is off (asgart.rs:291-293)
        f.write(">chrU\n")
        for i in range(12000, 20000, 70):
            f.write(seq[i:i + 70] + "\n")
    st = oracle.make_settings()
    js = oracle.run_files([str(fa)], st)
    import json
    d = json.loads(js)
    assert js.endswith("}\n")
    assert list(d.keys()) == ["strand", "settings", "families"]
    assert d["strand"] == {"name": str(fa), "length": 20000,
                           "map": [{"name": "chrT", "position": 0, "length": 12000},
                                   {"name": "chrU", "position": 12000, "length": 8000}]}
    assert d["settings"] == {"probe_size": 20, "max_gap_size": 120, "min_duplication_length": 1000,
                             "max_cardinality": 500, "trim": None, "skip_masked": False}
    assert d["families"] == [[{
        "chr_left": "chrT", "chr_right": "chrU", "global_left_position": 5000, "global_right_position": 12000,
        "chr_left_position": 5000, "chr_right_position": 0, "left_length": 3000, "right_length": 3000,
        "left_seq": None, "right_seq": None, "identity": 0.0, "reversed": False, "complemented": False}]]
    assert '"identity": 0.0,' in js and '      {\n        "chr_left": "chrT",' in js
    # -S: every lower-case base becomes N -> first fragment is one big N run -> no duplication survives
    st_s = oracle.make_settings(skip_masked=True)
    d2 = json.loads(oracle.run_files([str(fa)], st_s))
    assert d2["families"] == [] and d2["settings"]["skip_masked"] is True


def test_post_steps_unit():
    # reduce_overlap with the bug-compatible merge (asgart.rs:497-513, Q5) and ReOrder (Q4)
    off = np.array([0, 3], dtype=np.int64)
    fields = np.array([[100, 5000, 1000, 1200], [600, 5500, 1000, 900], [150, 5050, 100, 100]], dtype=np.uint64)
    fam = oracle.Families(off, fields, np.zeros(3, np.float32), np.zeros((3, 2), np.uint8))
    t = np.frombuffer(b"A" * 8000 + b"$", dtype=np.uint8)
    red = oracle.post_steps(fam, t, oracle.POST_REDUCE_OVERLAP)
    # x=(600,5500,1000,900) overlaps y=(100,5000,1000,1200) on both arms:
    #   left=min=100, lsize=max(600+1000, 100+1200)-100=1500 ; right=5000, rsize=max(5500+1000, 5000+1200)-5000=1500
    # then (150,5050,100,100) is a sub-segment of the merged one -> dropped
    assert red.as_lists() == [[(100, 5000, 1500, 1500, False, False)]]
    fam2 = oracle.Families(np.array([0, 1], np.int64), np.array([[900, 100, 10, 20]], np.uint64),
                           np.zeros(1, np.float32), np.zeros((1, 2), np.uint8))
    assert oracle.post_steps(fam2, t, oracle.POST_REORDER).as_lists() == [[(100, 900, 10, 20, False, False)]]


def test_filter_ns_f32_threshold():
    # n_content counts N over len+1 bytes and divides by len, in f32, keeps <= 0.2 (structs.rs:454-467)
    t = np.full(3001, ord("A"), dtype=np.uint8)
    t[-1] = ord("$")
    t[0:200] = ord("N")           # arm [0, 1000]: 200 N / 1000 = 0.2 -> kept
    t[1200:1401] = ord("N")       # arm [1100, 2100]: 201 / 1000 -> dropped
    mk = lambda l, r: oracle.Families(np.array([0, 1], np.int64), np.array([[l, r, 1000, 1000]], np.uint64),
                                      np.zeros(1, np.float32), np.zeros((1, 2), np.uint8))
    assert len(oracle.post_steps(mk(0, 2000), t, oracle.POST_FILTER_NS).as_lists()) == 1
    assert oracle.post_steps(mk(0, 1100), t, oracle.POST_FILTER_NS).as_lists() == []
    # the inclusive end: byte at p+len counts
    t2 = np.full(2002, ord("A"), dtype=np.uint8); t2[-1] = ord("$")
    t2[0:200] = ord("N"); t2[1000] = ord("N")
    assert oracle.post_steps(mk(0, 1001), t2, oracle.POST_FILTER_NS).as_lists() == []
