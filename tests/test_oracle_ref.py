"""Pin the oracle's SA / sa_search / LUT restatements to the reference's real C library (oracle/_ref, compiled in place
from /root/reference/libdivsufsort by oracle/Makefile), and to the committed golden fixtures made from it
(tests/golden/make_golden.py)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle
from tests import kat

GOLD = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref/libdivsufsort64.so not built")


def _texts():
    rng = np.random.default_rng(2026)
    out = {}
    out["random4k"] = np.concatenate([kat.rand_dna(rng, 4000), [ord("$")]]).astype(np.uint8)
    t = kat.rand_dna(rng, 6000); t[1000:1700] = ord("N"); t[3000:3600] = t[200:800]; t[5990:] = ord("N")
    out["nrun_dup6k"] = np.concatenate([t, [ord("$")]]).astype(np.uint8)
    out["polyA"] = np.frombuffer(b"A" * 777 + b"$", dtype=np.uint8).copy()
    out["no_terminator"] = kat.rand_dna(rng, 3000)                 # divsufsort contract without '$'
    out["bytes"] = rng.integers(0, 256, size=5000).astype(np.uint8)  # full byte alphabet
    out["abab"] = np.frombuffer(b"AB" * 500, dtype=np.uint8).copy()
    out["tiny1"] = np.frombuffer(b"$", dtype=np.uint8).copy()
    out["tiny2"] = np.frombuffer(b"A$", dtype=np.uint8).copy()
    return out


@needs_ref
@pytest.mark.parametrize("name", list(_texts().keys()))
def test_sa_matches_divsufsort64(name):
    t = _texts()[name]
    sa_ref = oracle.ref_divsufsort64(t)
    assert oracle.ref_sufcheck64(t, sa_ref) == 0
    sa = oracle.suffix_array(t)
    assert np.array_equal(sa, sa_ref)


@needs_ref
def test_sa_search_matches_reference():
    t = _texts()["nrun_dup6k"]
    sa = oracle.ref_divsufsort64(t)
    R, L = oracle.ref(), oracle.lib()
    rng = np.random.default_rng(5)
    pats = [bytes(t[i:i + l]) for i, l in zip(rng.integers(0, 5900, 300), rng.integers(1, 30, 300))]
    pats += [b"ACGTACGTAC", b"NNNNNNNN", b"TTTTTTTTTTTTTTTT", b"A", b"T", b"$", b"N" * 700, b"N" * 701]
    for p in pats:
        pa = np.frombuffer(p, dtype=np.uint8)
        for (l, r) in [(0, len(sa)), (10, len(sa) - 7), (100, 100), (200, 3000)]:
            i1, i2 = C.c_int64(), C.c_int64()
            c1 = R.sa_searchb64(t.ctypes.data, len(t), pa.ctypes.data, len(pa), sa.ctypes.data, len(sa), C.byref(i1), l, r)
            c2 = L.oracle_sa_searchb(t.ctypes.data, len(t), pa.ctypes.data, len(pa), sa.ctypes.data, len(sa), C.byref(i2), l, r)
            assert c1 == c2, (p, l, r)
            if c1 > 0 or l < r:
                assert i1.value == i2.value, (p, l, r, c1)


@needs_ref
def test_lut_matches_reference_sa_searchb64():
    """Searcher::new (searcher.rs:99-143): every one of the 5^8 keys against the reference's sa_searchb64."""
    t = _texts()["nrun_dup6k"]
    sa = oracle.ref_divsufsort64(t)
    keys, lo, hi = oracle.lut(t, sa)
    R = oracle.ref()
    rng = np.random.default_rng(0)
    nonempty = np.nonzero(hi > lo)[0]
    sample = np.concatenate([nonempty, rng.integers(0, oracle.LUT_SIZE, 3000)])
    for e in sample:
        p = np.frombuffer(int(keys[e]).to_bytes(8, "little"), dtype=np.uint8)
        out = C.c_int64()
        cnt = R.sa_searchb64(t.ctypes.data, len(t), p.ctypes.data, 8, sa.ctypes.data, len(sa), C.byref(out), 0, len(sa))
        assert (out.value, out.value + cnt) == (int(lo[e]), int(hi[e]))
    # invariant of SURVEY §8c: counts over ACGT-only keys of an N-free text sum to n-7
    t2 = _texts()["random4k"]
    sa2 = oracle.ref_divsufsort64(t2)
    _, lo2, hi2 = oracle.lut(t2, sa2)
    assert int((hi2 - lo2).sum()) == (len(t2) - 1) - 7


def test_golden_fixtures():
    """Fixtures generated from the reference's C library (tests/golden/make_golden.py); runs without oracle/_ref."""
    with open(os.path.join(GOLD, "sa_golden.json")) as f:
        gold = json.load(f)
    assert gold["generator"].startswith("tests/golden/make_golden.py")
    for name, entry in gold["cases"].items():
        t = np.frombuffer(bytes.fromhex(entry["text_hex"]), dtype=np.uint8)
        sa = oracle.suffix_array(t)
        assert sa.tolist() == entry["sa"], name
        if "lut_nonempty" in entry:
            keys, lo, hi = oracle.lut(t, sa)
            got = {format(int(k), "016x"): [int(a), int(b)] for k, a, b in zip(keys, lo, hi) if b > a}
            assert got == entry["lut_nonempty"], name
