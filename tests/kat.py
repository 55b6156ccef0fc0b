"""Known-answer cases of SURVEY.md §8c (hand-derived; NOT reference-executed — "parity unpinned" for the Rust rows).

Each case: (name, strand-without-$ as bytes, chunks, settings kwargs, expected families after all post-steps) where a
family is a list of (left, right, left_length, right_length, reversed, complemented).
Texts are seeded numpy draws, so both the oracle tests and the GPU parity tests see identical inputs.
"""
from __future__ import annotations

import numpy as np

_COMP = {ord("A"): ord("T"), ord("T"): ord("A"), ord("C"): ord("G"), ord("G"): ord("C"), ord("N"): ord("N")}


def rand_dna(rng, n) -> np.ndarray:
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)].copy()


def complement(a: np.ndarray) -> np.ndarray:
    lut = np.full(256, ord("N"), dtype=np.uint8)
    for k, v in _COMP.items():
        lut[k] = v
    return lut[a]


def revcomp(a: np.ndarray) -> np.ndarray:
    return complement(a)[::-1].copy()


def _base(seed=1, n=20000, slen=3000):
    rng = np.random.default_rng(seed)
    text = rand_dna(rng, n)
    s = rand_dna(rng, slen)
    return rng, text, s


def cases():
    out = []
    D = dict(probe_size=20, gap_size=100, min_length=1000, max_cardinality=500)

    # KAT-D: S at 5000 and 12000, no flags
    rng, t, s = _base(1)
    t[5000:8000] = s; t[12000:15000] = s
    out.append(("KAT-D", t.copy(), [(0, 20000)], dict(D), [[(5000, 12000, 3000, 3000, False, False)]]))
    # KAT-OFFSET: same text, chunk (1000, 19000)
    out.append(("KAT-OFFSET", t.copy(), [(1000, 19000)], dict(D), [[(5000, 12000, 3000, 3000, False, False)]]))
    # KAT-END: arm still active at chunk end is dropped (Q3) / kept when the chunk runs on far enough
    out.append(("KAT-END-drop", t.copy(), [(0, 8050), (8050, 11950)], dict(D), []))
    out.append(("KAT-END-keep", t.copy(), [(0, 8200), (8200, 11800)], dict(D),
                [[(5000, 12000, 3000, 3000, False, False)]]))
    # direct text under -R / -RC finds nothing (Q7)
    out.append(("KAT-D-under-RC", t.copy(), [(0, 20000)], dict(D, reverse=True, complement=True), []))

    # KAT-RC: S at 5000, revcomp(S) at 13000
    rng, t, s = _base(2)
    t[5000:8000] = s; t[13000:16000] = revcomp(s)
    out.append(("KAT-RC", t.copy(), [(0, 20000)], dict(D, reverse=True, complement=True),
                [[(5000, 13000, 3000, 3000, True, True)]]))
    out.append(("KAT-RC-under-R", t.copy(), [(0, 20000)], dict(D, reverse=True), []))
    out.append(("KAT-RC-noflags", t.copy(), [(0, 20000)], dict(D), []))

    # KAT-R: reverse(S) at 13000, -R
    rng, t, s = _base(3)
    t[5000:8000] = s; t[13000:16000] = s[::-1]
    out.append(("KAT-R", t.copy(), [(0, 20000)], dict(D, reverse=True), [[(5000, 13000, 3000, 3000, True, False)]]))

    # KAT-C: complement(S) at 13000, -C
    rng, t, s = _base(4)
    t[5000:8000] = s; t[13000:16000] = complement(s)
    out.append(("KAT-C", t.copy(), [(0, 20000)], dict(D, complement=True), [[(5000, 13000, 3000, 3000, False, True)]]))

    # KAT-Q1: S at 5000, revcomp(S) at 12000, -RC: needle-local i == global m.start for every probe -> nothing
    rng, t, s = _base(5)
    t[5000:8000] = s; t[12000:15000] = revcomp(s)
    out.append(("KAT-Q1", t.copy(), [(0, 20000)], dict(D, reverse=True, complement=True), []))

    # KAT-UNALIGNED: S at 5003 and 13000
    rng, t, s = _base(6)
    t[5003:8003] = s; t[13000:16000] = s
    out.append(("KAT-UNALIGNED", t.copy(), [(0, 20000)], dict(D), [[(5010, 13007, 2990, 2990, False, False)]]))

    # KAT-SNP: one substitution every 150 bp in the second copy
    rng, t, s = _base(7)
    s2 = s.copy()
    for p in range(75, 3000, 150):
        s2[p] = ord("ACGT"[("ACGT".index(chr(s2[p])) + 1) % 4])
    t[5000:8000] = s; t[12000:15000] = s2
    out.append(("KAT-SNP", t.copy(), [(0, 20000)], dict(D), [[(5000, 12000, 3000, 3000, False, False)]]))

    # KAT-GAP80 / KAT-GAP130: random bp replacing part of the second copy
    rng, t, s = _base(8)
    s2 = s.copy(); s2[1500:1580] = rand_dna(rng, 80)
    t[5000:8000] = s; t[12000:15000] = s2
    out.append(("KAT-GAP80", t.copy(), [(0, 20000)], dict(D), [[(5000, 12000, 3000, 3000, False, False)]]))
    rng, t, s = _base(9)
    s2 = s.copy(); s2[1500:1630] = rand_dna(rng, 130)
    t[5000:8000] = s; t[12000:15000] = s2
    out.append(("KAT-GAP130", t.copy(), [(0, 20000)], dict(D),
                [[(5000, 12000, 1500, 1500, False, False)], [(6630, 13630, 1370, 1370, False, False)]]))

    # KAT-3COPIES: S at 5000, 13000, 21000 in a 30 000 text. The first family holds both pairs that start at 5000;
    # their order inside the family is the SA order of the two matches, undone by the final sort only when
    # `left` differs — here both have left == 5000, so compare as a set (see test).
    rng, t, s = _base(10, n=30000)
    t[5000:8000] = s; t[13000:16000] = s; t[21000:24000] = s
    out.append(("KAT-3COPIES", t.copy(), [(0, 30000)], dict(D),
                [[(5000, 13000, 3000, 3000, False, False), (5000, 21000, 3000, 3000, False, False)],
                 [(13000, 21000, 3000, 3000, False, False)]]))
    return out
