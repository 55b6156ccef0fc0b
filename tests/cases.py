"""Seeded stress inputs shared by the CPU emulation tests and the GPU parity tests."""
import numpy as np

from tests import kat


def stress_text(seed: int, n: int = 60000, n_dups: int = 12, with_n: bool = True, tandem: bool = True) -> np.ndarray:
    """Random DNA with planted direct / reverse / complement / reverse-complement copies (0-3 % SNPs, a few indels),
    short tandem repeats, a high-copy 200 bp element and N-runs of assorted lengths (also > 5000)."""
    rng = np.random.default_rng(seed)
    t = kat.rand_dna(rng, n)
    for _ in range(n_dups):
        L = int(rng.integers(300, 4000))
        src = int(rng.integers(0, n - L))
        dst = int(rng.integers(0, n - L))
        seg = t[src:src + L].copy()
        d = rng.random() * 0.03
        mut = rng.random(L) < d
        seg[mut] = kat.rand_dna(rng, int(mut.sum()))
        if rng.random() < 0.3 and L > 600:  # one small indel
            p = int(rng.integers(100, L - 100))
            seg = np.concatenate([seg[:p], kat.rand_dna(rng, int(rng.integers(1, 4))), seg[p:]])[:L]
        mode = int(rng.integers(0, 4))
        if mode == 1:
            seg = seg[::-1].copy()
        elif mode == 2:
            seg = kat.complement(seg)
        elif mode == 3:
            seg = kat.revcomp(seg)
        t[dst:dst + L] = seg
    if tandem:
        unit = kat.rand_dna(rng, 37)
        p = int(rng.integers(0, n - 3000))
        t[p:p + 37 * 60] = np.tile(unit, 60)
        elem = kat.rand_dna(rng, 200)
        for _ in range(40):
            q = int(rng.integers(0, n - 200))
            e = elem.copy()
            m = rng.random(200) < 0.05
            e[m] = kat.rand_dna(rng, int(m.sum()))
            t[q:q + 200] = e
    if with_n:
        for L in (7, 50, 700, 5000, 5001, 6500):
            p = int(rng.integers(0, n - L))
            t[p:p + L] = ord("N")
        for _ in range(30):  # scattered single Ns (probes containing N are still searched)
            t[int(rng.integers(0, n))] = ord("N")
    return t


def settings_grid():
    """(label, kwargs for oracle.make_settings / RunSettings)"""
    return [
        ("default", dict(probe_size=20, gap_size=100, min_length=1000, max_cardinality=500)),
        ("RC", dict(probe_size=20, gap_size=100, min_length=1000, max_cardinality=500, reverse=True, complement=True)),
        ("R", dict(probe_size=20, gap_size=100, min_length=500, max_cardinality=500, reverse=True)),
        ("C", dict(probe_size=20, gap_size=100, min_length=500, max_cardinality=500, complement=True)),
        ("k32", dict(probe_size=32, gap_size=200, min_length=400, max_cardinality=500, reverse=True, complement=True)),
        ("k12-lowcard", dict(probe_size=12, gap_size=30, min_length=200, max_cardinality=8)),
        ("k40", dict(probe_size=40, gap_size=10, min_length=300, max_cardinality=50)),
        ("k9-gap0", dict(probe_size=9, gap_size=0, min_length=100, max_cardinality=20)),
    ]
