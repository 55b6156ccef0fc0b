"""asgart-slice's duplicon filters — host code, no GPU: asgart_b200_slice_families against a restatement
of RunResult::remove_* / max_family_members (src/structs.rs:143-198) and the --min-length retain of
src/bin/asgart-slice.rs:149-154 applied to the JSON of the unfiltered families."""
import itertools
import json

import numpy as np

import asgart_b200 as ab
from asgart_b200.api import PROTOSD_DTYPE


def _families(rng, n_fam, n):
    off, rows = [0], []
    for _ in range(n_fam):
        for _ in range(int(rng.integers(0, 5))):
            rows.append((int(rng.integers(0, n - 10)), int(rng.integers(0, n - 10)), int(rng.integers(1, 3000)), int(rng.integers(1, 3000)),
                         0.0, int(rng.integers(0, 2)), int(rng.integers(0, 2)), (0, 0)))
        off.append(len(rows))
    return ab.Families(np.array(off, dtype=np.uint64), np.array(rows, dtype=PROTOSD_DTYPE))


def _restated(run, no_direct, no_reversed, no_uncomplemented, no_complemented, no_inter, no_intra, min_length, max_members):
    fams = [list(f) for f in run["families"]]

    def retain(pred):
        nonlocal fams
        fams = [[sd for sd in f if pred(sd)] for f in fams]
        fams = [f for f in fams if f]
    if no_direct:
        retain(lambda sd: sd["reversed"])                      # structs.rs:143-148
    if no_reversed:
        retain(lambda sd: not sd["reversed"])
    if no_uncomplemented:
        retain(lambda sd: sd["complemented"])
    if no_complemented:
        retain(lambda sd: not sd["complemented"])
    if no_inter:
        retain(lambda sd: sd["chr_left"] == sd["chr_right"])   # :171-176
    if no_intra:
        retain(lambda sd: sd["chr_left"] != sd["chr_right"])   # :189-194
    if min_length is not None:
        retain(lambda sd: min(sd["left_length"], sd["right_length"]) >= min_length)   # asgart-slice.rs:149-154
    if max_members is not None:
        fams = [f for f in fams if len(f) <= max_members]      # structs.rs:196-198
    return fams


def test_slice_filters_match_the_reference_semantics():
    rng = np.random.default_rng(4)
    n = 50_000
    strand = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    # two fragments share a name: the reference compares names, not fragments
    prep = ab.Prepared.from_memory(strand, [("chrA", 0, 20_000), ("chrB", 20_000, 10_000), ("chrA", 30_000, 20_000)], "x.fa")
    st = ab.RunSettings()
    fam = _families(rng, 40, n)
    run = json.loads(prep.to_json(st, fam))
    assert any(not f for f in fam.as_lists()) and len(run["families"]) == 40
    flags = list(itertools.product([False, True], repeat=6))
    for k, (nd, nr, nu, nc, ni, na) in enumerate(flags):
        ml = [None, 0, 500, 1500][k % 4]
        mm = [None, 0, 1, 3][(k // 4) % 4]
        got = prep.slice(fam, no_direct=nd, no_reversed=nr, no_uncomplemented=nu, no_complemented=nc, no_inter=ni, no_intra=na,
                         min_length=ml, max_family_members=mm)
        want = _restated(run, nd, nr, nu, nc, ni, na, ml, mm)
        assert json.loads(prep.to_json(st, got))["families"] == want, (nd, nr, nu, nc, ni, na, ml, mm)
    assert prep.slice(fam).as_lists() == fam.as_lists()       # no filter: empty families stay, as in the reference
