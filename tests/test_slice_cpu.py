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


# ---------------------------------------------------------------------------------------------------------------------
# The rest of asgart-slice (--collapse, --no-inter-relaxed, --keep/--restrict/--exclude-fragments [--regexp]) against a
# restatement of src/structs.rs:178-187, 204-416 and src/bin/asgart-slice.rs:126-191 on the JSON dict.
COLLAPSED = "ASGART_COLLAPSED"          # src/structs.rs:9


def _rs_flatten(run):                   # RunResult::flatten, src/structs.rs:350-416
    m = run["strand"]["map"]
    if len(m) < 2:
        return
    n = float(len(m))
    lengths = [float(c["length"]) for c in m]
    avg = sum(lengths) / n
    std = (1.0 / (n - 1.0) * sum((x - avg) ** 2.0 for x in lengths)) ** 0.5
    to_flatten = [dict(c) for c in m if float(c["length"]) <= avg + std and len(c["name"].encode()) > 2]
    to_flatten_len = sum(c["length"] for c in to_flatten)
    to_keep = [dict(c) for c in m if not any(c["name"] == r["name"] for r in to_flatten)]
    to_keep_len = sum(c["length"] for c in to_keep)
    i = 0
    for c in to_keep:
        c["position"] = i
        i += c["length"]
    for c in to_flatten:
        c["position"] = i
        i += c["length"]
    pos = {c["name"]: c["position"] for c in to_flatten}
    run["strand"]["map"] = to_keep + [{"name": COLLAPSED, "position": to_keep_len + 1, "length": to_flatten_len}]
    for f in run["families"]:
        for sd in f:
            lm = any(c["name"] == sd["chr_left"] for c in to_flatten)
            rm = any(c["name"] == sd["chr_right"] for c in to_flatten)
            if lm:
                sd["chr_left_position"] += pos[sd["chr_left"]]
                sd["chr_left"] = COLLAPSED
            if rm:
                sd["chr_right_position"] += pos[sd["chr_right"]]
                sd["chr_right"] = COLLAPSED


def _rs_find_chr(run, name):            # StrandResult::find_chr, :77-79
    for c in run["strand"]["map"]:
        if c["name"] == name:
            return c
    return None


def _rs_consolidate(run, keep):         # consolidate_families, :204-230 (keep: predicate on fragment names)
    run["families"] = [f for f in run["families"] if f]
    run["strand"]["map"] = [c for c in run["strand"]["map"] if keep(c["name"])]
    run["strand"]["length"] = sum(c["length"] for c in run["strand"]["map"])
    i = 0
    for c in run["strand"]["map"]:
        c["position"] = i
        i += c["length"]
    for f in run["families"]:
        for sd in f:
            l, r = _rs_find_chr(run, sd["chr_left"]), _rs_find_chr(run, sd["chr_right"])
            sd["global_left_position"] = l["position"] + sd["chr_left_position"] if l else 0
            sd["global_right_position"] = r["position"] + sd["chr_right_position"] if r else 0


class _Panic(Exception):
    pass


def _rs_exclude(run, excluded):         # exclude_fragments(_regexp), :277-348: same layout code, but find_chr(..).unwrap()
    run["families"] = [[sd for sd in f if not excluded(sd["chr_left"]) and not excluded(sd["chr_right"])] for f in run["families"]]
    run["families"] = [f for f in run["families"] if f]
    run["strand"]["map"] = [c for c in run["strand"]["map"] if not excluded(c["name"])]
    run["strand"]["length"] = sum(c["length"] for c in run["strand"]["map"])
    i = 0
    for c in run["strand"]["map"]:
        c["position"] = i
        i += c["length"]
    for f in run["families"]:
        for sd in f:
            l, r = _rs_find_chr(run, sd["chr_left"]), _rs_find_chr(run, sd["chr_right"])
            if l is None or r is None:
                raise _Panic()
            sd["global_left_position"] = l["position"] + sd["chr_left_position"]
            sd["global_right_position"] = r["position"] + sd["chr_right_position"]


def _rs_slice(run, collapse=False, no_inter=False, no_inter_relaxed=False, no_intra=False, min_length=None, max_family_members=None,
              keep=None, restrict=None, exclude=None, regexp=False):
    import re

    def retain(pred):
        run["families"] = [[sd for sd in f if pred(sd)] for f in run["families"]]
        run["families"] = [f for f in run["families"] if f]
    if collapse:
        _rs_flatten(run)
    if no_inter:
        retain(lambda sd: sd["chr_left"] == sd["chr_right"])
    if no_inter_relaxed:                # :178-187
        retain(lambda sd: sd["chr_left"] == sd["chr_right"] or sd["chr_left"] == COLLAPSED or sd["chr_right"] == COLLAPSED)
    if no_intra:
        retain(lambda sd: sd["chr_left"] != sd["chr_right"])
    if min_length is not None:
        retain(lambda sd: min(sd["left_length"], sd["right_length"]) >= min_length)
    if max_family_members is not None:
        run["families"] = [f for f in run["families"] if len(f) <= max_family_members]

    def select(items, mode):            # asgart-slice.rs:159-191; *_fragments / *_fragments_regexp, structs.rs:232-348
        if items is None:
            return
        matchers = [(lambda n, rx=re.compile(p): rx.search(n) is not None) for p in items] if regexp else [lambda n: n in items]
        for m in matchers:
            if mode == 2:
                _rs_exclude(run, m)
                continue
            if mode == 0:
                run["families"] = [[sd for sd in f if m(sd["chr_left"]) or m(sd["chr_right"])] for f in run["families"]]
            else:
                run["families"] = [[sd for sd in f if m(sd["chr_left"]) and m(sd["chr_right"])] for f in run["families"]]
            _rs_consolidate(run, m)
    select(keep, 0)
    select(restrict, 1)
    select(exclude, 2)
    return run


def test_collapse_and_fragment_selection_match_the_reference_semantics():
    rng = np.random.default_rng(8)
    frags = [("chr1", 0, 30_000), ("chr2", 30_000, 24_000), ("scaffold_17", 54_000, 900), ("scaffold_18", 54_900, 1_300),
             ("un", 56_200, 700), ("chrUn_KI270", 56_900, 2_100), ("chrX", 59_000, 10_000)]      # positions from 69 000 on: "unknown"
    n = 70_000
    strand = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    prep = ab.Prepared.from_memory(strand, frags, "x.fa")
    st = ab.RunSettings(reverse=True)
    fam = _families(rng, 60, n - 20)      # some duplicons reach past the last fragment: "unknown"
    base = prep.to_json(st, fam)
    cases = [
        dict(collapse=True),
        dict(collapse=True, no_inter_relaxed=True),
        dict(no_inter_relaxed=True),
        dict(collapse=True, no_intra=True, min_length=300),
        dict(keep=["chr1"]),
        dict(keep=["chr1", "chrX", "nope"]),
        dict(restrict=["chr1", "chr2"]),
        dict(restrict=["chr2"], max_family_members=2),
        dict(exclude=["scaffold_17", "un", "unknown"]),
        dict(collapse=True, keep=[COLLAPSED]),
        dict(collapse=True, restrict=[COLLAPSED, "chr1"], no_inter_relaxed=True),
        dict(keep=["^chr[0-9]+$"], regexp=True),
        dict(keep=["^chr", "X$"], regexp=True),
        dict(restrict=["chr(1|2)"], regexp=True),
        dict(exclude=["scaffold", "^un"], regexp=True),
        dict(keep=["chr1"], restrict=["chr1"], exclude=["chr2"]),
    ]
    n_checked = 0
    for kw in cases:
        want_run = json.loads(base)
        try:
            _rs_slice(want_run, **kw)
            want = want_run
        except _Panic:
            want = None
        api_kw = dict(kw)
        for a, b in (("keep", "keep_fragments"), ("restrict", "restrict_fragments"), ("exclude", "exclude_fragments")):
            if a in api_kw:
                api_kw[b] = api_kw.pop(a)
        if want is None:
            import pytest
            with pytest.raises(ab.AsgartB200Error, match="panics"):
                prep.slice_json(st, fam, **api_kw)
        else:
            got = json.loads(prep.slice_json(st, fam, **api_kw))
            assert got == want, kw
            n_checked += 1
    assert n_checked >= 14
    # a duplicon on "unknown" makes --exclude-fragments panic in the reference unless "unknown" itself is excluded
    assert any(sd["chr_left"] == "unknown" or sd["chr_right"] == "unknown" for f in json.loads(base)["families"] for sd in f)
    import pytest
    with pytest.raises(ab.AsgartB200Error, match="panics"):
        prep.slice_json(st, fam, exclude_fragments=["chr2"])
    with pytest.raises(ab.AsgartB200Error, match="compiling"):
        prep.slice_json(st, fam, keep_fragments=["chr(1"], regexp=True)
    # nothing asked: the JSON of the run itself
    assert prep.slice_json(st, fam) == base


def test_random_option_combinations_match_the_restatement():
    """Random fragment maps (names with dots, blanks, shared prefixes; sometimes a tail of positions outside every fragment)
    and random combinations of all the asgart-slice options, in asgart-slice's order, panics included. (An offline run of
    1800 combinations found no difference; 240 stay in the suite.)"""
    rng = np.random.default_rng(1)
    names_pool = ["chr1", "chr2", "chr10", "scaffold_17", "scaffold_18", "un", "chrUn_KI270", "chrX", "a.b", "x y", "chr1_alt"]
    patterns = ["^chr", "chr[0-9]+$", "scaffold", "^un", "X$", "chr(1|2)", ".", "nope", "_", "^" + COLLAPSED + "$", "unknown"]
    checked = panics = 0
    for _ in range(40):
        nf = int(rng.integers(1, 8))
        names = list(rng.choice(names_pool, size=nf, replace=False))
        lens = [int(rng.integers(200, 20000)) for _ in range(nf)]
        pos = np.concatenate([[0], np.cumsum(lens)])
        frags = [(names[i], int(pos[i]), lens[i]) for i in range(nf)]
        n = int(pos[-1]) + int(rng.integers(0, 3000))
        strand = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
        prep = ab.Prepared.from_memory(strand, frags, "x.fa")
        st = ab.RunSettings(reverse=bool(rng.integers(0, 2)), complement=bool(rng.integers(0, 2)))
        fam = _families(rng, int(rng.integers(1, 40)), n - 20)
        base = prep.to_json(st, fam)
        for _ in range(6):
            kw = {}
            for name, p_on, values in (("collapse", 0.4, [True]), ("no_inter", 0.2, [True]), ("no_inter_relaxed", 0.3, [True]), ("no_intra", 0.2, [True]),
                                       ("min_length", 0.3, [1, 100, 300, 1000]), ("max_family_members", 0.3, [1, 2, 3, 10])):
                if rng.random() < p_on:
                    kw[name] = values[int(rng.integers(0, len(values)))]
            regexp = bool(rng.random() < 0.4)
            pool = patterns if regexp else names_pool + [COLLAPSED, "unknown", "nope"]
            for name in ("keep", "restrict", "exclude"):
                if rng.random() < 0.4:
                    kw[name] = [str(x) for x in rng.choice(pool, size=int(rng.integers(1, 4)), replace=False)]
            if regexp:
                kw["regexp"] = True
            want = json.loads(base)
            try:
                _rs_slice(want, **kw)
            except _Panic:
                want = None
            api_kw = {{"keep": "keep_fragments", "restrict": "restrict_fragments", "exclude": "exclude_fragments"}.get(k, k): v for k, v in kw.items()}
            checked += 1
            if want is None:
                panics += 1
                import pytest
                with pytest.raises(ab.AsgartB200Error, match="panics"):
                    prep.slice_json(st, fam, **api_kw)
            else:
                assert json.loads(prep.slice_json(st, fam, **api_kw)) == want, (kw, frags)
    assert checked == 240 and panics > 5
