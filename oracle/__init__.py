"""ctypes loader for the CPU oracle — TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may import
this module; the product package ``asgart_b200`` never does. See oracle/asgart_oracle.cpp for what it restates
(reference file:line per function) and what pins it.

Two libraries:
  * ``oracle/_build/liboracle.so``  — our restatement (built by ``make -C oracle``)
  * ``oracle/_ref/libdivsufsort64.so`` — the reference's real C library, compiled in place from
    /root/reference/libdivsufsort (git-ignored, travels to the GPU box prebuilt). Optional.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libdivsufsort64.so")


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is mounted)."""
    src = os.path.join(_HERE, "asgart_oracle.cpp")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale or (os.path.isdir("/root/reference/libdivsufsort/lib") and not os.path.exists(_REF_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)


class Settings(C.Structure):
    _fields_ = [
        ("probe_size", C.c_uint64),
        ("max_gap_size", C.c_uint32),
        ("reverse", C.c_uint32),
        ("complement", C.c_uint32),
        ("skip_masked", C.c_uint32),
        ("min_duplication_length", C.c_uint64),
        ("max_cardinality", C.c_uint64),
        ("has_trim", C.c_uint32),
        ("compute_score", C.c_uint32),
        ("trim_a", C.c_uint64),
        ("trim_b", C.c_uint64),
    ]


def make_settings(probe_size=20, gap_size=100, min_length=1000, max_cardinality=500, reverse=False,
                  complement=False, skip_masked=False, compute_score=False, trim=None) -> Settings:
    """RunSettings as bin/asgart.rs:679-692 builds it: max_gap_size = gap_size + probe_size; trim = the raw CLI pair."""
    t = trim or (0, 0)
    return Settings(probe_size, gap_size + probe_size, int(reverse), int(complement), int(skip_masked), min_length,
                    max_cardinality, int(trim is not None), int(compute_score), int(t[0]), int(t[1]))


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, i64p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_int64), C.POINTER(C.c_uint64)
        L.oracle_suffix_array.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.oracle_suffix_array.restype = C.c_int
        L.oracle_lut.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_sa_searchb.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                        i64p, C.c_int64, C.c_int64]
        L.oracle_sa_searchb.restype = C.c_int64
        L.oracle_searcher_new.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.oracle_searcher_new.restype = C.c_void_p
        L.oracle_searcher_free.argtypes = [C.c_void_p]
        L.oracle_searcher_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                             C.c_void_p, C.c_int64]
        L.oracle_searcher_search.restype = C.c_int64
        L.oracle_search.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(Settings),
                                    C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_search.restype = C.c_void_p
        L.oracle_search_trim.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                         C.POINTER(Settings), C.c_int, C.c_int]
        L.oracle_search_trim.restype = C.c_void_p
        L.oracle_sa_search_literal.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        L.oracle_sa_search_literal.restype = C.c_int64
        L.oracle_effective_trim.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.oracle_effective_trim.restype = C.c_int
        L.oracle_result_from_arrays.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_result_from_arrays.restype = C.c_void_p
        L.oracle_result_post.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.oracle_result_post.restype = C.c_int
        L.oracle_levenshtein.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        L.oracle_levenshtein.restype = C.c_uint32
        L.oracle_result_n_families.argtypes = [C.c_void_p]
        L.oracle_result_n_families.restype = C.c_int64
        L.oracle_result_n_sds.argtypes = [C.c_void_p]
        L.oracle_result_n_sds.restype = C.c_int64
        L.oracle_result_copy.argtypes = [C.c_void_p] * 5
        L.oracle_result_free.argtypes = [C.c_void_p]
        L.oracle_prepare.argtypes = [C.c_char_p, C.c_int]
        L.oracle_prepare.restype = C.c_void_p
        L.oracle_prepared_error.argtypes = [C.c_void_p]
        L.oracle_prepared_error.restype = C.c_char_p
        L.oracle_prepared_strand_len.argtypes = [C.c_void_p]
        L.oracle_prepared_strand_len.restype = C.c_int64
        L.oracle_prepared_strand.argtypes = [C.c_void_p]
        L.oracle_prepared_strand.restype = C.c_void_p
        L.oracle_prepared_n_chunks.argtypes = [C.c_void_p]
        L.oracle_prepared_n_chunks.restype = C.c_int64
        L.oracle_prepared_chunks.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_prepared_n_fragments.argtypes = [C.c_void_p]
        L.oracle_prepared_n_fragments.restype = C.c_int64
        L.oracle_prepared_fragment.argtypes = [C.c_void_p, C.c_int64, u64p, u64p]
        L.oracle_prepared_fragment.restype = C.c_char_p
        L.oracle_prepare_from_memory.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_char_p, C.c_void_p,
                                                 C.c_void_p, C.c_int64]
        L.oracle_prepare_from_memory.restype = C.c_void_p
        L.oracle_prepared_free.argtypes = [C.c_void_p]
        L.oracle_to_json.argtypes = [C.c_void_p, C.POINTER(Settings), C.c_void_p]
        L.oracle_to_json.restype = C.c_void_p
        L.oracle_free_string.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def ref() -> Optional[C.CDLL]:
    """The reference's own libdivsufsort64 (None when oracle/_ref was never built)."""
    global _ref
    if _ref is None:
        build()
        if not os.path.exists(_REF_PATH):
            return None
        R = C.CDLL(_REF_PATH)
        R.divsufsort64.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        R.divsufsort64.restype = C.c_int32
        R.sufcheck64.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]
        R.sufcheck64.restype = C.c_int32
        R.sa_search64.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                  C.POINTER(C.c_int64)]
        R.sa_search64.restype = C.c_int64
        R.sa_searchb64.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                   C.POINTER(C.c_int64), C.c_int64, C.c_int64]
        R.sa_searchb64.restype = C.c_int64
        _ref = R
    return _ref


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def as_strand(text) -> np.ndarray:
    """bytes/str/ndarray -> contiguous uint8 array."""
    if isinstance(text, str):
        text = text.encode()
    if isinstance(text, (bytes, bytearray)):
        return np.frombuffer(bytes(text), dtype=np.uint8).copy()
    return np.ascontiguousarray(text, dtype=np.uint8)


# ------------------------------------------------------------------ suffix array
def suffix_array(text) -> np.ndarray:
    """Oracle's own restated SA (prefix doubling, slow): int64[n]."""
    t = as_strand(text)
    sa = np.empty(len(t), dtype=np.int64)
    rc = lib().oracle_suffix_array(_ptr(t), _ptr(sa), len(t))
    assert rc == 0
    return sa


def ref_divsufsort64(text) -> np.ndarray:
    """The reference's divsufsort64 (libdivsufsort/lib/divsufsort.c:331)."""
    R = ref()
    assert R is not None, "oracle/_ref/libdivsufsort64.so missing (run make -C oracle where /root/reference exists)"
    t = as_strand(text)
    sa = np.empty(len(t), dtype=np.int64)
    rc = R.divsufsort64(_ptr(t), _ptr(sa), len(t))
    assert rc == 0, rc
    return sa


def ref_sufcheck64(text, sa: np.ndarray) -> int:
    R = ref()
    assert R is not None
    t = as_strand(text)
    sa = np.ascontiguousarray(sa, dtype=np.int64)
    return int(R.sufcheck64(_ptr(t), _ptr(sa), len(t), 0))


def best_suffix_array(text) -> np.ndarray:
    """divsufsort64 from oracle/_ref when present (fast), else the restated one."""
    return ref_divsufsort64(text) if ref() is not None else suffix_array(text)


# ------------------------------------------------------------------ LUT / search
LUT_SIZE = 5 ** 8


def lut(text, sa: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Searcher::new restated: (keys uint64[5^8] = LE u64 of each 8-mer, lo int64, hi int64)."""
    t = as_strand(text)
    sa = np.ascontiguousarray(sa, dtype=np.int64)
    keys = np.empty(LUT_SIZE, dtype=np.uint64)
    lo = np.empty(LUT_SIZE, dtype=np.int64)
    hi = np.empty(LUT_SIZE, dtype=np.int64)
    lib().oracle_lut(_ptr(t), len(t), _ptr(sa), _ptr(keys), _ptr(lo), _ptr(hi))
    return keys, lo, hi


class OracleSearcher:
    """searcher.rs Searcher restated (LUT + equal_range_by)."""

    def __init__(self, text, sa):
        self.t = as_strand(text)
        self.sa = np.ascontiguousarray(sa, dtype=np.int64)
        self.h = lib().oracle_searcher_new(_ptr(self.t), len(self.t), _ptr(self.sa))

    def search(self, pattern: bytes) -> np.ndarray:
        p = as_strand(pattern)
        cap = 1 << 16
        while True:
            out = np.empty(cap, dtype=np.int64)
            n = lib().oracle_searcher_search(self.h, _ptr(self.t), len(self.t), _ptr(self.sa), _ptr(p), len(p),
                                             _ptr(out), cap)
            if n <= cap:
                return out[:n].copy()
            cap = int(n)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_searcher_free(self.h)
            self.h = None


# ------------------------------------------------------------------ families
@dataclass
class Families:
    """CSR view of Vec<ProtoSDsFamily>: fam_offsets int64[n_fam+1]; fields uint64[n_sd,4] =
    (left, right, left_length, right_length); identity float32[n_sd]; flags uint8[n_sd,2] = (reversed, complemented)."""
    fam_offsets: np.ndarray
    fields: np.ndarray
    identity: np.ndarray
    flags: np.ndarray

    def as_lists(self) -> List[List[Tuple[int, int, int, int, bool, bool]]]:
        out = []
        for f in range(len(self.fam_offsets) - 1):
            fam = []
            for j in range(int(self.fam_offsets[f]), int(self.fam_offsets[f + 1])):
                l, r, ll, rl = (int(x) for x in self.fields[j])
                fam.append((l, r, ll, rl, bool(self.flags[j, 0]), bool(self.flags[j, 1])))
            out.append(fam)
        return out

    def canonical(self):
        """SURVEY §4 canonical form: SDs sorted inside each family, families sorted."""
        return sorted(sorted(f) for f in self.as_lists())


def _copy_result(h) -> Families:
    L = lib()
    nf, ns = L.oracle_result_n_families(h), L.oracle_result_n_sds(h)
    off = np.zeros(nf + 1, dtype=np.int64)
    fields = np.zeros((ns, 4), dtype=np.uint64)
    ident = np.zeros(ns, dtype=np.float32)
    flags = np.zeros((ns, 2), dtype=np.uint8)
    L.oracle_result_copy(h, _ptr(off), _ptr(fields), _ptr(ident), _ptr(flags))
    return Families(off, fields, ident, flags)


POST_FILTER_NS, POST_REORDER, POST_REDUCE_OVERLAP, POST_SORT = 1, 2, 4, 8
POST_ALL = 15
POST_COMPUTE_SCORE = 16   # --compute-score (bin/asgart.rs:98-111, 744-746); not part of the default pipeline


class RefPanic(RuntimeError):
    pass


def levenshtein(a: bytes, b: bytes) -> int:
    """bio::alignment::distance::levenshtein restated (plain DP)."""
    aa = np.frombuffer(bytes(a), dtype=np.uint8) if len(a) else np.zeros(1, dtype=np.uint8)
    bb = np.frombuffer(bytes(b), dtype=np.uint8) if len(b) else np.zeros(1, dtype=np.uint8)
    return int(lib().oracle_levenshtein(_ptr(aa), len(a), _ptr(bb), len(b)))


@dataclass
class SearchOutput:
    families: Families
    seconds: dict
    counters: dict


def search(text_with_dollar, sa, chunks: Sequence[Tuple[int, int]], settings: Settings, post_mask: int = POST_ALL,
           threads: int = 1) -> SearchOutput:
    """SearchDuplications::run given SA (bin/asgart.rs:137-258) + selected post-steps."""
    t = as_strand(text_with_dollar)
    sa = np.ascontiguousarray(sa, dtype=np.int64)
    ch = np.ascontiguousarray(np.array(chunks, dtype=np.uint64).reshape(-1, 2))
    secs = np.zeros(3, dtype=np.float64)
    ctr = np.zeros(6, dtype=np.uint64)
    h = lib().oracle_search(_ptr(t), len(t), _ptr(sa), _ptr(ch), len(ch), C.byref(settings), post_mask, threads,
                            _ptr(secs), _ptr(ctr))
    if not h:
        raise RefPanic("the reference panics on this input (FilterNs / ComputeScore slice or complement panic)")
    try:
        fam = _copy_result(h)
    finally:
        lib().oracle_result_free(h)
    return SearchOutput(
        fam,
        {"lut": float(secs[0]), "search": float(secs[1]), "post": float(secs[2])},
        dict(zip(["probes", "searched", "skipped_n", "skipped_card", "matches", "alg_bytes"], (int(x) for x in ctr))),
    )


def effective_trim(trim: Tuple[int, int], strand_len_with_dollar: int) -> Optional[Tuple[int, int]]:
    """prepare_data's validation of --trim (bin/asgart.rs:432-463): None when the reference skips trimming."""
    out = np.zeros(2, dtype=np.uint64)
    ok = lib().oracle_effective_trim(int(trim[0]), int(trim[1]), int(strand_len_with_dollar), _ptr(out))
    return (int(out[0]), int(out[1])) if ok else None


def trimmed_suffix_array(text_with_dollar, trim: Tuple[int, int]) -> np.ndarray:
    """bin/asgart.rs:142-147: suffix array of strand[a..b] + '$', every entry shifted by a."""
    t = as_strand(text_with_dollar)
    a, b = trim
    sub = np.concatenate([t[a:b], np.frombuffer(b"$", dtype=np.uint8)])
    return best_suffix_array(sub) + a


def sa_search_literal(text, pattern: bytes, sa: np.ndarray) -> Tuple[int, int]:
    """sa_search (libdivsufsort/lib/utils.c:282-349) step for step: (first index, count)."""
    t = as_strand(text)
    p = as_strand(pattern)
    sa = np.ascontiguousarray(sa, dtype=np.int64)
    idx = C.c_int64()
    n = lib().oracle_sa_search_literal(_ptr(t), len(t), _ptr(p), len(p), _ptr(sa), len(sa), C.byref(idx))
    return idx.value, n


def search_trim(text_with_dollar, trim: Tuple[int, int], chunks: Sequence[Tuple[int, int]], settings: Settings,
                post_mask: int = POST_ALL, threads: int = 1) -> Families:
    """SearchDuplications::run with --trim (bin/asgart.rs:137-258): the suffix array covers strand[a..b] only, the LUT
    and every comparison read the whole strand. `trim` is the effective one (see effective_trim)."""
    t = as_strand(text_with_dollar)
    sa = np.ascontiguousarray(trimmed_suffix_array(t, trim), dtype=np.int64)
    ch = np.ascontiguousarray(np.array(chunks, dtype=np.uint64).reshape(-1, 2))
    h = lib().oracle_search_trim(_ptr(t), len(t), _ptr(sa), len(sa), _ptr(ch), len(ch), C.byref(settings), post_mask, threads)
    if not h:
        raise RefPanic("the reference panics on this input (FilterNs / ComputeScore slice or complement panic)")
    try:
        return _copy_result(h)
    finally:
        lib().oracle_result_free(h)


def post_steps(fam: Families, text_with_dollar, post_mask: int) -> Families:
    t = as_strand(text_with_dollar)
    off = np.ascontiguousarray(fam.fam_offsets, dtype=np.int64)
    fields = np.ascontiguousarray(fam.fields, dtype=np.uint64)
    ident = np.ascontiguousarray(fam.identity, dtype=np.float32)
    flags = np.ascontiguousarray(fam.flags, dtype=np.uint8)
    h = lib().oracle_result_from_arrays(_ptr(off), len(off) - 1, _ptr(fields), _ptr(ident), _ptr(flags))
    try:
        if lib().oracle_result_post(h, _ptr(t), len(t), post_mask) != 0:
            raise RefPanic("the reference panics on this input (FilterNs / ComputeScore: '$' under complement, or arm past the strand)")
        return _copy_result(h)
    finally:
        lib().oracle_result_free(h)


# ------------------------------------------------------------------ prepare_data / JSON
class Prepared:
    """prepare_data output (bin/asgart.rs:273-471): strand incl. '$', fragment map, chunks."""

    def __init__(self, handle):
        L = lib()
        self.h = handle
        err = L.oracle_prepared_error(handle)
        if err:
            msg = err.decode()
            L.oracle_prepared_free(handle)
            self.h = None
            raise IOError(msg)
        n = L.oracle_prepared_strand_len(handle)
        buf = (C.c_uint8 * n).from_address(L.oracle_prepared_strand(handle))
        self.strand = np.frombuffer(buf, dtype=np.uint8).copy()
        nc = L.oracle_prepared_n_chunks(handle)
        ch = np.zeros((nc, 2), dtype=np.uint64)
        L.oracle_prepared_chunks(handle, _ptr(ch))
        self.chunks = [(int(a), int(b)) for a, b in ch]
        self.map = []
        for i in range(L.oracle_prepared_n_fragments(handle)):
            pos, ln = C.c_uint64(), C.c_uint64()
            nm = L.oracle_prepared_fragment(handle, i, C.byref(pos), C.byref(ln))
            self.map.append((nm.decode(), pos.value, ln.value))

    @classmethod
    def from_files(cls, files: Sequence[str], skip_masked: bool = False) -> "Prepared":
        return cls(lib().oracle_prepare("\n".join(files).encode(), int(skip_masked)))

    @classmethod
    def from_memory(cls, strand_no_dollar, fragments: Sequence[Tuple[str, int, int]], file_names: str = "mem.fa"):
        t = as_strand(strand_no_dollar)
        pos = np.array([f[1] for f in fragments], dtype=np.uint64)
        ln = np.array([f[2] for f in fragments], dtype=np.uint64)
        names = "\n".join(f[0] for f in fragments).encode()
        return cls(lib().oracle_prepare_from_memory(file_names.encode(), _ptr(t), len(t), names, _ptr(pos), _ptr(ln),
                                                    len(fragments)))

    def to_json(self, settings: Settings, fam: Families) -> str:
        L = lib()
        off = np.ascontiguousarray(fam.fam_offsets, dtype=np.int64)
        fields = np.ascontiguousarray(fam.fields, dtype=np.uint64)
        ident = np.ascontiguousarray(fam.identity, dtype=np.float32)
        flags = np.ascontiguousarray(fam.flags, dtype=np.uint8)
        rh = L.oracle_result_from_arrays(_ptr(off), len(off) - 1, _ptr(fields), _ptr(ident), _ptr(flags))
        try:
            p = L.oracle_to_json(self.h, C.byref(settings), rh)
            s = C.string_at(p).decode()
            L.oracle_free_string(p)
            return s
        finally:
            L.oracle_result_free(rh)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_prepared_free(self.h)
            self.h = None


def run_files(files: Sequence[str], settings: Settings, threads: int = 1) -> str:
    """The whole reference pipeline on FASTA files -> JSON text (bin/asgart.rs:731-822 + exporters.rs:12-25)."""
    prep = Prepared.from_files(files, bool(settings.skip_masked))
    mask = POST_ALL | (POST_COMPUTE_SCORE if settings.compute_score else 0)   # bin/asgart.rs:744-746
    eff = effective_trim((settings.trim_a, settings.trim_b), len(prep.strand)) if settings.has_trim else None
    if eff is not None:                                                       # bin/asgart.rs:142-147
        return prep.to_json(settings, search_trim(prep.strand, eff, prep.chunks, settings, mask, threads))
    sa = best_suffix_array(prep.strand)
    out = search(prep.strand, sa, prep.chunks, settings, mask, threads)
    return prep.to_json(settings, out.families)
