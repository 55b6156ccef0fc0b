"""ctypes wrapper for tests/emul/libemul.so (host build of the kernels' logic headers; test-only)."""
import ctypes as C
import os
import subprocess

import numpy as np

from asgart_b200 import _lib as ablib

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "emul.cpp")
SO = os.path.join(HERE, "emul", "libemul.so")
_L = None


def lib():
    global _L
    if _L is None:
        deps = [SRC] + [os.path.join(HERE, "..", "asgart_b200", "csrc", f) for f in ("kmer_core.h", "automaton_core.h", "fasta_core.h")]
        if not os.path.exists(SO) or any(os.path.getmtime(SO) < os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", SO, SRC])
        L = C.CDLL(SO)
        L.emul_search.restype = C.c_void_p
        L.emul_search.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(ablib.Settings), C.c_void_p]
        L.emul_lut.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.emul_literal_bucket.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.emul_code_tab_mismatches.restype = C.c_int
        L.emul_pack_check.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.emul_window.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        for n in ("emul_result_n_families", "emul_result_n_sds"):
            getattr(L, n).restype = C.c_int64
            getattr(L, n).argtypes = [C.c_void_p]
        L.emul_result_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.emul_result_free.argtypes = [C.c_void_p]
        L.emul_ingest.restype = C.c_int64
        L.emul_ingest.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        _L = L
    return _L


def pack_check(text: np.ndarray):
    """(fast-path words, fast-path mismatches, table-path mismatches) of the packing kernel's logic on `text`, all four modes"""
    t = np.ascontiguousarray(text, dtype=np.uint8)
    out = np.zeros(3, dtype=np.uint64)
    lib().emul_pack_check(t.ctypes.data, len(t), out.ctypes.data)
    return tuple(int(x) for x in out)


def search(strand, sa, chunks, settings_c):
    L = lib()
    t = np.ascontiguousarray(strand, dtype=np.uint8)
    sa = np.ascontiguousarray(sa, dtype=np.int64)
    ch = np.ascontiguousarray(np.array(chunks, dtype=np.uint64).reshape(-1, 2))
    ctr = np.zeros(6, dtype=np.uint64)
    h = L.emul_search(t.ctypes.data, len(t), sa.ctypes.data, ch.ctypes.data, len(ch), C.byref(settings_c), ctr.ctypes.data)
    try:
        nf, ns = L.emul_result_n_families(h), L.emul_result_n_sds(h)
        off = np.zeros(nf + 1, dtype=np.int64)
        fields = np.zeros((ns, 4), dtype=np.uint64)
        L.emul_result_copy(h, off.ctypes.data, fields.ctypes.data)
    finally:
        L.emul_result_free(h)
    fams = [[tuple(int(x) for x in fields[j]) for j in range(off[f], off[f + 1])] for f in range(nf)]
    return fams, dict(zip(["probes", "searched", "skipped_n", "skipped_card", "matches", "alg_bytes"], (int(x) for x in ctr)))


def ingest(blob: bytes, skip_masked: bool):
    """fasta_ingest.cuh's passes emulated on the CPU: (strand without '$', [(name, position, length)], [(start, length)])."""
    L = lib()
    b = np.frombuffer(blob, dtype=np.uint8) if len(blob) else np.zeros(0, dtype=np.uint8)
    cap = len(b) + 1
    strand = np.zeros(cap, dtype=np.uint8)
    rec_off = np.zeros(cap, dtype=np.uint64)
    rec_pos = np.zeros(cap, dtype=np.uint64)
    chunks = np.zeros((2 * cap + 2, 2), dtype=np.uint64)
    frag = np.zeros((cap, 2), dtype=np.uint64)
    n_rec, n_chunks, n_frag = C.c_int64(), C.c_int64(), C.c_int64()
    kept = L.emul_ingest(b.ctypes.data if len(b) else None, len(b), int(skip_masked), strand.ctypes.data, cap, rec_off.ctypes.data,
                         rec_pos.ctypes.data, cap, C.byref(n_rec), chunks.ctypes.data, len(chunks), C.byref(n_chunks),
                         frag.ctypes.data, cap, C.byref(n_frag))
    if kept == -3:
        raise IOError("Unable to parse: expected > at record start")
    assert kept >= 0, kept
    names = []
    for r in range(n_rec.value):       # record ids: header up to the first white space (what api.cu reads out of the file)
        o = int(rec_off[r]) + 1
        e = o
        while e < len(blob) and blob[e] not in b" \t\n\x0b\x0c\r":
            e += 1
        names.append(blob[o:e].decode())
    fr = [(names[i], int(frag[i][0]), int(frag[i][1])) for i in range(n_frag.value)]
    return strand[:kept].copy(), fr, [(int(a), int(c)) for a, c in chunks[:n_chunks.value]]
