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
        deps = [SRC] + [os.path.join(HERE, "..", "asgart_b200", "csrc", f) for f in ("kmer_core.h", "automaton_core.h")]
        if not os.path.exists(SO) or any(os.path.getmtime(SO) < os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", SO, SRC])
        L = C.CDLL(SO)
        L.emul_search.restype = C.c_void_p
        L.emul_search.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(ablib.Settings), C.c_void_p]
        L.emul_lut.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.emul_window.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        for n in ("emul_result_n_families", "emul_result_n_sds"):
            getattr(L, n).restype = C.c_int64
            getattr(L, n).argtypes = [C.c_void_p]
        L.emul_result_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.emul_result_free.argtypes = [C.c_void_p]
        _L = L
    return _L


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
