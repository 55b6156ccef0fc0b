"""Generate tests/golden/sa_golden.json from the REFERENCE's own C library (oracle/_ref/libdivsufsort64.so, compiled in
place from /root/reference/libdivsufsort): suffix arrays (divsufsort64, validated by sufcheck64) and non-empty 8-mer LUT
intervals (sa_searchb64, the call Searcher::new makes, src/searcher.rs:118-128) for small fixed texts.

Run here (where /root/reference exists):  python tests/golden/make_golden.py
The fixtures travel to the GPU box; /root/reference does not.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from tests import kat  # noqa: E402


def main():
    R = oracle.ref()
    assert R is not None, "needs oracle/_ref (make -C oracle ref)"
    rng = np.random.default_rng(424242)
    texts = {}
    texts["dna600"] = np.concatenate([kat.rand_dna(rng, 600), [ord("$")]]).astype(np.uint8)
    t = kat.rand_dna(rng, 900); t[100:160] = ord("N"); t[500:620] = t[300:420]; t[880:] = ord("N")
    texts["dna900_N_dup"] = np.concatenate([t, [ord("$")]]).astype(np.uint8)
    texts["polyN"] = np.frombuffer(b"ACGT" + b"N" * 120 + b"TGCA" + b"N" * 40 + b"$", dtype=np.uint8).copy()
    texts["bytes300_noterm"] = rng.integers(0, 256, size=300).astype(np.uint8)
    texts["banana"] = np.frombuffer(b"banana", dtype=np.uint8).copy()
    texts["mississippi$"] = np.frombuffer(b"mississippi$", dtype=np.uint8).copy()
    cases = {}
    alphabet = b"ATGCN"
    for name, t in texts.items():
        sa = oracle.ref_divsufsort64(t)
        assert oracle.ref_sufcheck64(t, sa) == 0
        entry = {"text_hex": t.tobytes().hex(), "sa": sa.tolist()}
        if name.startswith("dna") or name == "polyN":
            lutd = {}
            # every 8-mer occurring in the text, plus a few absent ones, through the reference's sa_searchb64
            seen = {bytes(t[i:i + 8]) for i in range(len(t) - 8)}
            for p in seen:
                if any(c not in alphabet for c in p):
                    continue
                pa = np.frombuffer(p, dtype=np.uint8)
                out = C.c_int64()
                cnt = R.sa_searchb64(t.ctypes.data, len(t), pa.ctypes.data, 8, sa.ctypes.data, len(sa), C.byref(out), 0, len(sa))
                assert cnt > 0
                lutd[format(int.from_bytes(p, "little"), "016x")] = [out.value, out.value + cnt]
            entry["lut_nonempty"] = lutd
        cases[name] = entry
    with open(os.path.join(os.path.dirname(__file__), "sa_golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py (reference libdivsufsort @ 523b07c submodule, divsufsort64 + sa_searchb64)",
                   "cases": cases}, f, separators=(",", ":"))
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
