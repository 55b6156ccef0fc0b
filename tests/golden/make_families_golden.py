#!/usr/bin/env python
"""Full-size golden digests of the BASELINE configs, made OFFLINE on the CPU by the oracle — test infrastructure.

    python tests/golden/make_families_golden.py 1 3 5 4        # writes tests/golden/families_c{N}.json

For each config the script generates the synthetic workload at BASELINE size (same generator and seeds as bench.py and
the GPU tests), runs prepare_data's chunker as restated by the oracle, builds the suffix array with the REFERENCE's own
libdivsufsort64 (oracle/_ref, compiled in place from /root/reference/libdivsufsort — so it needs this container),
runs the oracle port of SearchDuplications::run + the four post-steps (src/bin/asgart.rs:137-258, :33-96) on all host
threads and records

  * families_sha256   — asgart_b200.api.families_digest of the result (reference order, bit-exact fields)
  * json_sha256       — sha256 of the JSON text as JSONExporter::save would write it (src/exporters.rs:12-25)
  * sa_fingerprint    — asgart_b200.api.sa_fingerprint_host of divsufsort64's output (order-sensitive 64-bit sum)
  * counts (families, duplicons, probes, matches …), the chunk list digest and the oracle's phase seconds.

`tests/test_gpu_golden.py` compares the CUDA path's digests with these on the GPU box, where /root/reference is absent.
C4 needs about 30 GB of host memory and 15-25 minutes (the suffix array is single-threaded, as build.rs builds it).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import asgart_b200 as ab  # noqa: E402  (synthetic generator + digest helpers only: host code, no GPU)
import oracle  # noqa: E402
from asgart_b200.api import families_digest, sa_fingerprint_host  # noqa: E402
from bench import CONFIG_FLAGS, CONFIG_NAMES, FULL_N, oracle_settings, oracle_workload  # noqa: E402


def make(config: int, scale_n: int = 0) -> dict:
    threads = os.cpu_count() or 8
    t0 = time.perf_counter()
    st, prep = oracle_workload(config, scale_n)
    t_gen = time.perf_counter() - t0
    strand = prep.strand
    so = oracle_settings(st)
    t0 = time.perf_counter()
    sa = oracle.ref_divsufsort64(strand)
    t_sa = time.perf_counter() - t0
    assert sa[0] == len(strand) - 1
    fp = sa_fingerprint_host(sa)
    out = oracle.search(strand, sa, prep.chunks, so, oracle.POST_ALL, threads=threads)
    fam = out.families
    del sa
    js = prep.to_json(so, fam)
    ch = np.array(prep.chunks, dtype="<u8").reshape(-1, 2)
    return {
        "generator": "tests/golden/make_families_golden.py (oracle: reference libdivsufsort64 + restated Rust path)",
        "config": config, "workload": CONFIG_NAMES[config], "flags": CONFIG_FLAGS[config],
        "strand_bp": int(len(strand) - 1), "full_size": int(len(strand) - 1) == FULL_N[config],
        "n_chunks": len(prep.chunks), "searched_bp": int(ch[:, 1].sum()) if len(ch) else 0,
        "chunks_sha256": hashlib.sha256(ch.tobytes()).hexdigest(),
        "map": [[nm, int(p), int(ln)] for nm, p, ln in prep.map],
        "sa_fingerprint": fp,
        "families": int(len(fam.fam_offsets) - 1), "duplicons": int(len(fam.fields)),
        "families_sha256": families_digest(fam.fam_offsets, fam.fields, fam.identity, fam.flags),
        "json_sha256": hashlib.sha256(js.encode()).hexdigest(), "json_bytes": len(js),
        "counters": out.counters,
        "oracle_seconds": {"generate": round(t_gen, 2), "sa_divsufsort64_1thread": round(t_sa, 2), "lut": round(out.seconds["lut"], 2),
                           "search_automaton": round(out.seconds["search"], 2), "post": round(out.seconds["post"], 3),
                           "threads": threads, "host": "CPU container (no GPU), %d cores" % threads},
    }


if __name__ == "__main__":
    for c in [int(x) for x in sys.argv[1:]] or [1, 3, 5, 4]:
        row = make(c)
        path = os.path.join(HERE, f"families_c{c}.json")
        with open(path, "w") as f:
            json.dump(row, f, indent=1)
            f.write("\n")
        print(path, row["families"], "families", row["duplicons"], "duplicons", row["oracle_seconds"], flush=True)
