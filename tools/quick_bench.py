"""Developer timing probe (not the bench contract): python tools/quick_bench.py <config> [scale_n] [flags]
Generates the synthetic config, runs load -> build_index -> search on cuda:0 a few times and prints the stats."""
import json
import sys
import time

import numpy as np

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asgart_b200 as ab  # noqa: E402


def main():
    config = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    scale_n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    device = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    flags = {1: dict(), 2: dict(reverse=True, complement=True, skip_masked=True), 3: dict(reverse=True, complement=True),
             4: dict(reverse=True, complement=True), 0: dict()}[config]
    st = ab.RunSettings(compute_score=bool(os.environ.get("QB_SCORE")), **flags)   # QB_SCORE=1: with --compute-score
    t0 = time.time()
    g, fr = ab.synth_genome(config, scale_n=scale_n)
    prep = ab.Prepared.from_memory(ab.normalise(g, st.skip_masked), fr)
    print(f"generated n={len(g)} in {time.time() - t0:.2f}s; chunks={len(prep.chunks)}", flush=True)
    strand = np.array(prep.strand)
    with ab.Context(device) as ctx:
        for r in range(reps):
            ctx.reset_stats()
            t0 = time.time()
            ctx.load_strand(strand)
            t1 = time.time()
            ctx.build_index()
            t2 = time.time()
            fam = ctx.search(prep.chunks, st, ab.POST_ALL)
            t3 = time.time()
            s = ctx.stats()
            print(json.dumps({"rep": r, "wall_load": round(t1 - t0, 4), "wall_build": round(t2 - t1, 4), "wall_search": round(t3 - t2, 4),
                              "families": fam.n_families, "sds": len(fam.sds),
                              **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in s.items()}}), flush=True)


if __name__ == "__main__":
    main()
