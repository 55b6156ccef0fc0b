import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, asgart_b200 as ab, oracle
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
g, fr = ab.synth_genome(2, scale_n=scale)
t = ab.normalise(g, True)
strand = np.concatenate([t, np.frombuffer(b"$", dtype=np.uint8)])
want = oracle.ref_divsufsort64(strand)
for trial in range(3):
    got = ab.r_divsufsort(strand)
    bad = np.nonzero(got != want)[0]
    print("trial", trial, "n", len(strand), "mismatches", len(bad), "perm", bool((np.sort(got) == np.arange(len(strand))).all()))
    if len(bad):
        i = int(bad[0])
        print(" first bad SA index", i, "last", int(bad[-1]))
        for k in range(max(0, i - 1), min(len(strand), i + 4)):
            a, b = int(got[k]), int(want[k])
            print("  ", k, "got", a, bytes(strand[a:a + 48]).decode(), "| want", b, bytes(strand[b:b + 48]).decode())
        # run lengths at mismatching suffixes
        def runlen(p):
            x = strand[p]; e = p
            while e < len(strand) and strand[e] == x: e += 1
            return e - p
        rl = [runlen(int(got[k])) for k in bad[:2000]]
        print(" run lengths of first mismatching suffixes: min", min(rl), "max", max(rl), "share >=21:", np.mean(np.array(rl) >= 21))
