"""Developer probe: distribution of automaton segments for a config (needs a GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, asgart_b200 as ab
config = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = int(sys.argv[2]) if len(sys.argv) > 2 else 0
flags = {1: dict(), 2: dict(reverse=True, complement=True, skip_masked=True), 3: dict(reverse=True, complement=True), 4: dict(reverse=True, complement=True)}[config]
st = ab.RunSettings(**flags)
g, fr = ab.synth_genome(config, scale_n=scale)
prep = ab.Prepared.from_memory(ab.normalise(g, st.skip_masked), fr)
with ab.Context(0) as ctx:
    ctx.load_strand(np.array(prep.strand)); ctx.build_index()
    part = ctx.search_shard(prep.chunks, st, 0, 1)
hdr = part[:48].view(np.uint64)
nb, ne, nm = int(hdr[3]), int(hdr[4]), int(hdr[5])
o = 48
bits = part[o:o + nb * 4].view(np.uint32); o += nb * 4
ev_probe = part[o:o + ne * 8].view(np.uint64).astype(np.int64); o += ne * 8
ev_cnt = part[o:o + ne * 4].view(np.uint32).astype(np.int64)
proc = np.unpackbits(bits.view(np.uint8), bitorder="little")
pre = np.concatenate([[0], np.cumsum(proc)])
k, s = st.probe_size, st.probe_size // 2
q = max(1, -(-st.max_gap_size // s))
# chunk of each event
bases = np.cumsum([0] + [max(0, -(-(c[1] - k - s) // s)) if c[1] >= st.min_duplication_length and c[1] >= k + s else 0 for c in prep.chunks])
chunk = np.searchsorted(bases, ev_probe, side="right") - 1
t = pre[ev_probe] - pre[bases[chunk]]
head = np.ones(ne, dtype=bool)
head[1:] = (chunk[1:] != chunk[:-1]) | ((t[1:] - t[:-1]) > q)
seg = np.cumsum(head) - 1
nseg = seg[-1] + 1
n_ev = np.bincount(seg, minlength=nseg)
sum_cnt = np.bincount(seg, weights=ev_cnt, minlength=nseg)
max_cnt = np.zeros(nseg, dtype=np.int64); np.maximum.at(max_cnt, seg, ev_cnt)
print(f"events {ne} matches {nm} segments {nseg}; events/segment: mean {n_ev.mean():.1f} max {n_ev.max()}; matches/segment max {int(sum_cnt.max())}")
order = np.argsort(-n_ev)[:8]
print("longest segments (events, matches, max cnt):", [(int(n_ev[i]), int(sum_cnt[i]), int(max_cnt[i])) for i in order])
order = np.argsort(-sum_cnt)[:8]
print("heaviest segments (events, matches, max cnt):", [(int(n_ev[i]), int(sum_cnt[i]), int(max_cnt[i])) for i in order])
print("cnt histogram:", {b: int(((ev_cnt > a) & (ev_cnt <= b)).sum()) for a, b in [(0, 1), (1, 4), (4, 32), (32, 128), (128, 512)]})
