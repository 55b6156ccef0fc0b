"""Developer probe (torchrun): the sharded step of bench.py, phase by phase."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import asgart_b200 as ab

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
st = ab.RunSettings(reverse=True, complement=True, skip_masked=True)
g, fr = ab.synth_genome(2)
prep = ab.Prepared.from_memory(ab.normalise(g, True), fr)
strand = np.array(prep.strand)
torch.cuda.set_device(local)
ctx = ab.Context(local)
ctx.load_strand(strand)

# ---- the sharded step, phase by phase (wall clock, after a device sync)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from asgart_b200.dist import all_gather_bytes
dev = torch.device("cuda", local)
acc = {"build": 0.0, "shard": 0.0, "gather": 0.0, "finish": 0.0}
for it in range(8):
    if it == 3:
        acc = {k: 0.0 for k in acc}
        ctx.reset_stats()
    t0 = time.perf_counter(); ctx.build_index()
    t1 = time.perf_counter(); part = ctx.search_shard(prep.chunks, st, rank, world)
    t2 = time.perf_counter(); parts = all_gather_bytes(part, dev)
    t3 = time.perf_counter(); fam = ctx.finish(prep.chunks, st, parts, ab.POST_ALL)
    t4 = time.perf_counter()
    acc["build"] += t1 - t0; acc["shard"] += t2 - t1; acc["gather"] += t3 - t2; acc["finish"] += t4 - t3
s = ctx.stats()
print(json.dumps({"tag": "sharded step", "rank": rank, **{k: round(v / 5 * 1e3, 2) for k, v in acc.items()},
                  "ms_sa_build": round(s["ms_sa_build"] / 5, 2), "ms_automaton": round(s["ms_automaton"] / 5, 2), "part_bytes": len(part)}), flush=True)
dist.barrier()
dist.destroy_process_group()
ctx.close()
