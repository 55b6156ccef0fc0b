"""Multi-process check of the sharded index build (run under torchrun on N GPUs of one box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_index_check.py [config] [scale_n]
Every rank loads the same synthetic strand, the ranks build the index together (NCCL + CUDA IPC), then each rank
checks its copy on the device (sufcheck) and rank 0 compares families and LUT with a lone build."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asgart_b200 as ab  # noqa: E402
from asgart_b200.dist import join_index_group, sharded_search  # noqa: E402


def main():
    config = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    scale_n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    flags = {1: dict(), 2: dict(reverse=True, complement=True, skip_masked=True), 3: dict(reverse=True, complement=True, max_cardinality=500),
             4: dict(reverse=True, complement=True)}[config]
    st = ab.RunSettings(**flags)
    g, fr = ab.synth_genome(config, scale_n=scale_n, threads=max(1, (os.cpu_count() or 8) // world))
    prep = ab.Prepared.from_memory(ab.normalise(g, st.skip_masked), fr)
    strand = np.array(prep.strand)
    ctx = ab.Context(local)
    ctx.load_strand(strand)
    join_index_group(ctx, dev)
    for rep in range(3):
        torch.cuda.synchronize(dev)
        dist.barrier()
        t0 = time.perf_counter()
        ctx.build_index()
        t1 = time.perf_counter()
        fam = sharded_search(ctx, prep.chunks, st, ab.POST_ALL, dev)
        t2 = time.perf_counter()
        print(f"[rank {rank}] rep {rep}: sharded build {1e3 * (t1 - t0):.2f} ms, search {1e3 * (t2 - t1):.2f} ms, {fam.n_families} families", flush=True)
    bad = ctx.check_sa()
    print(f"[rank {rank}] sufcheck violations: {bad}", flush=True)
    ok = bad == 0
    # the collective load (every rank copies 1/world of the strand, the pieces travel over NVLink) gives the same strand
    ctx.load_strand(strand)
    ok = ok and np.array_equal(ctx.download_strand(), strand)
    ctx.build_index()
    fam_again = sharded_search(ctx, prep.chunks, st, ab.POST_ALL, dev)
    ok = ok and fam_again.digest() == fam.digest()
    print(f"[rank {rank}] families digest {fam.digest()}", flush=True)
    if rank == 0:
        lut = ctx.download_lut()
        with ab.Context(local) as one:
            one.load_strand(strand)
            t0 = time.perf_counter()
            one.build_index()
            t1 = time.perf_counter()
            ref = one.search(prep.chunks, st, ab.POST_ALL)
            lut1 = one.download_lut()
            print(f"[rank 0] lone build {1e3 * (t1 - t0):.2f} ms (first call, cold pool)", flush=True)
            t0 = time.perf_counter()
            one.build_index()
            print(f"[rank 0] lone build {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
        ne = lut1[1] > lut1[0]
        ok = ok and fam.as_lists() == ref.as_lists() and np.array_equal(lut[0][ne], lut1[0][ne]) and np.array_equal(lut[1][ne], lut1[1][ne])
        print("DIST INDEX CHECK:", "OK" if ok else "MISMATCH", flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    ctx.dist_shutdown()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
