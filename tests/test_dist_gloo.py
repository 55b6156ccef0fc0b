"""world_size-2 gloo test of the multi-GPU host logic (asgart_b200/dist.py): variable-length all-gather of the
serialised stage-A partials, and the shard partition arithmetic of the C ABI (restated here)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from asgart_b200.dist import all_gather_bytes
    rng = np.random.default_rng(100 + rank)
    mine = rng.integers(0, 256, size=1000 * (rank + 1) + 7 * rank, dtype=np.uint8) if rank != 1 else np.zeros(0, np.uint8)
    parts = all_gather_bytes(mine)
    ok = len(parts) == world
    for r in range(world):
        exp = np.random.default_rng(100 + r).integers(0, 256, size=1000 * (r + 1) + 7 * r, dtype=np.uint8) if r != 1 else np.zeros(0, np.uint8)
        ok = ok and np.array_equal(parts[r], exp)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_all_gather_bytes_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_all_gather_bytes_gloo_world3_empty_and_large():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True), (2, True)]


def test_shard_ranges_partition_probe_space():
    """shard_range (api.cu): consecutive, 32-aligned, covering [0, total) for every world size."""
    def shard_range(total, shard, n):
        per = -(-(-(-total // n)) // 32) * 32 or 32
        return min(total, shard * per), min(total, (shard + 1) * per)
    for total in (0, 1, 31, 32, 33, 1000, 5_722_741, 308_826_983):
        for n in (1, 2, 3, 4, 8):
            ranges = [shard_range(total, r, n) for r in range(n)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            for (a, b), (c, d) in zip(ranges, ranges[1:]):
                assert b == c and a <= b
            assert all(a % 32 == 0 or a == total for a, _ in ranges)   # empty trailing shards sit at `total`
