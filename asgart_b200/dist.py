"""Multi-GPU plumbing: one process per GPU (torch.distributed), strand + index replicated, probes partitioned by
position, one exchange of the per-rank stage-A partials (NCCL all-gather over NVLink; gloo on CPU for the tests),
then the automaton + post-steps on every rank (identical results everywhere, rank 0 reports).

The reference has no distributed mode (rayon threads only, src/bin/asgart.rs:201-240); this is the probe-position
sharding BASELINE.json:north_star prescribes.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def all_gather_bytes(buf: np.ndarray, device: torch.device | None = None) -> List[np.ndarray]:
    """All-gather variable-length byte blobs. Works on gloo (CPU tensors) and NCCL (tensors on `device`)."""
    world = dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    n = torch.tensor([len(buf)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    mine = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if len(buf):
        mine[:len(buf)] = torch.from_numpy(np.ascontiguousarray(buf, dtype=np.uint8)).to(dev)
    outs = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(outs, mine)
    return [o[:s].cpu().numpy() for o, s in zip(outs, sizes)]


def sharded_search(ctx, chunks: Sequence, settings, post_mask: int, device: torch.device | None = None):
    """Stage A on this rank's probe range, exchange, stage B on the merged events. Same families on every rank."""
    rank, world = dist.get_rank(), dist.get_world_size()
    part = ctx.search_shard(chunks, settings, rank, world)
    parts = all_gather_bytes(part, device)
    return ctx.finish(chunks, settings, parts, post_mask)
