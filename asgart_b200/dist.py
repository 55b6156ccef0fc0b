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


def join_index_group(ctx, device: torch.device | None = None):
    """Make `ctx.build_index()` a sharded, collective build over the ranks of the default process group (one process per
    GPU): rank 0 creates the library's NCCL unique id, torch.distributed carries it to the others."""
    from .api import dist_unique_id
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    dev = device if device is not None else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(dist_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    ctx.dist_init(rank, world, bytes(t.cpu().numpy().tobytes()))


class _DevMem:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface v3)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


def sharded_search_device(ctx, chunks: Sequence, settings, post_mask: int, device: torch.device):
    """NCCL path: the stage-A partial never leaves HBM. One tiny all-gather of the meta words, one all-gather of the
    (padded) device blobs over NVLink, then stage B on every rank from the gathered device memory."""
    rank, world = dist.get_rank(), dist.get_world_size()
    ptr, nbytes, meta = ctx.search_shard_dev(chunks, settings, rank, world)
    mine = torch.empty(5, dtype=torch.int64, device=device)
    mine.copy_(torch.from_numpy(np.concatenate([meta.astype(np.int64), [nbytes]])))
    allm = torch.empty(world * 5, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allm, mine)
    allm = allm.cpu().numpy().reshape(world, 5)
    cap = (int(allm[:, 4].max()) + 255) // 256 * 256
    cap = max(cap, 256)
    send = torch.empty(cap, dtype=torch.uint8, device=device)
    if nbytes:
        send[:nbytes].copy_(torch.as_tensor(_DevMem(ptr, nbytes), device=device))
    recv = torch.empty(world * cap, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(recv, send)
    torch.cuda.current_stream(device).synchronize()   # the library reads `recv` on its own stream
    base = recv.data_ptr()
    return ctx.finish_dev(chunks, settings, [base + r * cap for r in range(world)], allm[:, :4].astype(np.uint64), post_mask)


def sharded_search(ctx, chunks: Sequence, settings, post_mask: int, device: torch.device | None = None):
    """Stage A on this rank's probe range, exchange, stage B on the merged events. Same families on every rank."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if device is not None and device.type == "cuda" and dist.get_backend() == "nccl":
        return sharded_search_device(ctx, chunks, settings, post_mask, device)
    part = ctx.search_shard(chunks, settings, rank, world)
    parts = all_gather_bytes(part, device)
    return ctx.finish(chunks, settings, parts, post_mask)
