"""File-list sharding for the one-process-per-GPU launch (torchrun), SURVEY.md section 8(e).

Each file's digest depends on nothing else, so the list is cut into one contiguous run per
rank, balanced by padded 128-byte block count; digests (64 bytes per file) are gathered on
rank 0 by index.  That gather is the only communication and it is not on the data path, so
it runs over whatever backend the process group has (nccl on the GPU box, gloo in the CPU
tests).  Inside one process, libsnapgpu shards across its bound devices itself
(csrc/snapgpu.cu: shard_items).
"""
from __future__ import annotations

import numpy as np


def contiguous_shards(lengths, world: int) -> list[tuple[int, int]]:
    """[(begin, end)) per rank: contiguous runs with near-equal total (L+144)//128."""
    lengths = np.asarray(lengths, dtype=np.uint64)
    n = len(lengths)
    if world <= 1:
        return [(0, n)]
    blocks = ((lengths + np.uint64(144)) // np.uint64(128)).astype(np.float64)
    csum = np.concatenate([[0.0], np.cumsum(blocks)])
    total = csum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(csum, target, side="left"))
        # pick the boundary closer to the target
        if k > 0 and abs(csum[k - 1] - target) <= abs(csum[min(k, n)] - target):
            k -= 1
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_digests(local: np.ndarray, shards: list[tuple[int, int]], rank: int, world: int, group=None):
    """Rank 0 returns the (n, 64) digest array of the whole list; other ranks return None."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    most = max(e - b for b, e in shards)
    mine = torch.zeros((most, 64), dtype=torch.uint8, device=dev)
    if len(local):
        mine[: len(local)] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    bucket = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(bucket, mine, group=group)
    if rank != 0:
        return None
    n = shards[-1][1]
    out = np.zeros((n, 64), dtype=np.uint8)
    for r, (b, e) in enumerate(shards):
        out[b:e] = bucket[r][: e - b].cpu().numpy()
    return out
