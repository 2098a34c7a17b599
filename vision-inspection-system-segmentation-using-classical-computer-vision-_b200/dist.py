"""Multi-GPU plumbing: shard images by rank, gather the per-unit record table.

The path shards naturally (every unit of every image is independent):
image i goes to rank i % world; the only exchange is one all-gather of the
64-byte per-unit records (torch.distributed: NCCL over NVLink on GPUs, gloo in
the CPU tests).  Masks stay on the GPU that produced them."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from ._lib import RECORD_DTYPE


def shard_images(n_images: int, rank: int, world: int) -> List[int]:
    """Global image indices owned by `rank` (round robin)."""
    return list(range(rank, n_images, world))


def max_shard(n_images: int, world: int) -> int:
    return (n_images + world - 1) // world


def gather_record_table(local_records, n_images: int, n_units: int, group=None):
    """All-gather the ranks' record blocks into one image-major table.

    local_records: this rank's records, uint8 tensor [n_local * n_units, 64] (CUDA
    for NCCL, CPU for gloo) in the order of shard_images().  Returns a numpy
    structured array [n_images, n_units] with `image` rewritten to the global index."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    cap = max_shard(n_images, world) * n_units
    buf = torch.zeros((cap, 64), dtype=torch.uint8, device=local_records.device)
    buf[: local_records.shape[0]] = local_records
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    table = np.zeros((n_images, n_units), RECORD_DTYPE)
    for r in range(world):
        imgs = shard_images(n_images, r, world)
        rec = out[r].cpu().numpy().view(RECORD_DTYPE).reshape(-1)[: len(imgs) * n_units].reshape(len(imgs), n_units).copy()
        for k, gi in enumerate(imgs):
            rec[k]['image'] = gi
            table[gi] = rec[k]
    return table
