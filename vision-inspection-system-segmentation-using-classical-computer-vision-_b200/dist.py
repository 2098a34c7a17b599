"""Multi-GPU plumbing: shard images by rank, gather the per-unit record table.

The path shards naturally (every unit of every image is independent):
image i goes to rank i % world; the only exchange is the table of 64-byte per-unit
records.  Two ways to get it:

* `gather_record_table` -- one all-gather through torch.distributed (NCCL over NVLink
  on GPUs, gloo in the CPU tests), after the batch;
* `RecordExchange` -- fused into the kernel: every rank's table lives in peer-mapped
  device memory (CUDA IPC) and the kernel stores each record into all of them over
  NVLink as it finishes a unit.  No collective runs on the data path; the table is
  complete after the ranks' streams are synchronised and a barrier is passed.

Masks stay on the GPU that produced them."""
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


class RecordExchange:
    """Kernel-fused gather of the record table over peer-mapped memory (include/vi_b200.h, vi_set_record_peers).

    Every rank calls this with the same `n_images` (global) while a process group (NCCL) is up.  Afterwards
    `inspector.inspect_batch(local_frames)` -- local image k being global image rank + k * world -- also stores each
    record at [global image, unit] of every rank's table.  `table()` (after `complete()`) returns it."""

    def __init__(self, inspector, n_images: int, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from ._lib import check
        self.insp, self.n_images, self.group = inspector, int(n_images), group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("RecordExchange: one NVSwitch domain, at most 8 ranks")
        lib, ctx = inspector._lib, inspector._ctx
        self.n_rec = self.n_images * inspector.n_units
        self._own = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        check(lib.vi_peer_table_create(ctx, self.n_rec, C.byref(self._own), C.cast(handle, C.c_void_p)))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self._ptrs, self._opened = [], []
        for r in range(self.world):
            if r == self.rank:
                self._ptrs.append(self._own.value)
                continue
            p = C.c_void_p()
            buf = (C.c_uint8 * 64).from_buffer_copy(handles[r])
            check(lib.vi_peer_table_open(ctx, C.cast(buf, C.c_void_p), C.byref(p)))
            self._ptrs.append(p.value)
            self._opened.append(p.value)
        arr = (C.c_void_p * self.world)(*self._ptrs)
        check(lib.vi_set_record_peers(ctx, C.cast(arr, C.c_void_p), self.world, self.world, self.rank))
        dist.barrier(group=group)                      # every rank has mapped every table before anyone launches

    def complete(self):
        """All ranks' batches are done and their stores have landed: the local table is the whole table."""
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        torch.cuda.synchronize()

    def table(self):
        """The local copy of the gathered table as a numpy structured array [n_images, n_units] (call complete() first)."""
        from ._lib import check
        out = np.empty(self.n_rec, RECORD_DTYPE)
        check(self.insp._lib.vi_peer_table_read(self.insp._ctx, self._own, self.n_rec, out.ctypes.data))
        return out.reshape(self.n_images, self.insp.n_units)

    def close(self):
        from ._lib import check
        lib, ctx = self.insp._lib, self.insp._ctx
        import torch
        import torch.distributed as dist
        check(lib.vi_set_record_peers(ctx, None, 0, 1, 0))
        torch.cuda.synchronize()
        dist.barrier(group=self.group)                 # nobody still stores into a table that is about to go
        for p in self._opened:
            check(lib.vi_peer_table_close(ctx, p))
        self._opened = []
        if self._own.value:
            check(lib.vi_peer_table_destroy(ctx, self._own))
            self._own.value = None
