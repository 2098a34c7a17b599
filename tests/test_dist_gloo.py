"""Host-side multi-rank logic on CPU: world_size-2 gloo, image sharding and the
all-gather of the per-unit record table (the only collective of the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vi_b200 import RECORD_DTYPE
from vi_b200.dist import gather_record_table, max_shard, shard_images


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_records(global_image, n_units):
    rec = np.zeros(n_units, RECORD_DTYPE)
    rec['unit'] = np.arange(n_units)
    rec['defect_area'] = global_image * 1000 + np.arange(n_units)
    rec['status'] = (np.arange(n_units) + global_image) % 3
    rec['cx'] = global_image + 0.5
    return rec


def _worker(rank, world, port, n_images, n_units, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_images(n_images, rank, world)
    local = np.concatenate([_fake_records(gi, n_units) for gi in mine]) if mine else np.zeros(0, RECORD_DTYPE)
    local['image'] = np.repeat(np.arange(len(mine)), n_units)          # batch-local indices, as the kernel writes them
    t = torch.from_numpy(local.view(np.uint8).reshape(-1, 64).copy())
    table = gather_record_table(t, n_images, n_units)
    ok = True
    for gi in range(n_images):
        exp = _fake_records(gi, n_units)
        ok &= bool(np.array_equal(table[gi]['defect_area'], exp['defect_area']))
        ok &= bool(np.array_equal(table[gi]['status'], exp['status']))
        ok &= bool((table[gi]['image'] == gi).all())
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [5, 8])
def test_record_gather_two_ranks(n_images):
    world, n_units = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_sharding_is_a_partition():
    for n, w in ((1024, 8), (5, 2), (3, 4), (64, 1)):
        seen = sorted(i for r in range(w) for i in shard_images(n, r, w))
        assert seen == list(range(n))
        assert max(len(shard_images(n, r, w)) for r in range(w)) == max_shard(n, w)
