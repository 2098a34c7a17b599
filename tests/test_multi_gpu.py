"""Two ranks on two GPUs (NCCL): the per-unit record table gathered by the kernel over peer memory
(vi_b200.dist.RecordExchange) and by an NCCL all-gather (gather_record_table) equals the table of a 1-rank run of the
same global image list.  Skips below 2 GPUs (the CPU suite covers the host logic on gloo: test_dist_gloo.py)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_IMAGES = 7            # odd on purpose: the ranks' shards differ in length
FRAME_H, FRAME_W = 760, 1456


def _boxes():
    import vi_b200
    return vi_b200.generate_grid((21, 18, 316, 315), 4, 2, 1, 1, 30, 40, 0, 0)


def _frames():
    from vi_b200 import synth
    boxes = [b for b, _ in _boxes()]
    return np.stack([synth.make_frame(300 + i, boxes, H=FRAME_H, W=FRAME_W) for i in range(N_IMAGES)])


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import vi_b200
    from vi_b200 import dist as vdist
    from vi_b200.grid import Grid
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    try:
        boxes = _boxes()
        frames = _frames()
        excl = [{'shape': 'rect', 'x': 50, 'y': 60, 'w': 70, 'h': 30}, {'shape': 'circle', 'cx': 200, 'cy': 180, 'r': 25}]
        refc = {i: (158.0 + 0.25 * i, 157.0 - 0.5 * i) for i in range(len(boxes))}
        grid = Grid(boxes=boxes, exclusions=excl, ref_centroids=refc)
        insp = vi_b200.Inspector(rank)
        insp.configure(grid, is_reference=False)
        mine = vdist.shard_images(N_IMAGES, rank, world)
        d = torch.from_numpy(frames[mine]).cuda()
        ex = vdist.RecordExchange(insp, N_IMAGES)
        for _ in range(3):                                   # repeated batches overwrite the same table entries
            rec, seg, dfm = insp.inspect_batch(d)
        ex.complete()
        table = ex.table()
        table_nccl = vdist.gather_record_table(rec, N_IMAGES, len(boxes))
        # local records carry the global image index too
        loc = rec.cpu().numpy().view(vi_b200.RECORD_DTYPE).reshape(len(mine), len(boxes))
        assert [int(x) for x in loc['image'][:, 0]] == mine
        ex.close()
        if rank == 0:
            insp1 = vi_b200.Inspector(0)
            insp1.configure(grid, is_reference=False)
            r1, _, _ = insp1.inspect_batch(torch.from_numpy(frames).cuda())
            torch.cuda.synchronize()
            np.save(os.path.join(out_dir, "one_rank.npy"), r1.cpu().numpy())
        np.save(os.path.join(out_dir, f"table_{rank}.npy"), table.view(np.uint8))
        np.save(os.path.join(out_dir, f"nccl_{rank}.npy"), table_nccl.view(np.uint8))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_table_equals_one_rank_table(tmp_path):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    import vi_b200
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    one = np.load(tmp_path / "one_rank.npy").reshape(-1).view(vi_b200.RECORD_DTYPE)
    assert (one['dx'] != 0).any() and (one['status'] == vi_b200.STATUS_NG).any()
    for rank in range(2):
        for kind in ("table", "nccl"):
            t = np.load(tmp_path / f"{kind}_{rank}.npy").reshape(-1).view(vi_b200.RECORD_DTYPE)
            assert t.shape == one.shape
            for k in vi_b200.RECORD_DTYPE.names:
                a, b = t[k], one[k]
                same = np.array_equal(a, b) or (a.dtype.kind == 'f' and np.array_equal(np.isnan(a), np.isnan(b)) and
                                                np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]))
                assert same, (kind, rank, k)
