"""world_size-2 gloo tests of the multi-GPU host logic (frame sharding, reference-frame broadcast)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kvazaar_b200 import dist as kd


def test_shard_frames_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in kd.shard_frames(37, r, world))
        assert seen == list(range(37))
        assert all(kd.owner_of(i, world) == r for r in range(world) for i in kd.shard_frames(37, r, world))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames = 5
        mine = kd.shard_frames(frames, rank, world)
        ok = True
        for f in range(frames):
            buf = torch.full((64 * 64 * 3 // 2,), f + 1, dtype=torch.uint8) if f in mine else torch.zeros(64 * 64 * 3 // 2, dtype=torch.uint8)
            kd.broadcast_reference_frame(buf, f, world)
            ok &= bool((buf == f + 1).all())
        sizes = kd.gather_result_sizes(100 + rank)
        ok &= sizes == [100 + r for r in range(world)]
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_broadcast_reference_frame_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}
    assert np.all([out[0], out[1]])


def test_tile_grid_matches_reference_rule():
    """Uniform tile split (encoder.c:383-391): boundary i = i * size_in_ctus / count, in CTUs."""
    xs, ys = kd.tile_grid(7680, 4320, 4, 2)
    assert xs == [0, 1920, 3840, 5760, 7680] and ys == [0, 2176, 4320]          # 120 x 68 CTUs -> 30-CTU columns, 34-CTU rows
    xs, ys = kd.tile_grid(1920, 1080, 4, 2)
    assert xs == [0, 448, 960, 1408, 1920] and ys == [0, 512, 1080]              # 30 x 17 CTUs -> 7,8,7,8 columns; 8,9 rows
    assert [kd.tile_of_rank(r, 4, 2) for r in range(8)] == [(c, r) for r in range(2) for c in range(4)]


def _tile_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, H, cols, rows = 256, 192, 2, 1
        rng = np.random.default_rng(3)
        full = torch.from_numpy(rng.integers(0, 256, W * H * 3 // 2, dtype=np.uint8))
        xs, ys = kd.tile_grid(W, H, cols, rows)
        # every rank starts with only its own tile filled in
        mine = torch.zeros_like(full)
        tx, ty = kd.tile_of_rank(rank, cols, rows)
        for off, pw, ph, sub in ((0, W, H, 1), (W * H, W // 2, H // 2, 2), (W * H * 5 // 4, W // 2, H // 2, 2)):
            x0, x1, y0, y1 = xs[tx] // sub, xs[tx + 1] // sub, ys[ty] // sub, ys[ty + 1] // sub
            mine[off:off + pw * ph].view(ph, pw)[y0:y1, x0:x1] = full[off:off + pw * ph].view(ph, pw)[y0:y1, x0:x1]
        kd.allgather_tile_reconstructions(mine, W, H, cols, rows)
        out[rank] = bool(torch.equal(mine, full))
    finally:
        dist.destroy_process_group()


def test_allgather_tile_reconstructions_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_tile_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}
