"""N>1 host logic on CPU: tile ownership + the radiance reduce over a 2-rank gloo group (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from moonshine_b200 import scenes, shard


@pytest.mark.parametrize("w,h,world", [(96, 64, 2), (1920, 1080, 8), (100, 37, 3), (31, 31, 4)])
def test_every_pixel_has_exactly_one_owner(w, h, world):
    om = shard.owner_map(w, h, world)
    assert om.shape == (h, w) and om.min() >= 0 and om.max() < world
    counts = [shard.owned_pixel_count(w, h, r, world) for r in range(world)]
    assert sum(counts) == w * h
    T = shard.TILE
    if w * h >= T * T * world * 4:
        assert max(counts) - min(counts) <= 4 * T * T  # dealing tiles keeps ranks balanced to a few tiles
    # tiles are never split between ranks, and neither a tile row nor a tile column belongs to one rank
    assert (om[:T, :T] == om[0, 0]).all()
    if world > 1 and w >= 2 * T and h >= 2 * T:
        assert om[0, 0] != om[0, T] and om[0, 0] != om[T, 0]


def _worker(rank, world, port, w, h, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.binding import Oracle
        sc = scenes.single_triangle(extent=(w, h))
        o = Oracle()
        scenes.load_scene(o, sc)
        full, counts = o.render(sc.lens, w, h, n_launches=2, threads=1)
        mine = shard.owner_map(w, h, world) == rank
        part = np.where(mine[..., None], full, 0.0).astype(np.float32)  # what a rank's device buffer holds
        t = torch.from_numpy(part.copy())
        shard.reduce_radiance(t, dst=0)
        c = torch.tensor([int(mine.sum()), 1], dtype=torch.int64)
        shard.allreduce_counts(c)
        if rank == 0:
            np.save(out_path, np.concatenate([t.numpy().reshape(-1), full.reshape(-1), c.numpy().astype(np.float32)]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_reduce_reassembles_the_frame(tmp_path):
    w, h, world = 96, 64, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r.npy")
    mp.spawn(_worker, args=(world, port, w, h, out), nprocs=world, join=True)
    data = np.load(out)
    n = w * h * 4
    reduced, full, counts = data[:n], data[n:2 * n], data[2 * n:]
    np.testing.assert_array_equal(reduced.view(np.uint32), full.view(np.uint32))  # exact: one non-zero contributor per texel
    assert counts[0] == w * h and counts[1] == world
