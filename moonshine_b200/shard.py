"""Pixel-tile data parallelism: the image is cut into 32x32 tiles dealt round-robin to ranks.

Every pixel's estimate depends only on (sample, x, y) seeds and read-only scene data
(shaders/hrtsystem/main.hlsl:85), so ranks never exchange anything on the data path. Each rank renders ALL
samples of its own tiles in sample order (keeps the FP32 running mean of main.hlsl:43-51 bit-identical to
the 1-GPU result) into a full-frame RGBA32F buffer that is zero outside its tiles; one reduce(sum) to rank 0
assembles the frame exactly (x + 0 = x).

The ownership rule lives in the library (runtime.cpp `buildPixIds`, driven by HdMoonshineExtSetTileShard);
this module mirrors it for host-side logic/tests and wraps the collective.
"""
from __future__ import annotations

import numpy as np

TILE = 32


def tile_grid(width: int, height: int):
    return (width + TILE - 1) // TILE, (height + TILE - 1) // TILE


def owner_map(width: int, height: int, world: int) -> np.ndarray:
    """(H, W) int32 array: rank that owns each pixel (tile index t = ty*tiles_x + tx; owner = t % world)."""
    tx, _ = tile_grid(width, height)
    y, x = np.mgrid[0:height, 0:width]
    return (((y // TILE) * tx + (x // TILE)) % world).astype(np.int32)


def owned_pixel_count(width: int, height: int, rank: int, world: int) -> int:
    return int((owner_map(width, height, world) == rank).sum())


def reduce_radiance(buf, dst: int = 0, group=None):
    """Sum the per-rank full-frame radiance buffers onto `dst` (NCCL for CUDA tensors, gloo for CPU tensors)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return buf


def allreduce_counts(counts, group=None):
    """Sum ray counters over ranks (u64 carried as int64)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts
