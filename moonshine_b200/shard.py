"""Pixel-tile data parallelism: the image is cut into 16x16 tiles dealt to the ranks, one rank per GPU.

Every pixel's estimate depends only on (sample, x, y) seeds and read-only scene data
(shaders/hrtsystem/main.hlsl:85), so ranks never exchange anything on the data path. Each rank renders ALL
samples of its own tiles in sample order (keeps the FP32 running mean of main.hlsl:43-51 bit-identical to
the 1-GPU result) into a full-frame RGBA32F buffer that is zero outside its tiles; one reduce(sum) to rank 0
assembles the frame exactly (x + 0 = x).

The ownership rule AND the collective live in the library (runtime.cpp `tileOwner` / `buildPixIds`, `reduceSensor`:
HdMoonshineExtSetTileShard, HdMoonshineExtCommInit, HdMoonshineExtCommReduceSensor over NCCL). This module is the host-side
mirror of the rule (tests, tooling), the rendezvous that brings a communicator up on the ranks of one node, and a
torch.distributed twin of the reduce used by the CPU-only gloo test.
"""
from __future__ import annotations

import os

import numpy as np

TILE = 16      # runtime.cpp tileEdge


def stagger(world: int) -> int:
    """runtime.cpp tileStagger: every tile row starts this many ranks further - the smallest number >= 2 coprime to `world`."""
    import math
    if world == 2:
        return 3
    s = 2
    while world > 1 and math.gcd(s, world) != 1:
        s += 1
    return s


def tile_grid(width: int, height: int):
    return (width + TILE - 1) // TILE, (height + TILE - 1) // TILE


def owner_map(width: int, height: int, world: int) -> np.ndarray:
    """(H, W) int32 array: rank that owns each pixel; tile (cx, cy) belongs to rank (cx + stagger(world) * cy) % world."""
    y, x = np.mgrid[0:height, 0:width]
    return (((x // TILE) + stagger(world) * (y // TILE)) % world).astype(np.int32)


def owned_pixel_count(width: int, height: int, rank: int, world: int) -> int:
    return int((owner_map(width, height, world) == rank).sum())


def init_comm(m, rank: int = None, world: int = None, tag: str = None):
    """Bring up the library's NCCL communicator on this rank (env RANK / WORLD_SIZE / MASTER_PORT as set by torchrun or
    `offline --gpus N`). Returns (rank, world). A world of 1 needs no communicator."""
    from . import hydra
    rank = int(os.environ.get("RANK", 0)) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", 1)) if world is None else world
    if world > 1:
        tag = tag or os.environ.get("MSNE_JOB_TAG", os.environ.get("MASTER_PORT", "0"))
        m.CommInit(hydra.exchange_unique_id(rank, world, tag), rank, world)
        if rank == 0:
            hydra.retire_unique_id(tag)
    return rank, world


def reduce_radiance(buf, dst: int = 0, group=None):
    """torch.distributed twin of the library's reduceSensor (gloo for CPU tensors: the world_size-2 CPU test)."""
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
