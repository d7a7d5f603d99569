"""Partition of a frame over ranks and the in-place frame gather.

Two partitions, one process per GPU, backend agnostic (nccl over NVLink on GPUs, gloo on CPU
for the host-logic tests):
  * stripes (default for rendering): rank r owns the 40-row tile rows t with t % N == r and
    writes them stripe-major into a staging frame [N][T][40*W*4 bytes], T = ceil(H/40/N), so
    its output is contiguous; one in-place all-gather + an un-stripe copy give the raster frame.
    Interleaving balances the strongly row-dependent cost of the shadow walks.
  * bands: rank r owns rows [r*H/N, (r+1)*H/N) in place in the raster frame."""
from __future__ import annotations


def band_rows(H: int, world: int, rank: int) -> tuple[int, int]:
    """Rows [row0,row1) of rank `rank`; the last rank takes the remainder when world∤H."""
    if not (0 <= rank < world) or H < world:
        raise ValueError(f"bad band request H={H} world={world} rank={rank}")
    rows = H // world
    return rank * rows, (H if rank == world - 1 else (rank + 1) * rows)


def gather_bands(frame, W: int, H: int, world: int, rank: int, group=None) -> None:
    """frame: flat uint8 tensor of H*W*4 bytes whose band `rank` is valid; on return every
    rank holds the whole frame.  In place: band r lives at offset r of the output."""
    if world == 1:
        return
    import torch.distributed as dist
    row_bytes = W * 4
    if H % world == 0:
        r0, r1 = band_rows(H, world, rank)
        dist.all_gather_into_tensor(frame, frame[r0 * row_bytes:r1 * row_bytes], group=group)
    else:  # unequal bands: one broadcast per band
        for b in range(world):
            r0, r1 = band_rows(H, world, b)
            dist.broadcast(frame[r0 * row_bytes:r1 * row_bytes], src=b, group=group)


def stripes_per_rank(H: int, world: int) -> int:
    """T: 40-row stripes per rank in the stripe-major staging frame (the last ones may be padding)."""
    return (H // 40 + world - 1) // world


def owned_rows(H: int, world: int, rank: int) -> list[tuple[int, int]]:
    """Raster row ranges of the stripes rank `rank` renders."""
    return [(t * 40, t * 40 + 40) for t in range(rank, H // 40, world)]


def stripe_split_for(W: int, H: int, world: int) -> int:
    """par_config.stripe_split that gives every rank the same number of stripes: the smallest s in
    (1, 2, 4, 8) dividing W / 40 with (H / 40 * s) % world == 0 — e.g. 2 for the 108 tile rows of a
    7680x4320 frame on 8 GPUs; 1 (whole tile rows) when there is none or the rows already divide."""
    tile_rows, tile_cols = H // 40, W // 40
    if world <= 1 or tile_rows % world == 0:
        return 1
    for s in (2, 4, 8):
        if tile_cols % s == 0 and (tile_rows * s) % world == 0:
            return s
    return 1


def stripe_column_segment(v: int, world: int, split: int) -> int:
    """Column segment of stripe v within its tile row v // split: v % split rotated by one for every
    lcm(world, split) stripes, so that a rank's stripes visit all column segments in turn (v % world alone
    gives a rank the same columns in every tile row when world and split share a factor); par_device.cuh."""
    import math
    return (v % split + v // math.lcm(world, split)) % split if split > 1 else 0


def owned_rects(W: int, H: int, world: int, rank: int, split: int = 1) -> list[tuple[int, int, int, int]]:
    """(row0, row1, col0, col1) of the stripes rank `rank` renders with par_config.stripe_split = split."""
    split = max(split, 1)
    cols = (W // 40 // split) * 40
    out = []
    for v in range(rank, (H // 40) * split, world):
        t, seg = v // split, stripe_column_segment(v, world, split)
        out.append((t * 40, t * 40 + 40, seg * cols, (seg + 1) * cols if split > 1 else W))
    return out


def gather_stripes(staging, world: int, rank: int, group=None) -> None:
    """staging: flat uint8 tensor [world][T][40*W*4] whose block `rank` is valid; in-place all-gather."""
    if world == 1:
        return
    import torch.distributed as dist
    block = staging.numel() // world
    dist.all_gather_into_tensor(staging, staging[rank * block:(rank + 1) * block], group=group)


def unstripe(staging, W: int, H: int, world: int):
    """Stripe-major staging frame -> raster frame (flat uint8 tensor of H*W*4 bytes); torch ops only
    (the GPU path uses par_unstripe_device, this is the reference implementation for tests)."""
    T = stripes_per_rank(H, world)
    stripe = 40 * W * 4
    return staging.view(world, T, stripe).permute(1, 0, 2).reshape(-1)[: H * W * 4]
