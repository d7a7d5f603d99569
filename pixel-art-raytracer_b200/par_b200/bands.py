"""Row-band partition of a frame over ranks and the in-place frame gather.

One process per GPU: rank r renders rows [r*H/N, (r+1)*H/N) in place into a full-size frame
and a single in-place all-gather completes the frame on every rank.  Backend agnostic
(nccl on GPUs over NVLink, gloo on CPU for the host-logic tests)."""
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
