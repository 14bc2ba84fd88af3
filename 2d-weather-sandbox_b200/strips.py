"""x-strip domain decomposition plan.

The grid is cut into N vertical strips, one per GPU / process; rank r owns global columns
[r*W/N, (r+1)*W/N).  x is periodic (textures use REPEAT wrap, app.js:5191-5230), so the ranks form
a ring: the left neighbour of rank 0 is rank N-1.  Every strip keeps the full height, so the
periodic y-wrap stays local.

Each rank stores its strip with GHOST columns on both sides.  One iteration of the fused
schedule has a dependency radius of 3 (pressure+velocity+curl+vorticity+boundary kernel: its
output is valid from the 4th ghost column inwards) + 1 + ceil(|v|max) (the bilinear footprint of
the advection back-trace; the sun-ray fetch of the lighting pass needs 2), so with GHOST = 8 a
single exchange of GHOST columns per iteration keeps every owned cell bit-identical to the
single-GPU run while |v| <= MAX_STRIP_VELOCITY = 4 cells/iteration (the shipped saves peak at
0.36).  libwsb200 uses exactly these numbers (csrc/wsb200.cu: kGhost, kMaxStripVelocity) and
every synchronising call of a strip fails once the running maximum of |v| over its own columns
exceeds the budget.
"""
from __future__ import annotations

GHOST = 8
MAX_STRIP_VELOCITY = 4.0


def strip_bounds(width: int, n_ranks: int, rank: int) -> tuple[int, int]:
    """(x_begin, local_width) of `rank`."""
    if not (0 <= rank < n_ranks):
        raise ValueError(f"rank {rank} outside 0..{n_ranks - 1}")
    x0 = (rank * width) // n_ranks
    x1 = ((rank + 1) * width) // n_ranks
    if n_ranks > 1 and x1 - x0 < 2 * GHOST:
        raise ValueError(f"strip of {x1 - x0} columns is narrower than 2*GHOST={2 * GHOST}")
    return x0, x1 - x0


def neighbours(rank: int, n_ranks: int) -> tuple[int, int]:
    """(left, right) ranks on the periodic ring."""
    return (rank - 1) % n_ranks, (rank + 1) % n_ranks


def padded_columns(width: int, n_ranks: int, rank: int) -> list[int]:
    """Global column index of every column of the rank's padded array (ghost | owned | ghost)."""
    x0, wl = strip_bounds(width, n_ranks, rank)
    g = GHOST if n_ranks > 1 else 0
    return [(x0 - g + i) % width for i in range(wl + 2 * g)]


def send_slices(local_width: int) -> dict[str, slice]:
    """Column slices (in padded-array coordinates) a rank SENDS: its leftmost owned GHOST columns
    go to the left neighbour's right ghost zone, its rightmost owned columns to the right
    neighbour's left ghost zone."""
    g = GHOST
    return {"to_left": slice(g, 2 * g), "to_right": slice(local_width, local_width + g)}


def recv_slices(local_width: int) -> dict[str, slice]:
    g = GHOST
    return {"from_left": slice(0, g), "from_right": slice(local_width + g, local_width + 2 * g)}
