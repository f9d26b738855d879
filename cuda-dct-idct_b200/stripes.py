"""Block-row striping of an image over ranks (one process per GPU).

Every 8x8 block is independent (the reference's host function
already accepts rectangular H x W, main_newAppr.cu:261-262), so GPU g simply owns a
contiguous range of block-rows: no halo, no exchange, no collective on the data path.
"""
from __future__ import annotations


def stripe_rows(H: int, world_size: int, rank: int) -> tuple[int, int]:
    """[row0, row1) of the image rows owned by `rank`; whole block-rows, sizes differ by
    at most one block-row, empty stripes allowed when there are fewer block-rows than ranks."""
    if H % 8:
        raise ValueError("H must be a multiple of 8")
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    nb = H // 8
    base, extra = divmod(nb, world_size)
    b0 = rank * base + min(rank, extra)
    b1 = b0 + base + (1 if rank < extra else 0)
    return b0 * 8, b1 * 8


def batch_images(n_images: int, world_size: int, rank: int) -> list[int]:
    """Indices of the images of a batch that `rank` processes: image b goes to rank b mod world_size
    (BASELINE configs[4] "a batch of 64 8192^2 images"); every image exactly once,
    counts differ by at most one, empty shares allowed."""
    if n_images < 0:
        raise ValueError("n_images must not be negative")
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_images, world_size))
