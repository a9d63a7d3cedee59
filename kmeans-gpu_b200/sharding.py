"""Host-side partitioning rules for the two multi-GPU modes."""
from __future__ import annotations


def row_shards(height: int, n_ranks: int) -> list[tuple[int, int]]:
    """Contiguous row blocks of one image: rank r owns rows [r*h/G, (r+1)*h/G)."""
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    return [((r * height) // n_ranks, ((r + 1) * height) // n_ranks) for r in range(n_ranks)]


def frame_shards(n_frames: int, n_ranks: int) -> list[tuple[int, int]]:
    """Contiguous frame ranges of a batch; no collective is needed between them."""
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    return [((r * n_frames) // n_ranks, ((r + 1) * n_frames) // n_ranks) for r in range(n_ranks)]
