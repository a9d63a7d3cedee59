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


NO_CANDIDATE = (1 << 64) - 1


def init_key(distance_bits: int, global_pixel: int) -> int:
    """Arg-max key of a farthest-point round (kmg_kernels.cuh: k_init_round): distance bits in the
    high word, (pixel ^ 15) in the low one — among equal distances the highest 16-pixel chunk wins
    and, inside it, the lowest pixel (plus_plus_init.wgsl:62-68,:92,:102,:136,:142)."""
    return (int(distance_bits) << 32) | ((int(global_pixel) & 0xFFFFFFFF) ^ 15)


def key_to_pixel(key: int) -> int:
    """A zero maximum selects pixel 0 (every thread starts from Candidate(0, 0.0))."""
    return 0 if (key >> 32) == 0 else ((key & 0xFFFFFFFF) ^ 15)


def local_init_candidate(key: int, first_pixel: int, n_pixels: int) -> tuple[int, int]:
    """What a rank posts to the peers' mailboxes in round j: its best key and the global pixel it
    resolves to — NO_CANDIDATE when a zero maximum resolves to pixel 0 and pixel 0 lives elsewhere."""
    p = key_to_pixel(key)
    return key, (p if first_pixel <= p < first_pixel + n_pixels else NO_CANDIDATE)


def merge_init_candidates(candidates: list[tuple[int, int]]) -> tuple[int, int]:
    """The rule every rank applies to the posted (key, pixel) pairs (k_init_round<., 2>): the largest
    key wins — keys of different shards never tie, they embed the global pixel index — and the
    colour comes from the rank whose candidate pixel is the winner's.  Returns (key, owner rank)."""
    kmax = max(k for k, _ in candidates)
    want = key_to_pixel(kmax)
    owners = [r for r, (_, p) in enumerate(candidates) if p == want]
    if len(owners) != 1:
        raise ValueError(f"winner pixel {want} is held by {len(owners)} ranks")
    return kmax, owners[0]
