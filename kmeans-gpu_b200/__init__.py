"""kmeans_gpu_b200 — B200-native (sm_100a CUDA) image hot path of redwarp/kmeans-gpu.

Host-side mirror of the reference's public interface (core/src/lib.rs) over the C ABI of
libkmeans_gpu.so (include/kmeans_gpu.h).  PyTorch is only used by `device` helpers for device
memory and streams; all arithmetic is in the hand-written CUDA library.
"""
from .processor import (  # noqa: F401
    Algorithm,
    ColorSpace,
    Image,
    ImageProcessor,
    KmgError,
    Opts,
    ReduceMode,
    fixed_centroids,
    centroids_to_rgba8,
    sort_palette_by_lightness,
    resized_dims,
    pinned_empty,
)
from .palette import parse_colors, parse_palette, validate_palette  # noqa: F401
from .sharding import row_shards, frame_shards  # noqa: F401

__all__ = [
    "Algorithm", "ColorSpace", "Image", "ImageProcessor", "KmgError", "Opts", "ReduceMode",
    "fixed_centroids", "centroids_to_rgba8", "sort_palette_by_lightness", "resized_dims", "pinned_empty",
    "parse_colors", "parse_palette", "validate_palette", "row_shards", "frame_shards",
]
