"""Host-side mirror of the reference's public API (core/src/lib.rs, core/src/image.rs).

Same names, argument meaning and error behaviour as the Rust crate: `ImageProcessor.palette`,
`.find`, `.reduce`, enums `ColorSpace`, `Algorithm`, `ReduceMode`, and `Image`.  Images are numpy
uint8 arrays of shape (h, w, 4) (tightly packed RGBA8, as `&[RGBA8]` in the reference).
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass

import numpy as np

from . import _native
from ._native import KmgError, KmgOpts


class ColorSpace(enum.IntEnum):
    """core/src/lib.rs:168-213"""

    Lab = 0
    Rgb = 1

    @staticmethod
    def from_str(s: str) -> "ColorSpace":
        if s == "lab":
            return ColorSpace.Lab
        if s == "rgb":
            return ColorSpace.Rgb
        raise ValueError(f"Unsupported color space {s}")

    @property
    def label(self) -> str:
        return "lab" if self is ColorSpace.Lab else "rgb"

    def convergence(self) -> float:
        return 1.0 if self is ColorSpace.Lab else 0.01

    def __str__(self) -> str:
        return self.label


class Algorithm(enum.Enum):
    """core/src/lib.rs:216-232"""

    Kmeans = "kmeans"
    Octree = "octree"

    def __str__(self) -> str:
        return self.value


class ReduceMode(enum.IntEnum):
    """core/src/lib.rs:235-253"""

    Replace = 0
    Dither = 1
    Meld = 2

    def __str__(self) -> str:
        return self.name.lower()


@dataclass
class Opts:
    """The reference's hard-coded constants (kmg_opts in include/kmeans_gpu.h)."""

    max_dim: int = 256
    max_iter: int = 128
    check_every: int = 8
    convergence: float = -1.0
    seed_x_frac: float = 0.5625
    seed_y_frac: float = 0.93359375
    seed_x: int = -1
    seed_y: int = -1
    fused_kmeans: bool = True  # False: KMG_OPT_NO_FUSED_KMEANS (stage-by-stage launches; same results)

    def to_c(self) -> KmgOpts:
        return KmgOpts(C.sizeof(KmgOpts), self.max_dim, self.max_iter, self.check_every, self.convergence,
                       self.seed_x_frac, self.seed_y_frac, self.seed_x, self.seed_y,
                       0 if self.fused_kmeans else _native.KMG_OPT_NO_FUSED_KMEANS)


class Image:
    """core/src/image.rs:20-48 — (width, height) + RGBA8 pixels."""

    def __init__(self, dimensions, rgba):
        w, h = int(dimensions[0]), int(dimensions[1])
        arr = np.ascontiguousarray(rgba, dtype=np.uint8)
        if arr.size != w * h * 4:
            raise ValueError(f"pixel buffer has {arr.size} bytes, expected {w * h * 4}")
        self.dimensions = (w, h)
        self.rgba = arr.reshape(h, w, 4)

    @staticmethod
    def new(dimensions, rgba) -> "Image":
        return Image(dimensions, rgba)

    @staticmethod
    def from_array(arr: np.ndarray) -> "Image":
        arr = np.asarray(arr)
        if arr.ndim != 3 or arr.shape[2] != 4:
            raise ValueError("expected an (h, w, 4) uint8 array")
        return Image((arr.shape[1], arr.shape[0]), arr)

    def get_pixel(self, x: int, y: int):
        return tuple(int(v) for v in self.rgba[y, x])

    def into_raw_pixels(self) -> np.ndarray:
        return self.rgba.reshape(-1, 4)


def _as_image(image) -> Image:
    return image if isinstance(image, Image) else Image.from_array(image)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def fixed_centroids(colors_rgba8, color_space: ColorSpace = ColorSpace.Lab) -> np.ndarray:
    """CentroidsBuffer::fixed_centroids (core/src/structures.rs:523-553)."""
    cols = np.ascontiguousarray(colors_rgba8, dtype=np.uint8).reshape(-1, 4)
    out = np.empty((cols.shape[0], 4), np.float32)
    _native.load().kmg_fixed_centroids(cols.ctypes.data_as(C.POINTER(C.c_uint8)), cols.shape[0], int(color_space),
                                       out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def centroids_to_rgba8(centroids, color_space: ColorSpace = ColorSpace.Lab) -> np.ndarray:
    """CentroidsBuffer::pull_values (core/src/structures.rs:600-617)."""
    cent = np.ascontiguousarray(centroids, dtype=np.float32).reshape(-1, 4)
    out = np.empty((cent.shape[0], 4), np.uint8)
    _native.load().kmg_centroids_to_rgba8(cent.ctypes.data_as(C.POINTER(C.c_float)), cent.shape[0], int(color_space),
                                          out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def sort_palette_by_lightness(colors_rgba8) -> np.ndarray:
    """core/src/lib.rs:276-284."""
    cols = np.ascontiguousarray(colors_rgba8, dtype=np.uint8).reshape(-1, 4).copy()
    _native.load().kmg_sort_palette_by_lightness(cols.ctypes.data_as(C.POINTER(C.c_uint8)), cols.shape[0])
    return cols


def octree_colors(pixels_rgba8, color_count: int) -> np.ndarray:
    """operations::extract_palette_octree (core/src/operations.rs:90-97): the reference's CPU octree
    quantiser over the given pixels; <= color_count colours sorted as (r,g,b,a) tuples."""
    px = np.ascontiguousarray(pixels_rgba8, dtype=np.uint8).reshape(-1, 4)
    out = np.empty((max(int(color_count), 1), 4), np.uint8)
    n = C.c_uint32(0)
    code = _native.load().kmg_octree_palette(px.ctypes.data_as(C.POINTER(C.c_uint8)), px.shape[0], int(color_count),
                                             out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(n))
    if code != 0:
        raise KmgError(code, "kmg_octree_palette: bad argument")
    return out[: n.value].copy()


def resized_dims(w: int, h: int, max_size: int = 256):
    ow, oh = C.c_uint32(), C.c_uint32()
    _native.load().kmg_resized_dims(w, h, max_size, C.byref(ow), C.byref(oh))
    return ow.value, oh.value


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """A numpy array in page-locked host memory (kmg_alloc_pinned): host<->device copies of such
    buffers run at full PCIe speed and overlap with kernels.  Freed when the array is collected."""
    import weakref

    lib = _native.load()
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    ptr = lib.kmg_alloc_pinned(max(n, 1))
    if not ptr:
        raise KmgError(3, lib.kmg_last_error().decode("utf-8", "replace"))
    buf = (C.c_uint8 * max(n, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, lib.kmg_free_pinned, ptr)
    return arr


def _out_image(out, h, w):
    if out is None:
        return np.empty((h, w, 4), np.uint8)
    if out.dtype != np.uint8 or out.shape != (h, w, 4) or not out.flags["C_CONTIGUOUS"]:
        raise ValueError(f"out must be a contiguous uint8 array of shape {(h, w, 4)}")
    return out


class ImageProcessor:
    """core/src/lib.rs:24-165.  One instance may be shared by many threads
    (core/examples/parallel.rs:23,36-51)."""

    def __init__(self, device: int = 0):
        self._lib = _native.load()
        handle = C.c_void_p()
        _native.check(self._lib.kmg_create(device, C.byref(handle)))
        self._ctx = handle
        self.device = device

    @staticmethod
    def new(device: int = 0) -> "ImageProcessor":
        return ImageProcessor(device)

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._lib.kmg_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def ctx(self):
        return self._ctx

    def launch_count(self) -> int:
        return int(self._lib.kmg_launch_count(self._ctx))

    # -- the three reference operations -------------------------------------------------------

    def palette(self, color_count: int, image, algo: Algorithm = Algorithm.Kmeans, opts: Opts | None = None,
                color_space: ColorSpace = ColorSpace.Lab) -> np.ndarray:
        """lib.rs:67-77 + kmeans_palette :255-286: k colours (RGBA8, alpha 255) sorted by Lab L."""
        if algo is Algorithm.Octree:
            return self.octree_palette(color_count, image)
        cent, _ = self.kmeans_centroids(color_count, image, color_space, opts)
        return sort_palette_by_lightness(centroids_to_rgba8(cent, color_space))

    def octree_palette(self, color_count: int, image) -> np.ndarray:
        """octree_palette (core/src/lib.rs:288-331): shrink to <= 128 px on the device
        (InputTexture::resized + pull_image), quantise on the CPU as the reference does
        (core/src/octree.rs), sort by Lab L."""
        img = _as_image(image)
        w, h = img.dimensions
        pixels = self.resize(img, 128).rgba if (w > 128 or h > 128) else img.rgba
        return sort_palette_by_lightness(octree_colors(pixels, color_count))

    def find(self, image, colors, reduce_mode: ReduceMode = ReduceMode.Replace,
             color_space: ColorSpace = ColorSpace.Lab, out: np.ndarray | None = None) -> Image:
        """lib.rs:79-114: remap onto a fixed palette given as RGBA8 colours."""
        cent = fixed_centroids(colors, color_space)
        return self.remap(image, cent, reduce_mode, color_space, out=out)

    def reduce(self, color_count: int, image, algo: Algorithm = Algorithm.Kmeans,
               reduce_mode: ReduceMode = ReduceMode.Replace, opts: Opts | None = None,
               color_space: ColorSpace = ColorSpace.Lab, return_details: bool = False,
               out: np.ndarray | None = None):
        """lib.rs:116-164.  `out`: optional (h, w, 4) uint8 result buffer (e.g. from pinned_empty)."""
        if algo is Algorithm.Octree:
            # lib.rs:133-136: the octree palette goes through fixed_centroids into the same remap
            colors = self.octree_palette(color_count, image)
            res = self.find(image, colors, reduce_mode, color_space, out=out)
            if return_details:
                return res, fixed_centroids(colors, color_space), 0
            return res
        img = _as_image(image)
        w, h = img.dimensions
        out = _out_image(out, h, w)
        cent = np.empty((int(color_count), 4), np.float32) if color_count > 0 else np.empty((0, 4), np.float32)
        passes = C.c_uint32(0)
        o = (opts or Opts()).to_c()
        _native.check(self._lib.kmg_reduce(self._ctx, _ptr(img.rgba), w, h, int(color_count), int(color_space),
                                           int(reduce_mode), C.byref(o), _ptr(out),
                                           cent.ctypes.data_as(C.POINTER(C.c_float)), C.byref(passes)))
        res = Image((w, h), out)
        return (res, cent, passes.value) if return_details else res

    # -- pieces of the boundary exposed for callers that keep centroids ------------------------

    def kmeans_centroids(self, color_count: int, image, color_space: ColorSpace = ColorSpace.Lab,
                         opts: Opts | None = None):
        """operations::extract_palette_kmeans (operations.rs:15-88): raw centroids + pass count."""
        img = _as_image(image)
        w, h = img.dimensions
        cent = np.empty((max(int(color_count), 0), 4), np.float32)
        passes = C.c_uint32(0)
        o = (opts or Opts()).to_c()
        _native.check(self._lib.kmg_kmeans_palette(self._ctx, _ptr(img.rgba), w, h, int(color_count),
                                                   int(color_space), C.byref(o),
                                                   cent.ctypes.data_as(C.POINTER(C.c_float)), C.byref(passes)))
        return cent, passes.value

    def remap(self, image, centroids, reduce_mode: ReduceMode = ReduceMode.Replace,
              color_space: ColorSpace = ColorSpace.Lab, out: np.ndarray | None = None) -> Image:
        """operations::{find_colors,dither_colors,meld_colors} (operations.rs:99-271)."""
        img = _as_image(image)
        w, h = img.dimensions
        cent = np.ascontiguousarray(centroids, dtype=np.float32).reshape(-1, 4)
        out = _out_image(out, h, w)
        _native.check(self._lib.kmg_remap(self._ctx, _ptr(img.rgba), w, h, cent.ctypes.data_as(C.POINTER(C.c_float)),
                                          cent.shape[0], int(color_space), int(reduce_mode), _ptr(out)))
        return Image((w, h), out)

    def resize(self, image, max_size: int) -> Image:
        """InputTexture::resized (structures.rs:76-182) + pull_image."""
        img = _as_image(image)
        w, h = img.dimensions
        ow, oh = resized_dims(w, h, max_size)
        out = np.empty((oh, ow, 4), np.uint8)
        _native.check(self._lib.kmg_resize(self._ctx, _ptr(img.rgba), w, h, max_size, _ptr(out)))
        return Image((ow, oh), out)

    def reduce_batch(self, color_count: int, frames: np.ndarray, reduce_mode: ReduceMode = ReduceMode.Replace,
                     opts: Opts | None = None, color_space: ColorSpace = ColorSpace.Lab,
                     out: np.ndarray | None = None):
        """Batch of equally sized frames, array (n, h, w, 4).  Chunks of frames are pipelined
        (upload / kernels / read-back overlap) when `frames` and `out` are page-locked (pinned_empty)."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        n, h, w, _ = frames.shape
        if out is None:
            out = np.empty_like(frames)
        elif out.dtype != np.uint8 or out.shape != frames.shape or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a contiguous uint8 array shaped like frames")
        cent = np.empty((n, int(color_count), 4), np.float32)
        passes = np.zeros(n, np.uint32)
        o = (opts or Opts()).to_c()
        _native.check(self._lib.kmg_reduce_batch(self._ctx, _ptr(frames), n, w, h, int(color_count), int(color_space),
                                                 int(reduce_mode), C.byref(o), _ptr(out),
                                                 cent.ctypes.data_as(C.POINTER(C.c_float)),
                                                 passes.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out, cent, passes
