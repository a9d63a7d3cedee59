"""Device-resident entry points (kmg_dev_* / kmg_job_* of include/kmeans_gpu.h) over torch CUDA
tensors.  PyTorch is plumbing here — device memory and streams — all arithmetic is in
libkmeans_gpu.so.  Used by the stage-level parity tests, bench.py and batch callers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native
from .processor import ColorSpace, ImageProcessor, Opts, ReduceMode


def _dptr(t: torch.Tensor) -> C.c_void_p:
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("expected a contiguous CUDA tensor")
    return C.c_void_p(t.data_ptr())


def _stream_ptr(stream) -> C.c_void_p:
    if stream is None:
        stream = torch.cuda.current_stream()
    return C.c_void_p(stream.cuda_stream)


def _f32p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def convert(proc: ImageProcessor, rgba: torch.Tensor, color_space: ColorSpace = ColorSpace.Lab, stream=None) -> torch.Tensor:
    """K1/K3: (.., 4) uint8 -> (n, 4) float32 work plane."""
    n = rgba.numel() // 4
    work = torch.empty((n, 4), dtype=torch.float32, device=rgba.device)
    _native.check(proc._lib.kmg_dev_convert(proc.ctx, _dptr(rgba), n, int(color_space), _dptr(work), _stream_ptr(stream)))
    return work


def resize(proc: ImageProcessor, rgba: torch.Tensor, dw: int, dh: int, stream=None) -> torch.Tensor:
    """K15: (h, w, 4) uint8 -> (dh, dw, 4) uint8."""
    sh, sw = rgba.shape[:2]
    out = torch.empty((dh, dw, 4), dtype=torch.uint8, device=rgba.device)
    _native.check(proc._lib.kmg_dev_resize(proc.ctx, _dptr(rgba), sw, sh, _dptr(out), dw, dh, _stream_ptr(stream)))
    return out


def assign(proc: ImageProcessor, work: torch.Tensor, centroids: np.ndarray, stream=None) -> torch.Tensor:
    """K5: labels (uint32 stored in an int32 tensor) for a work plane."""
    n = work.shape[0]
    cent = np.ascontiguousarray(centroids, np.float32).reshape(-1, 4)
    labels = torch.empty(n, dtype=torch.int32, device=work.device)
    _native.check(proc._lib.kmg_dev_assign(proc.ctx, _dptr(work), n, _f32p(cent), cent.shape[0], _dptr(labels),
                                           _stream_ptr(stream)))
    return labels


def remap(proc: ImageProcessor, rgba: torch.Tensor, centroids: np.ndarray, mode: ReduceMode = ReduceMode.Replace,
          color_space: ColorSpace = ColorSpace.Lab, out: torch.Tensor | None = None, stream=None) -> torch.Tensor:
    h, w = rgba.shape[:2]
    cent = np.ascontiguousarray(centroids, np.float32).reshape(-1, 4)
    if out is None:
        out = torch.empty_like(rgba)
    _native.check(proc._lib.kmg_dev_remap(proc.ctx, _dptr(rgba), w, h, _f32p(cent), cent.shape[0], int(color_space),
                                          int(mode), _dptr(out), _stream_ptr(stream)))
    return out


def synth(proc: ImageProcessor, n: int, first_pixel: int = 0, frame: int = 0, seed: int = 0, blobs: int = 0,
          device=None, out: torch.Tensor | None = None, stream=None) -> torch.Tensor:
    if out is None:
        out = torch.empty((n, 4), dtype=torch.uint8, device=device or f"cuda:{proc.device}")
    _native.check(proc._lib.kmg_dev_synth(proc.ctx, _dptr(out), first_pixel, n, frame, seed, blobs, _stream_ptr(stream)))
    return out


def srgb_table(proc: ImageProcessor) -> np.ndarray:
    t = np.empty(256, np.float32)
    _native.check(proc._lib.kmg_dev_srgb_table(proc.ctx, _f32p(t)))
    return t


def fast_lab_error(proc: ImageProcessor) -> float:
    v = C.c_float(0)
    _native.check(proc._lib.kmg_dev_fast_lab_error(proc.ctx, C.byref(v)))
    return float(v.value)


def audit(proc: ImageProcessor, centroids: np.ndarray, search: int, mode: int, work: torch.Tensor | None = None,
          rgba: torch.Tensor | None = None, w: int | None = None, h: int | None = None,
          color_space: ColorSpace = ColorSpace.Lab, stream=None) -> tuple[int, int]:
    """Certificate audit (kmg_dev_audit): returns (certified-but-wrong pixels, uncertified pixels)."""
    cent = np.ascontiguousarray(centroids, np.float32).reshape(-1, 4)
    n = work.shape[0] if work is not None else rgba.numel() // 4
    if w is None:
        w, h = n, 1
    wrong, unc = C.c_uint64(0), C.c_uint64(0)
    _native.check(proc._lib.kmg_dev_audit(proc.ctx, _dptr(work) if work is not None else None,
                                          _dptr(rgba) if rgba is not None else None, w, h, _f32p(cent), cent.shape[0],
                                          int(color_space), search, mode, C.byref(wrong), C.byref(unc), _stream_ptr(stream)))
    return int(wrong.value), int(unc.value)


def fp32_peak(proc: ImageProcessor) -> float:
    """Measured non-tensor FP32 peak of the device, fused multiply-adds per second."""
    v = C.c_double(0)
    _native.check(proc._lib.kmg_dev_fp32_peak(proc.ctx, C.byref(v)))
    return float(v.value)


class Job:
    """A k-means problem resident on the device (kmg_job)."""

    def __init__(self, proc: ImageProcessor, work: torch.Tensor, w: int, h: int, k: int,
                 color_space: ColorSpace = ColorSpace.Lab, opts: Opts | None = None):
        self.proc = proc
        self.work = work  # keep alive: the job borrows the plane
        self.k = k
        self.w, self.h = w, h
        handle = C.c_void_p()
        o = (opts or Opts()).to_c()
        _native.check(proc._lib.kmg_job_create(proc.ctx, _dptr(work), w, h, k, int(color_space), C.byref(o), C.byref(handle)))
        self._job = handle

    def close(self):
        if getattr(self, "_job", None):
            self.proc._lib.kmg_job_destroy(self._job)
            self._job = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_shard(self, global_w: int, global_h: int, row_offset: int):
        _native.check(self.proc._lib.kmg_job_set_shard(self._job, global_w, global_h, row_offset))

    def init(self, stream=None):
        idx = np.zeros(self.k, np.uint32)
        dist = np.zeros(self.k, np.float32)
        _native.check(self.proc._lib.kmg_job_init(self._job, idx.ctypes.data_as(C.POINTER(C.c_uint32)), _f32p(dist),
                                                  _stream_ptr(stream)))
        return idx, dist

    def set_centroids(self, centroids: np.ndarray, stream=None):
        cent = np.ascontiguousarray(centroids, np.float32).reshape(-1, 4)
        assert cent.shape[0] == self.k
        _native.check(self.proc._lib.kmg_job_set_centroids(self._job, _f32p(cent), _stream_ptr(stream)))

    def centroids(self, stream=None) -> np.ndarray:
        cent = np.empty((self.k, 4), np.float32)
        _native.check(self.proc._lib.kmg_job_get_centroids(self._job, _f32p(cent), _stream_ptr(stream)))
        return cent

    def step(self, count: int = 1, stream=None):
        _native.check(self.proc._lib.kmg_job_step(self._job, count, _stream_ptr(stream)))

    def run(self, stream=None) -> int:
        passes = C.c_uint32(0)
        _native.check(self.proc._lib.kmg_job_run(self._job, C.byref(passes), _stream_ptr(stream)))
        return passes.value

    def stats(self, stream=None):
        conv, passes, slow = C.c_uint32(0), C.c_uint32(0), C.c_uint64(0)
        _native.check(self.proc._lib.kmg_job_stats(self._job, C.byref(conv), C.byref(passes), C.byref(slow),
                                                   _stream_ptr(stream)))
        return {"converged": conv.value, "passes": passes.value, "slow_pixels": slow.value}

    def init_stats(self, stream=None):
        """Work of the lazy farthest-point rounds: sweeps, refreshed pixels, exact distances evaluated."""
        sw, re, fo, ex = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _native.check(self.proc._lib.kmg_job_init_stats(self._job, C.byref(sw), C.byref(re), C.byref(fo), C.byref(ex),
                                                        _stream_ptr(stream)))
        return {"sweeps": sw.value, "refreshed": re.value, "pairs": fo.value, "exact": ex.value}

    def sums(self, stream=None) -> np.ndarray:
        """k x (sum0, sum1, sum2, count) of the last pass, sums in units of 2^-15."""
        acc = np.zeros((self.k, 4), np.int64)
        _native.check(self.proc._lib.kmg_job_get_sums(self._job, acc.ctypes.data_as(C.POINTER(C.c_int64)),
                                                      _stream_ptr(stream)))
        return acc

    def remap(self, rgba: torch.Tensor, mode: ReduceMode = ReduceMode.Replace, out: torch.Tensor | None = None,
              stream=None) -> torch.Tensor:
        h, w = rgba.shape[:2]
        if out is None:
            out = torch.empty_like(rgba)
        _native.check(self.proc._lib.kmg_dev_remap_job(self.proc.ctx, _dptr(rgba), w, h, self._job, int(mode),
                                                       _dptr(out), _stream_ptr(stream)))
        return out


def reduce_batch(proc: ImageProcessor, frames: torch.Tensor, k: int, mode: ReduceMode = ReduceMode.Replace,
                 color_space: ColorSpace = ColorSpace.Lab, opts: Opts | None = None, out: torch.Tensor | None = None,
                 stream=None):
    """frames: (n, h, w, 4) uint8 on the device."""
    n, h, w, _ = frames.shape
    if out is None:
        out = torch.empty_like(frames)
    cent = np.empty((n, k, 4), np.float32)
    passes = np.zeros(n, np.uint32)
    o = (opts or Opts()).to_c()
    _native.check(proc._lib.kmg_dev_reduce_batch(proc.ctx, _dptr(frames), n, w, h, k, int(color_space), int(mode),
                                                 C.byref(o), _dptr(out), _f32p(cent),
                                                 passes.ctypes.data_as(C.POINTER(C.c_uint32)), _stream_ptr(stream)))
    return out, cent, passes


def comm_unique_id(proc: ImageProcessor) -> bytes:
    buf = (C.c_uint8 * 128)()
    _native.check(proc._lib.kmg_comm_unique_id(proc.ctx, buf))
    return bytes(buf)


def comm_init(proc: ImageProcessor, uid: bytes, n_ranks: int, rank: int):
    buf = (C.c_uint8 * 128).from_buffer_copy(uid)
    _native.check(proc._lib.kmg_comm_init(proc.ctx, buf, n_ranks, rank))


def comm_mode(proc: ImageProcessor) -> int:
    """0: none, 1: NCCL all-reduce per pass, 2: in-kernel exchange through peer-mapped mailboxes."""
    return int(proc._lib.kmg_comm_mode(proc.ctx))


def comm_destroy(proc: ImageProcessor):
    _native.check(proc._lib.kmg_comm_destroy(proc.ctx))
