"""ctypes binding of the CPU oracle (oracle/oracle.cpp).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_SO = ORACLE_DIR / "_build" / "liboracle.so"

LAB, RGB = 0, 1


class OrcOpts(C.Structure):
    _fields_ = [
        ("max_dim", C.c_uint32),
        ("max_iter", C.c_uint32),
        ("check_every", C.c_uint32),
        ("convergence", C.c_float),
        ("seed_x_frac", C.c_float),
        ("seed_y_frac", C.c_float),
        ("seed_x", C.c_int32),
        ("seed_y", C.c_int32),
        ("sum_mode", C.c_int32),
    ]


def default_opts(**kw) -> OrcOpts:
    o = OrcOpts(256, 128, 8, -1.0, 0.5625, 0.93359375, -1, -1, 1)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def build(force: bool = False) -> Path:
    src = ORACLE_DIR / "oracle.cpp"
    if force or not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(ORACLE_DIR)], stdout=subprocess.DEVNULL)
    return ORACLE_SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(ORACLE_SO))
        u8p, f32p, u32p, i64p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_float, C.c_uint32, C.c_int64, C.c_uint64))
        L.orc_num_threads.restype = C.c_int
        L.orc_pow_f32.restype = C.c_float
        L.orc_pow_f32.argtypes = [C.c_float, C.c_float]
        L.orc_cie94.restype = C.c_float
        L.orc_cie94.argtypes = [f32p, f32p]
        L.orc_srgb_decode.restype = C.c_float
        L.orc_srgb_decode.argtypes = [C.c_uint8]
        L.orc_convert.argtypes = [u8p, C.c_size_t, C.c_int, f32p]
        L.orc_revert.argtypes = [f32p, C.c_size_t, C.c_int, u8p]
        L.orc_assign.argtypes = [f32p, C.c_size_t, f32p, C.c_uint32, u32p, f32p, f32p]
        L.orc_update.restype = C.c_uint32
        L.orc_update.argtypes = [f32p, u32p, C.c_size_t, C.c_uint32, f32p, C.c_float, C.c_int, u64p]
        L.orc_partial_sums.argtypes = [f32p, u32p, C.c_size_t, C.c_uint32, i64p]
        L.orc_finalize.restype = C.c_uint32
        L.orc_finalize.argtypes = [i64p, C.c_uint32, f32p, C.c_float]
        L.orc_init.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, f32p, u32p, f32p]
        L.orc_resized_dims.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, u32p, u32p]
        L.orc_resize.argtypes = [u8p, C.c_uint32, C.c_uint32, u8p, C.c_uint32, C.c_uint32]
        L.orc_seed_pixel.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.orc_kmeans.restype = C.c_uint32
        L.orc_kmeans.argtypes = [u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(OrcOpts), f32p, f32p]
        L.orc_remap_replace.argtypes = [u8p, C.c_size_t, f32p, C.c_uint32, C.c_int, u8p, u32p, f32p, f32p]
        L.orc_dither_threshold.restype = C.c_float
        L.orc_dither_threshold.argtypes = [f32p, C.c_uint32]
        L.orc_remap_dither.argtypes = [u8p, C.c_uint32, C.c_uint32, f32p, C.c_uint32, C.c_int, u8p, u32p, f32p, f32p]
        L.orc_remap_meld.argtypes = [u8p, C.c_uint32, C.c_uint32, f32p, C.c_uint32, C.c_int, u8p]
        L.orc_pal_srgb8_to_lab.argtypes = [u8p, C.c_uint32, f32p]
        L.orc_pal_lab_to_srgb8.argtypes = [f32p, C.c_uint32, u8p]
        L.orc_synth.argtypes = [u8p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return lib().orc_num_threads()


def cie94(one, second) -> float:
    a, b = _f32(one), _f32(second)
    return float(lib().orc_cie94(_p(a, C.c_float), _p(b, C.c_float)))


def pow_f32(x: float, y: float) -> float:
    return float(lib().orc_pow_f32(x, y))


def srgb_decode_table() -> np.ndarray:
    return np.array([lib().orc_srgb_decode(i) for i in range(256)], dtype=np.float32)


def convert(rgba: np.ndarray, color_space: int = LAB) -> np.ndarray:
    rgba = _u8(rgba).reshape(-1, 4)
    out = np.empty((rgba.shape[0], 4), np.float32)
    lib().orc_convert(_p(rgba, C.c_uint8), rgba.shape[0], color_space, _p(out, C.c_float))
    return out


def revert(work: np.ndarray, color_space: int = LAB) -> np.ndarray:
    work = _f32(work).reshape(-1, 4)
    out = np.empty((work.shape[0], 4), np.uint8)
    lib().orc_revert(_p(work, C.c_float), work.shape[0], color_space, _p(out, C.c_uint8))
    return out


def assign(work: np.ndarray, cent: np.ndarray, with_margin: bool = False):
    work = _f32(work).reshape(-1, 4)
    cent = _f32(cent).reshape(-1, 4)
    n = work.shape[0]
    labels = np.empty(n, np.uint32)
    if with_margin:
        best = np.empty(n, np.float32)
        second = np.empty(n, np.float32)
        lib().orc_assign(_p(work, C.c_float), n, _p(cent, C.c_float), cent.shape[0], _p(labels, C.c_uint32),
                         _p(best, C.c_float), _p(second, C.c_float))
        return labels, best, second
    lib().orc_assign(_p(work, C.c_float), n, _p(cent, C.c_float), cent.shape[0], _p(labels, C.c_uint32), None, None)
    return labels


def update(work, labels, cent, threshold: float, sum_mode: int = 1):
    work = _f32(work).reshape(-1, 4)
    labels = np.ascontiguousarray(labels, np.uint32)
    cent = _f32(cent).reshape(-1, 4).copy()
    counts = np.zeros(cent.shape[0], np.uint64)
    conv = lib().orc_update(_p(work, C.c_float), _p(labels, C.c_uint32), work.shape[0], cent.shape[0],
                            _p(cent, C.c_float), threshold, sum_mode, _p(counts, C.c_uint64))
    return cent, int(conv), counts


def partial_sums(work, labels, k: int) -> np.ndarray:
    work = _f32(work).reshape(-1, 4)
    labels = np.ascontiguousarray(labels, np.uint32)
    acc = np.zeros((k, 4), np.int64)
    lib().orc_partial_sums(_p(work, C.c_float), _p(labels, C.c_uint32), work.shape[0], k, _p(acc, C.c_int64))
    return acc


def finalize(acc, cent, threshold: float):
    acc = np.ascontiguousarray(acc, np.int64).reshape(-1, 4)
    cent = _f32(cent).reshape(-1, 4).copy()
    conv = lib().orc_finalize(_p(acc, C.c_int64), acc.shape[0], _p(cent, C.c_float), threshold)
    return cent, int(conv)


def init(work, w: int, h: int, k: int, seed_x: int, seed_y: int):
    work = _f32(work).reshape(-1, 4)
    cent = np.zeros((k, 4), np.float32)
    idx = np.zeros(k, np.uint32)
    dist = np.zeros(k, np.float32)
    lib().orc_init(_p(work, C.c_float), w, h, k, seed_x, seed_y, _p(cent, C.c_float), _p(idx, C.c_uint32),
                   _p(dist, C.c_float))
    return cent, idx, dist


def resized_dims(w: int, h: int, max_size: int = 256):
    nw, nh = C.c_uint32(), C.c_uint32()
    lib().orc_resized_dims(w, h, max_size, C.byref(nw), C.byref(nh))
    return nw.value, nh.value


def resize(rgba: np.ndarray, dw: int, dh: int) -> np.ndarray:
    rgba = _u8(rgba)
    sh, sw = rgba.shape[:2]
    out = np.empty((dh, dw, 4), np.uint8)
    lib().orc_resize(_p(rgba, C.c_uint8), sw, sh, _p(out, C.c_uint8), dw, dh)
    return out


def shrunk(rgba: np.ndarray, max_dim: int = 256) -> np.ndarray:
    h, w = rgba.shape[:2]
    if max_dim and (w > max_dim or h > max_dim):
        nw, nh = resized_dims(w, h, max_dim)
        return resize(rgba, nw, nh)
    return _u8(rgba)


def seed_pixel(w: int, h: int, fx: float = 0.5625, fy: float = 0.93359375):
    sx, sy = C.c_int32(), C.c_int32()
    lib().orc_seed_pixel(w, h, fx, fy, C.byref(sx), C.byref(sy))
    return sx.value, sy.value


def kmeans(rgba: np.ndarray, k: int, color_space: int = LAB, opts: OrcOpts | None = None, want_trace: bool = False):
    rgba = _u8(rgba)
    h, w = rgba.shape[:2]
    o = opts or default_opts()
    cent = np.zeros((k, 4), np.float32)
    trace = np.zeros((o.max_iter, k, 4), np.float32) if want_trace else None
    passes = lib().orc_kmeans(_p(rgba, C.c_uint8), w, h, k, color_space, C.byref(o), _p(cent, C.c_float),
                              _p(trace, C.c_float) if want_trace else None)
    if want_trace:
        return cent, int(passes), trace[:passes]
    return cent, int(passes)


def remap_replace(rgba: np.ndarray, cent: np.ndarray, color_space: int = LAB, details: bool = False):
    rgba = _u8(rgba)
    shape = rgba.shape
    flat = rgba.reshape(-1, 4)
    cent = _f32(cent).reshape(-1, 4)
    n = flat.shape[0]
    out = np.empty_like(flat)
    labels = np.empty(n, np.uint32)
    best = np.empty(n, np.float32)
    second = np.empty(n, np.float32)
    lib().orc_remap_replace(_p(flat, C.c_uint8), n, _p(cent, C.c_float), cent.shape[0], color_space,
                            _p(out, C.c_uint8), _p(labels, C.c_uint32), _p(best, C.c_float), _p(second, C.c_float))
    out = out.reshape(shape)
    return (out, labels, best, second) if details else out


def dither_threshold(cent: np.ndarray) -> float:
    cent = _f32(cent).reshape(-1, 4)
    return float(lib().orc_dither_threshold(_p(cent, C.c_float), cent.shape[0]))


def remap_dither(rgba: np.ndarray, cent: np.ndarray, color_space: int = LAB, details: bool = False):
    rgba = _u8(rgba)
    h, w = rgba.shape[:2]
    cent = _f32(cent).reshape(-1, 4)
    out = np.empty_like(rgba)
    n = h * w
    labels = np.empty(n, np.uint32)
    best = np.empty(n, np.float32)
    second = np.empty(n, np.float32)
    lib().orc_remap_dither(_p(rgba, C.c_uint8), w, h, _p(cent, C.c_float), cent.shape[0], color_space,
                           _p(out, C.c_uint8), _p(labels, C.c_uint32), _p(best, C.c_float), _p(second, C.c_float))
    return (out, labels, best, second) if details else out


def remap_meld(rgba: np.ndarray, cent: np.ndarray, color_space: int = LAB) -> np.ndarray:
    rgba = _u8(rgba)
    h, w = rgba.shape[:2]
    cent = _f32(cent).reshape(-1, 4)
    out = np.empty_like(rgba)
    lib().orc_remap_meld(_p(rgba, C.c_uint8), w, h, _p(cent, C.c_float), cent.shape[0], color_space, _p(out, C.c_uint8))
    return out


def pal_srgb8_to_lab(colors: np.ndarray) -> np.ndarray:
    colors = _u8(colors).reshape(-1, 4)
    out = np.empty((colors.shape[0], 4), np.float32)
    lib().orc_pal_srgb8_to_lab(_p(colors, C.c_uint8), colors.shape[0], _p(out, C.c_float))
    return out


def pal_lab_to_srgb8(lab: np.ndarray) -> np.ndarray:
    lab = _f32(lab).reshape(-1, 4)
    out = np.empty((lab.shape[0], 4), np.uint8)
    lib().orc_pal_lab_to_srgb8(_p(lab, C.c_float), lab.shape[0], _p(out, C.c_uint8))
    return out


def synth(n: int, first_pixel: int = 0, frame: int = 0, seed: int = 0, blobs: int = 0) -> np.ndarray:
    out = np.empty((n, 4), np.uint8)
    lib().orc_synth(_p(out, C.c_uint8), first_pixel, n, frame, seed, blobs)
    return out


# ---- reference-level compositions (mirror core/src/lib.rs:67-164,255-286) ------------------------

def find(rgba: np.ndarray, colors_rgba8: np.ndarray, mode: str = "replace") -> np.ndarray:
    cent = pal_srgb8_to_lab(colors_rgba8)
    return {"replace": remap_replace, "dither": remap_dither, "meld": remap_meld}[mode](rgba, cent, LAB)


def reduce(rgba: np.ndarray, k: int, mode: str = "replace", opts: OrcOpts | None = None):
    cent, passes = kmeans(rgba, k, LAB, opts)
    out = {"replace": remap_replace, "dither": remap_dither, "meld": remap_meld}[mode](rgba, cent, LAB)
    return out, cent, passes


def palette(rgba: np.ndarray, k: int, opts: OrcOpts | None = None) -> np.ndarray:
    cent, _ = kmeans(rgba, k, LAB, opts)
    cols = pal_lab_to_srgb8(cent)
    key = pal_srgb8_to_lab(cols)[:, 0]
    return cols[np.argsort(key, kind="stable")]
