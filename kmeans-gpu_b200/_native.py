"""ctypes binding of libkmeans_gpu.so (include/kmeans_gpu.h).

There is no fallback: if the shared library is missing the import of any compute entry point
raises, and every compute call fails with KmgError when no B200 is present.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
import os

# KMG_LIB_PATH: development override (e.g. the `make TRACE=1` build); the product is lib/libkmeans_gpu.so
LIB_PATH = Path(os.environ["KMG_LIB_PATH"]) if os.environ.get("KMG_LIB_PATH") else PKG / "lib" / "libkmeans_gpu.so"

KMG_OK = 0
STATUS_NAMES = {1: "BAD_ARG", 2: "CUDA", 3: "OOM", 4: "NCCL", 5: "UNSUPPORTED"}


class KmgError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"kmeans_gpu error {STATUS_NAMES.get(code, code)}: {message}")
        self.code = code
        self.message = message


class KmgOpts(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("max_dim", C.c_uint32),
        ("max_iter", C.c_uint32),
        ("check_every", C.c_uint32),
        ("convergence", C.c_float),
        ("seed_x_frac", C.c_float),
        ("seed_y_frac", C.c_float),
        ("seed_x", C.c_int32),
        ("seed_y", C.c_int32),
        ("flags", C.c_uint32),
    ]


KMG_OPT_NO_FUSED_KMEANS = 1


_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/kmeans_gpu.h one to one
SIGNATURES = {
    "kmg_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "kmg_destroy": (None, [_vp]),
    "kmg_last_error": (C.c_char_p, []),
    "kmg_abi_version": (C.c_int, []),
    "kmg_default_opts": (None, [C.POINTER(KmgOpts)]),
    "kmg_kmeans_palette": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(KmgOpts), _f32p, _u32p]),
    "kmg_remap": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _f32p, C.c_uint32, C.c_int, C.c_int, _vp]),
    "kmg_reduce": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(KmgOpts), _vp, _f32p, _u32p]),
    "kmg_resized_dims": (None, [C.c_uint32, C.c_uint32, C.c_uint32, _u32p, _u32p]),
    "kmg_resize": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    "kmg_reduce_batch": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(KmgOpts), _vp, _f32p, _u32p]),
    "kmg_fixed_centroids": (None, [_u8p, C.c_uint32, C.c_int, _f32p]),
    "kmg_centroids_to_rgba8": (None, [_f32p, C.c_uint32, C.c_int, _u8p]),
    "kmg_sort_palette_by_lightness": (None, [_u8p, C.c_uint32]),
    "kmg_octree_palette": (C.c_int, [_u8p, C.c_uint64, C.c_uint32, _u8p, _u32p]),
    "kmg_dev_convert": (C.c_int, [_vp, _vp, C.c_uint64, C.c_int, _vp, _vp]),
    "kmg_dev_resize": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _vp, C.c_uint32, C.c_uint32, _vp]),
    "kmg_dev_assign": (C.c_int, [_vp, _vp, C.c_uint64, _f32p, C.c_uint32, _vp, _vp]),
    "kmg_job_create": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(KmgOpts), C.POINTER(_vp)]),
    "kmg_job_destroy": (None, [_vp]),
    "kmg_job_init": (C.c_int, [_vp, _u32p, _f32p, _vp]),
    "kmg_job_set_centroids": (C.c_int, [_vp, _f32p, _vp]),
    "kmg_job_get_centroids": (C.c_int, [_vp, _f32p, _vp]),
    "kmg_job_step": (C.c_int, [_vp, C.c_uint32, _vp]),
    "kmg_job_run": (C.c_int, [_vp, _u32p, _vp]),
    "kmg_job_stats": (C.c_int, [_vp, _u32p, _u32p, _u64p, _vp]),
    "kmg_job_init_stats": (C.c_int, [_vp, _u32p, _u64p, _u64p, _u64p, _vp]),
    "kmg_job_get_sums": (C.c_int, [_vp, C.POINTER(C.c_int64), _vp]),
    "kmg_comm_unique_id": (C.c_int, [_vp, _u8p]),
    "kmg_comm_init": (C.c_int, [_vp, _u8p, C.c_int, C.c_int]),
    "kmg_comm_destroy": (C.c_int, [_vp]),
    "kmg_comm_mode": (C.c_int, [_vp]),
    "kmg_job_set_shard": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32]),
    "kmg_dev_remap": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _f32p, C.c_uint32, C.c_int, C.c_int, _vp, _vp]),
    "kmg_dev_remap_job": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _vp, C.c_int, _vp, _vp]),
    "kmg_dev_reduce_batch": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(KmgOpts), _vp, _f32p, _u32p, _vp]),
    "kmg_dev_synth": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    "kmg_dev_srgb_table": (C.c_int, [_vp, _f32p]),
    "kmg_dev_fast_lab_error": (C.c_int, [_vp, _f32p]),
    "kmg_dev_fp32_peak": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "kmg_dev_audit": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_uint32, _f32p, C.c_uint32, C.c_int, C.c_int, C.c_int, _u64p, _u64p, _vp]),
    "kmg_launch_count": (C.c_uint64, [_vp]),
    "kmg_alloc_pinned": (C.c_void_p, [C.c_size_t]),
    "kmg_free_pinned": (None, [_vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). kmeans_gpu_b200 has no CPU or PyTorch fallback."
            )
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(code: int) -> None:
    if code != KMG_OK:
        raise KmgError(code, load().kmg_last_error().decode("utf-8", "replace"))
