"""In-tree build of libkmeans_gpu.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libkmeans_gpu.so"


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    srcs = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.cpp")) + [PKG.parent / "include" / "kmeans_gpu.h"]
    return any(s.stat().st_mtime > t for s in srcs)


def build(force: bool = False) -> Path:
    if force or needs_build():
        # -B always: needs_build() has already decided that the library is stale, whatever make's
        # own dependency list thinks
        subprocess.check_call(["make", "-C", str(CSRC), "-B"])
    return LIB


if __name__ == "__main__":
    print(build())
