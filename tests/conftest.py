import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def load_rgba(name: str) -> np.ndarray:
    from PIL import Image

    return np.array(Image.open(GOLDEN / name).convert("RGBA"))


@pytest.fixture(scope="session")
def tokyo() -> np.ndarray:
    return load_rgba("tokyo.png")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def native_lib():
    """The product library, built in-tree (nvcc cross-compiles without a GPU)."""
    import kmeans_gpu_b200  # noqa: F401
    from importlib import import_module

    build = import_module("kmeans_gpu_b200.build")
    build.build()
    native = import_module("kmeans_gpu_b200._native")
    return native.load()


@pytest.fixture(scope="session")
def proc(native_lib):
    import kmeans_gpu_b200 as K

    p = K.ImageProcessor(0)
    yield p
    p.close()
