"""Phase timeline of k_kmeans_small (needs `make -C kmeans-gpu_b200/csrc TRACE=1` and KMG_LIB_PATH pointing at it)."""
import ctypes as C, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import kmeans_gpu_b200 as K
from kmeans_gpu_b200 import _native
from PIL import Image as PILImage
proc = K.ImageProcessor(0)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tokyo = np.array(PILImage.open(ROOT / "tests" / "golden" / "tokyo.png").convert("RGBA"))
for _ in range(3):
    cent, passes = proc.kmeans_centroids(k, tokyo)
buf = np.zeros((16, 512), np.uint64)
lib = _native.load()
lib.kmg_debug_small_trace.argtypes = [C.c_void_p]
assert lib.kmg_debug_small_trace(buf.ctypes.data_as(C.c_void_p)) == 0
t = buf.astype(np.int64)
n_init = 4 * (k - 1)
names = ["start", "converted", "csync"] + [f"init{j}.{p}" for j in range(1, k) for p in "abcd"]
per_pass = ["table", "assigned", "folded", "csync", "final"]
names += [f"p{i}.{p}" for i in range(passes) for p in per_pass] + ["table_last", "end"]
print("passes", passes, "marks", len(names))
for r in (0, 7, 15):
    d = np.diff(t[r, :len(names)])
    print(f"rank {r}: total {t[r, len(names)-1]-t[r,0]} cycles")
    print("  head:", {names[i + 1]: int(d[i]) for i in range(2)})
    ini = d[2:2 + n_init].reshape(k - 1, 4) if k > 1 else np.zeros((0, 4))
    print("  init mean per round [scan, bsync+send, csync, fetch]:", ini.mean(axis=0).round(0) if k > 1 else None)
    ps = d[2 + n_init:2 + n_init + 5 * passes].reshape(passes, 5)
    print("  pass mean [table, assign, bsync+fold+send, csync, finalise]:", ps.mean(axis=0).round(0), "sum", ps.mean(axis=0).sum().round(0))
    print("  first pass:", ps[0], " last:", ps[-1])
    print("  tail:", d[2 + n_init + 5 * passes:])
