"""End-to-end time of reduce() on a large pinned host image, pipelined against serial
(KMG_NO_REDUCE_PIPELINE=1): python tools/time_reduce_large.py [side]"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import kmeans_gpu_b200 as K, oracle_lib as O
side = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
proc = K.ImageProcessor(0)
img = K.pinned_empty((side, side, 4)); img[...] = O.synth(side * side, seed=4, blobs=16).reshape(side, side, 4)
out = K.pinned_empty((side, side, 4))
for _ in range(2):
    proc.reduce(8, img, reduce_mode=K.ReduceMode.Dither, out=out)
t0 = time.perf_counter()
for _ in range(5):
    proc.reduce(8, img, reduce_mode=K.ReduceMode.Dither, out=out)
dt = (time.perf_counter() - t0) / 5
print("pipeline", "off" if os.environ.get("KMG_NO_REDUCE_PIPELINE") else "on", f"reduce k=8 dither {side}x{side}: {dt*1e3:.2f} ms  {1/dt:.1f} images/s")
