"""Minimal driver for ncu: the fused small-image k-means (tokyo k=8, and a 32-frame 1080p k=16 batch)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import kmeans_gpu_b200 as K, kmeans_gpu_b200.device as D
from PIL import Image as PILImage
proc = K.ImageProcessor(0)
what = sys.argv[1] if len(sys.argv) > 1 else "tokyo"
if what == "tokyo":
    tokyo = np.array(PILImage.open(ROOT / "tests" / "golden" / "tokyo.png").convert("RGBA"))
    for _ in range(4):
        out, cent, passes = proc.reduce(8, tokyo, reduce_mode=K.ReduceMode.Dither, return_details=True)
    print("passes", passes)
else:
    nf = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    frames = torch.empty((nf, 1080, 1920, 4), dtype=torch.uint8, device="cuda")
    for f in range(nf):
        D.synth(proc, 1920 * 1080, frame=f, seed=3, blobs=32, out=frames[f].view(-1, 4))
    out = torch.empty_like(frames)
    for _ in range(3):
        _, cent, passes = D.reduce_batch(proc, frames, 16, K.ReduceMode.Dither, out=out)
    print("passes", passes.mean())
