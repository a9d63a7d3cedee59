"""Time k_convert (8192x8192) with CUDA events."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import kmeans_gpu_b200 as K, kmeans_gpu_b200.device as D
proc = K.ImageProcessor(0)
side = 8192
img = D.synth(proc, side * side, seed=2, blobs=16).view(side, side, 4)
for _ in range(3):
    work = D.convert(proc, img)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    work = D.convert(proc, img)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"convert 8192^2: {ms:.3f} ms, {side*side*20/ms/1e6:.0f} GB/s (20 B/px)")
