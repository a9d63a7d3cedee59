"""Time the fused remap kernel: python tools/prof_remap.py <k> <mode 0|1|2> <w> <h> [blobs]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import kmeans_gpu_b200 as K, kmeans_gpu_b200.device as D
k = int(sys.argv[1]); mode = int(sys.argv[2]); w = int(sys.argv[3]); h = int(sys.argv[4])
blobs = int(sys.argv[5]) if len(sys.argv) > 5 else 0
proc = K.ImageProcessor(0)
img = D.synth(proc, w * h, seed=1, blobs=blobs).view(h, w, 4)
rng = np.random.default_rng(k)
if k == 64:
    cols = K.parse_palette(ROOT / "tests" / "golden" / "resurrect_64.png")
else:
    cols = rng.integers(0, 256, (k, 4), dtype=np.uint8); cols[:, 3] = 255
cent = K.fixed_centroids(cols)
work = torch.empty((8, 4), dtype=torch.float32, device="cuda")
job = D.Job(proc, work, 8, 1, k)
job.set_centroids(cent)
out = torch.empty_like(img)
for _ in range(3):
    job.remap(img, K.ReduceMode(mode), out=out)
torch.cuda.synchronize()
reps = 10
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    job.remap(img, K.ReduceMode(mode), out=out)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)])
st = job.stats()
print(f"remap k={k} mode={mode} {w}x{h}: {ms:.4f} ms  {w*h/ms/1e3:.1f} Mpix/s  {8*w*h/ms/1e6:.1f} GB/s  slow_px/pass={st['slow_pixels']/(reps+3):.0f} (incl. prepare launch)")
