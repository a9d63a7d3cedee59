"""Minimal driver for ncu / timing of the farthest-point init rounds: python tools/prof_init.py [k] [side]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import kmeans_gpu_b200 as K, kmeans_gpu_b200.device as D
k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
side = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
proc = K.ImageProcessor(0)
img = D.synth(proc, side * side, seed=2, blobs=2 * k).view(side, side, 4)
work = D.convert(proc, img)
job = D.Job(proc, work, side, side, k, opts=K.Opts(max_dim=0, max_iter=1 << 30, check_every=0))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
idx, dist = job.init()
b.record()
torch.cuda.synchronize()
print("k", k, "side", side, "init ms", round(a.elapsed_time(b), 3), "per round", round(a.elapsed_time(b) / (k - 1), 4),
      "picks", idx[:4], dist[:4])
