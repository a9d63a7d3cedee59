"""Minimal driver for ncu: a few Lloyd passes on the bench workload (8192x8192, k from argv)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import kmeans_gpu_b200 as K, kmeans_gpu_b200.device as D
k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
side = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
proc = K.ImageProcessor(0)
img = D.synth(proc, side * side, seed=2, blobs=2 * k).view(side, side, 4)
work = D.convert(proc, img)
job = D.Job(proc, work, side, side, k, opts=K.Opts(max_dim=0, max_iter=1 << 30, check_every=0))
job.init()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
ev[0].record()
for i in range(steps):
    job.step(1)
    ev[i + 1].record()
torch.cuda.synchronize()
print("k", k, "side", side, "ms per pass", [round(ev[i].elapsed_time(ev[i + 1]), 4) for i in range(steps)], job.stats())
