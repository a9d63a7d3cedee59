"""Time the farthest-point initialisation of BASELINE config 4 (8192 x 8192, k = 256) and print the
work the lazy rounds did.  usage: time_init.py [side] [k] [blobs]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import kmeans_gpu_b200 as K, kmeans_gpu_b200.device as D
side = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else 256
blobs = int(sys.argv[3]) if len(sys.argv) > 3 else 2 * k
proc = K.ImageProcessor(0)
img = D.synth(proc, side * side, seed=2, blobs=blobs).view(side, side, 4)
work = D.convert(proc, img)
job = D.Job(proc, work, side, side, k, opts=K.Opts(max_dim=0, max_iter=1 << 30, check_every=0))
for rep in range(3):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); idx, dist = job.init(); b.record(); torch.cuda.synchronize()
    print(f"init {side}x{side} k={k} blobs={blobs}: {a.elapsed_time(b):.3f} ms", job.init_stats(), "last picks", idx[-2:].tolist(), dist[-2:].tolist())
