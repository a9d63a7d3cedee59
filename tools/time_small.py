"""Timing of the small-image paths (tokyo reduce, 1080p frame batches) — development helper."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import kmeans_gpu_b200 as K
import kmeans_gpu_b200.device as D
from PIL import Image as PILImage

proc = K.ImageProcessor(0)
tokyo = np.array(PILImage.open(ROOT / "tests" / "golden" / "tokyo.png").convert("RGBA"))
tk = torch.from_numpy(tokyo).pin_memory().numpy()
for fused in (True, False):
    o = K.Opts(fused_kmeans=fused)
    for _ in range(3):
        proc.reduce(8, tk, reduce_mode=K.ReduceMode.Dither, opts=o)
    l0 = proc.launch_count()
    t0 = time.perf_counter()
    reps = 50
    for _ in range(reps):
        proc.reduce(8, tk, reduce_mode=K.ReduceMode.Dither, opts=o)
    dt = (time.perf_counter() - t0) / reps
    print(f"tokyo reduce k=8 dither fused={fused}: {dt*1e6:.1f} us/call, launches/call={(proc.launch_count()-l0)/reps}")
    t0 = time.perf_counter()
    for _ in range(reps):
        proc.palette(8, tk, opts=o)
    dt = (time.perf_counter() - t0) / reps
    print(f"tokyo palette k=8 fused={fused}: {dt*1e6:.1f} us/call")

dev = torch.device("cuda", 0)
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 256
frames = torch.empty((nf, 1080, 1920, 4), dtype=torch.uint8, device=dev)
for f in range(nf):
    D.synth(proc, 1920 * 1080, frame=f, seed=3, blobs=32, out=frames[f].view(-1, 4))
outb = torch.empty_like(frames)
torch.cuda.synchronize()
for fused in (True, False):
    o = K.Opts(fused_kmeans=fused)
    n = nf if fused else min(nf, 32)
    D.reduce_batch(proc, frames[:4], 16, K.ReduceMode.Dither, out=outb[:4], opts=o)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, cent, passes = D.reduce_batch(proc, frames[:n], 16, K.ReduceMode.Dither, out=outb[:n], opts=o)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"resident 1080p k=16 reduce+dither fused={fused}: {n/dt:.0f} frames/s ({n} frames, mean passes {passes.mean():.1f})")
# stage timing with events
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
import ctypes as C
# host pipeline
hn = min(nf, 128)
host = torch.empty((hn, 1080, 1920, 4), dtype=torch.uint8).pin_memory()
host.copy_(frames[:hn])
hout = torch.empty_like(host).pin_memory()
hnp = host.numpy()
proc.reduce_batch(16, hnp[:16], K.ReduceMode.Dither)
t0 = time.perf_counter()
out, cent, passes = proc.reduce_batch(16, hnp, K.ReduceMode.Dither)
dt = time.perf_counter() - t0
print(f"host-pipelined 1080p k=16 reduce+dither: {hn/dt:.0f} frames/s (output into a pageable numpy array)")
