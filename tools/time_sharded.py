"""Time the two sharded jobs of bench.py alone (for A/B runs of exchange code; KMG_LIB_PATH selects the
library): the k = 8 pass on 8192 x (8192 N) pixels and config 4 (one 8192^2 image, k = 256: init + 16 passes).
usage: python -m torch.distributed.run --nproc-per-node N ... tools/time_sharded.py [reps]"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
import kmeans_gpu_b200 as K
import kmeans_gpu_b200.device as D

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
proc = K.ImageProcessor(local)
if world > 1:
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid = torch.frombuffer(bytearray(D.comm_unique_id(proc)), dtype=torch.uint8).to(dev)
    dist.broadcast(uid, 0)
    D.comm_init(proc, bytes(uid.cpu().numpy().tobytes()), world, rank)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def tmax(ms):
    if world == 1:
        return ms
    t = torch.tensor(ms, device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


W = H = 8192
opts = K.Opts(max_dim=0, max_iter=1 << 30, check_every=0)
n = W * H
img = D.synth(proc, n, first_pixel=rank * n, seed=1, blobs=8, device=dev).view(H, W, 4)
work = D.convert(proc, img)
job = D.Job(proc, work, W, H, 8, opts=opts)
if world > 1:
    job.set_shard(W, H * world, rank * H)
job.init()
job.step(5)
for rep in range(reps):
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    job.step(300)
    ev[1].record()
    barrier()
    ms = tmax([ev[0].elapsed_time(ev[1]) / 300])
    if rank == 0:
        print(f"k=8 pass, {world} GPUs x 8192^2: {ms[0] * 1e3:.2f} us", flush=True)
job.close()
del work, img
torch.cuda.empty_cache()

rows = K.row_shards(H, world)[rank]
n4 = W * (rows[1] - rows[0])
img4 = D.synth(proc, n4, first_pixel=W * rows[0], seed=1, blobs=512, device=dev)
work4 = D.convert(proc, img4)
for rep in range(reps):
    job4 = D.Job(proc, work4, W, rows[1] - rows[0], 256, opts=opts)
    if world > 1:
        job4.set_shard(W, H, rows[0])
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    job4.init()
    ev[1].record()
    job4.step(16)
    ev[2].record()
    barrier()
    ms = tmax([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])])
    if rank == 0:
        print(f"config 4 on {world} GPUs: init {ms[0]:.3f} ms, 16 passes {ms[1]:.3f} ms", flush=True)
    job4.close()
if world > 1:
    D.comm_destroy(proc)
    dist.destroy_process_group()
