"""Multi-GPU parity check (run under torchrun, one rank per GPU): a row-sharded k-means job must
produce, bit for bit, the centroids / integer sums / pass count of the same job on one GPU.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi_gpu.py
  KMG_NO_P2P=1 ...   forces the NCCL all-reduce path instead of the in-kernel peer exchange
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch
import torch.distributed as dist

import kmeans_gpu_b200 as K
import kmeans_gpu_b200.device as D

world = int(os.environ["WORLD_SIZE"])
rank = int(os.environ["RANK"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
proc = K.ImageProcessor(local)
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    uid = torch.frombuffer(bytearray(D.comm_unique_id(proc)), dtype=torch.uint8).to(dev)
dist.broadcast(uid, 0)
D.comm_init(proc, bytes(uid.cpu().numpy().tobytes()), world, rank)
mode = D.comm_mode(proc)
want_mode = 1 if os.environ.get("KMG_NO_P2P") else 2
assert mode == want_mode, f"rank {rank}: comm mode {mode}, expected {want_mode}"

failures = 0
for (W, H, k, passes) in ((1024, 96 * world, 8, 12), (640, 50 * world, 16, 9), (333, 31 * world, 40, 5), (512, 64 * world, 256, 3)):
    rows = K.row_shards(H, world)[rank]
    n_loc = W * (rows[1] - rows[0])
    img = D.synth(proc, n_loc, first_pixel=W * rows[0], seed=5, blobs=2 * k, device=dev)
    work = D.convert(proc, img)
    opts = K.Opts(max_dim=0, max_iter=passes, check_every=4)
    job = D.Job(proc, work, W, rows[1] - rows[0], k, opts=opts)
    job.set_shard(W, H, rows[0])
    job.init()
    done = job.run()
    cent = job.centroids()
    sums = job.sums()
    job.close()
    if rank == 0:
        solo = K.ImageProcessor(local)
        img1 = D.synth(solo, W * H, seed=5, blobs=2 * k, device=dev)
        work1 = D.convert(solo, img1)
        job1 = D.Job(solo, work1, W, H, k, opts=opts)
        job1.init()
        done1 = job1.run()
        cent1 = job1.centroids()
        sums1 = job1.sums()
        job1.close()
        solo.close()
        ok = done == done1 and np.array_equal(cent.view(np.uint32), cent1.view(np.uint32)) and np.array_equal(sums, sums1)
        print(f"{W}x{H} k={k}: sharded over {world} GPUs (mode {mode}) passes {done} vs single {done1}: {'OK' if ok else 'MISMATCH'}")
        failures += 0 if ok else 1
    # every rank must hold the same centroids
    t = torch.from_numpy(cent.view(np.int32).copy()).to(dev)
    ref = t.clone()
    dist.broadcast(ref, 0)
    if not torch.equal(t, ref):
        print(f"rank {rank}: centroids differ from rank 0")
        failures += 1
# fewer distinct colours than clusters: the late init rounds see an all-zero maximum, which resolves
# to global pixel 0 — held by rank 0 alone, every other rank posts "no candidate"
W, H, k = 64, 16 * world, 6
host = np.zeros((H, W, 4), np.uint8)
host[..., 3] = 255
host[:, :20, 0] = 200
host[:, 20:40, 1] = 180
host[H // 2:, 40:, 2] = 90
rows = K.row_shards(H, world)[rank]
opts = K.Opts(max_dim=0, max_iter=3, check_every=0)
work = D.convert(proc, torch.from_numpy(host[rows[0]:rows[1]].copy()).to(dev))
job = D.Job(proc, work, W, rows[1] - rows[0], k, opts=opts)
job.set_shard(W, H, rows[0])
idx, dst = job.init()
job.run()
cent = job.centroids()
job.close()
if rank == 0:
    solo = K.ImageProcessor(local)
    work1 = D.convert(solo, torch.from_numpy(host).to(dev))
    job1 = D.Job(solo, work1, W, H, k, opts=opts)
    idx1, dst1 = job1.init()
    job1.run()
    cent1 = job1.centroids()
    job1.close()
    solo.close()
    ok = idx.tolist() == idx1.tolist() and np.array_equal(dst.view(np.uint32), dst1.view(np.uint32)) and \
        np.array_equal(cent.view(np.uint32), cent1.view(np.uint32))
    print(f"{W}x{H} four colours k={k}: picks {idx.tolist()} vs single {idx1.tolist()}: {'OK' if ok else 'MISMATCH'}")
    failures += 0 if ok else 1
flag = torch.tensor([failures], device=dev)
dist.all_reduce(flag)
D.comm_destroy(proc)
proc.close()
dist.destroy_process_group()
sys.exit(1 if int(flag[0]) else 0)
