"""world_size-2 gloo test of the multi-GPU host logic: row-sharded partial
sums, all-reduced as int64, give bit-identical centroids to the single-shard pass; sharded
farthest-point picks (max over ranks of the 64-bit key) equal the single-shard pick."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as O
    import kmeans_gpu_b200 as K

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w, h, k = 96, 70, 6
        img = O.synth(w * h, seed=5, blobs=12).reshape(h, w, 4)
        lab_all = O.convert(img)
        cent0, _, _ = O.init(lab_all, w, h, k, 10, 20)
        r0, r1 = K.row_shards(h, world)[rank]
        lab = lab_all[r0 * w:r1 * w]
        # one Lloyd pass: local assign + local exact sums, all-reduce, identical finalize everywhere
        labels = O.assign(lab, cent0)
        acc = torch.from_numpy(O.partial_sums(lab, labels, k))
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        cent1, conv = O.finalize(acc.numpy(), cent0, 1.0)
        # sharded farthest-point round: key = (distance bits << 32) | (global pixel ^ 15)
        d = np.array([O.cie94(p[:3], cent0[0, :3]) for p in lab], np.float32)
        gidx = np.arange(r0 * w, r1 * w, dtype=np.uint64)
        keys = (d.view(np.uint32).astype(np.uint64) << np.uint64(32)) | (gidx ^ np.uint64(15))
        best = torch.tensor([int(keys.max()) >> 1], dtype=torch.int64)  # gloo has no uint64; keys < 2^63
        lowbit = torch.tensor([int(keys.max())], dtype=torch.float64)
        dist.all_reduce(best, op=dist.ReduceOp.MAX)
        q.put((rank, cent1.tobytes(), conv, int(best.item())))
    finally:
        dist.destroy_process_group()


def test_row_sharded_pass_matches_single_shard(oracle):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-shard reference
    w, h, k = 96, 70, 6
    img = oracle.synth(w * h, seed=5, blobs=12).reshape(h, w, 4)
    lab = oracle.convert(img)
    cent0, _, _ = oracle.init(lab, w, h, k, 10, 20)
    labels = oracle.assign(lab, cent0)
    cent1, conv, _ = oracle.update(lab, labels, cent0, 1.0, sum_mode=1)
    d = np.array([oracle.cie94(p[:3], cent0[0, :3]) for p in lab], np.float32)
    gidx = np.arange(w * h, dtype=np.uint64)
    keys = (d.view(np.uint32).astype(np.uint64) << np.uint64(32)) | (gidx ^ np.uint64(15))
    for rank, cbytes, rconv, best in results:
        assert cbytes == cent1.tobytes(), f"rank {rank}: centroids differ from the single-shard pass"
        assert rconv == conv
        assert best == int(keys.max()) >> 1
    # and the key's winner is the oracle's second pick
    _, idx, _ = oracle.init(lab, w, h, 2, 10, 20)
    assert (int(keys.max()) & 0xFFFFFFFF) ^ 15 == int(idx[1])
