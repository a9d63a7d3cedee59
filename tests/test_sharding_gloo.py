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


def _init_worker(rank, world, port, q, case):
    """Sharded farthest-point init exactly as the peer-mailbox path runs it: every rank folds the
    previous centroid into its shard's running minimum, finds its local arg-max key, 'posts'
    (key, candidate pixel, colour) to all ranks (all_gather here) and applies the merge rule."""
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as O
    import kmeans_gpu_b200 as K
    from kmeans_gpu_b200 import sharding as S

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w, h, k, img = case
        lab_all = O.convert(img)
        r0, r1 = K.row_shards(h, world)[rank]
        first, n_loc = r0 * w, (r1 - r0) * w
        lab = lab_all[first:first + n_loc]
        sx, sy = O.seed_pixel(w, h)
        cent = [lab_all[sy * w + sx].copy()]  # the seed colour is shared by one 16-byte all-reduce
        dmin = np.full(n_loc, 1000000.0, np.float32)
        picks = [sy * w + sx]
        for j in range(1, k):
            d = np.array([O.cie94(p[:3], cent[-1][:3]) for p in lab], np.float32)
            dmin = np.minimum(dmin, d)
            keys = [S.init_key(int(b), first + i) for i, b in enumerate(dmin.view(np.uint32))]
            key, pix = S.local_init_candidate(max(keys), first, n_loc)
            colour = lab_all[pix] if pix != S.NO_CANDIDATE else np.zeros(4, np.float32)
            # "mailbox": key as two int64 halves (gloo has no uint64), candidate pixel, colour bits
            post = torch.tensor([key >> 32, key & 0xFFFFFFFF, pix if pix != S.NO_CANDIDATE else -1,
                                 *[int(x) for x in colour[:3].view(np.uint32)]], dtype=torch.int64)
            box = [torch.zeros_like(post) for _ in range(world)]
            dist.all_gather(box, post)
            cands = [((int(b[0]) << 32) | int(b[1]), int(b[2]) if int(b[2]) >= 0 else S.NO_CANDIDATE) for b in box]
            kmax, owner = S.merge_init_candidates(cands)
            c = np.array([int(x) for x in box[owner][3:6]], np.uint32).view(np.float32)
            cent.append(np.array([c[0], c[1], c[2], 1.0], np.float32))
            picks.append(S.key_to_pixel(kmax))
        q.put((rank, picks, np.stack(cent)[:, :3].tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case_name", ["blobs", "three_colours"])
def test_sharded_init_mailbox_rule_matches_single_shard(oracle, case_name):
    if case_name == "blobs":
        w, h, k = 40, 26, 6
        img = oracle.synth(w * h, seed=8, blobs=12).reshape(h, w, 4)
    else:
        # only three distinct colours, k = 5: rounds 4 and 5 see an all-zero maximum, which resolves
        # to global pixel 0 (held by rank 0 alone)
        w, h, k = 24, 10, 5
        img = np.zeros((h, w, 4), np.uint8)
        img[..., 3] = 255
        img[:, :8, 0] = 200
        img[:, 8:16, 1] = 180
        img[5:, 16:, 2] = 90
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_init_worker, args=(r, world, port, q, (w, h, k, img))) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lab = oracle.convert(img)
    sx, sy = oracle.seed_pixel(w, h)
    cent, idx, _ = oracle.init(lab, w, h, k, sx, sy)
    for rank, picks, cbytes in results:
        assert picks == [int(i) for i in idx], f"rank {rank}: picks differ from the single-shard init"
        assert cbytes == np.ascontiguousarray(cent[:, :3]).tobytes()
