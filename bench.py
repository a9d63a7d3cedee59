#!/usr/bin/env python
"""bench.py — headline benchmark of the image hot path (contract: see DESIGN.md "Measurement").

Workload (BASELINE.json target config): one k-means iteration = one fused assign+update pass over
an 8192x8192 image at k=8, Lab, cached float4 work plane (16 B/px), synthetic blobs(16) image made
by the device generator.  With N GPUs every rank owns an 8192x8192 row block of one 8192x(8192*N)
image (weak scaling) and the k x 4 integer sums are all-reduced over NCCL every pass.

  value  : Mpix/s per iteration, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e    : the same metric through the host C ABI (kmg_kmeans_palette on a pinned host image:
           H2D copy + convert + farthest-point init + `E2E_PASSES` passes + centroid read-back)
  extras : end-to-end images/s for the tokyo-sized reduce (configs 1/2) and 1080p frames (config 5)

`--impl reference` times the CPU oracle (restated reference; the Rust/wgpu reference cannot be
built in this image) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

W = H = 8192
K_CLUSTERS = 8
BLOBS = 16
SEED = 2
E2E_PASSES = 16
BYTES_PER_PX = 16  # one read of the cached float4 work plane
METRIC = "Mpix/s per k-means iteration (assign+update), 8192x8192 k=8"


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons under load.  nvidia-smi needs a few hundred ms to start
    and the timed region may be shorter than that, so the sampler is started before the warm-up and
    every line is stamped on arrival: stop() reports the samples that fall inside the timed window
    and, when there are fewer than two, all samples taken under the (identical) warm-up load too."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []  # (arrival time, text)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def count(self) -> int:
        return len(self.lines)

    def stop(self, t_begin: float, t_end: float, t_load: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        inside = [ln for (t, ln) in self.lines if t_begin <= t <= t_end + 0.02]
        window = "timed region"
        if len(inside) < 2:
            inside = [ln for (t, ln) in self.lines if t_load <= t <= t_end + 0.02]
            window = "warm-up + timed region (same load; the timed region is shorter than two sampling periods)"
        sm, mx, reasons = [], [], set()
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_iteration_sample(side: int, steps: int):
    """The oracle's assign + update on a side x side crop of the same synthetic image."""
    import oracle_lib as O

    img = O.synth(side * side, seed=SEED, blobs=BLOBS)
    lab = O.convert(img)
    cent, _, _ = O.init(lab, side, side, K_CLUSTERS, int(side * 0.5625), int(side * 0.93359375))
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        labels = O.assign(lab, cent)
        cent, _, _ = O.update(lab, labels, cent, 1.0, sum_mode=1)
        times.append(time.perf_counter() - t0)
    return side * side / 1e6 / float(np.mean(times)), O.num_threads(), float(np.mean(times))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    side = 2048
    cpu_iteration_sample(512, min(max(args.warmup, 1), 3))  # warm-up (page in, OpenMP pool)
    steps = min(args.steps, 40)  # bounded sample: ~0.05 s per 2048^2 iteration on 16 cores
    mpix, threads, sec = cpu_iteration_sample(side, steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{side}x{side} crop (bounded CPU sample) of the {W}x{H} blobs({BLOBS}) k={K_CLUSTERS} Lab image: "
                               "assign+update iteration, oracle port on the host cores",
                   "k": K_CLUSTERS, "bytes_per_px": BYTES_PER_PX, "same_config": False,
                   "note": "CPU restatement of the reference (oracle/oracle.cpp), not the reference's wgpu GPU path; "
                           "Mpix/s per iteration does not depend on the crop size"},
        "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": threads, "kind": "port",
                         "sample": f"{side}x{side} crop of the same synthetic image, {steps} iterations, "
                                   "oracle/oracle.cpp (restated reference; Rust+wgpu reference not buildable here)"},
        "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def parity_job(K, D, torch, proc, dev, w, h_total, k, blobs, world, rank, passes):
    """One row-sharded k-means job (farthest-point init + `passes` passes) on all ranks; returns this
    rank's (centroids, sums).  world == 1 runs the same job unsharded."""
    rows = K.row_shards(h_total, world)[rank]
    n_loc = w * (rows[1] - rows[0])
    img = D.synth(proc, n_loc, first_pixel=w * rows[0], seed=SEED, blobs=blobs, device=dev)
    work = D.convert(proc, img)
    del img
    job = D.Job(proc, work, w, rows[1] - rows[0], k, opts=K.Opts(max_dim=0, max_iter=passes, check_every=0))
    if world > 1:
        job.set_shard(w, h_total, rows[0])
    job.init()
    job.step(passes)
    cent, sums = job.centroids(), job.sums()
    job.close()
    del work
    torch.cuda.empty_cache()
    return cent, sums


def parity_check(K, D, torch, dist, proc, dev, world, rank):
    """identical results at 1/2/4/8 GPUs.  Two sharded jobs — the headline shape
    (8192 x 8192 per GPU, k = 8) and BASELINE config 4 (one 8192 x 8192 image, k = 256) — each with
    its init and 16 passes: the k x 4 centroids and int64 sums must be bit-identical on every rank,
    and equal to what rank 0 gets when it runs the same image alone, unsharded."""
    out = {"ranks_identical": True, "equals_single_gpu": True, "jobs": []}
    solo = None
    for name, w, h_total, k, blobs in (("headline 8192 x (8192 N) k=8", W, H * world, K_CLUSTERS, BLOBS),
                                       ("config 4: 8192 x 8192 k=256", W, H, 256, 512)):
        cent, sums = parity_job(K, D, torch, proc, dev, w, h_total, k, blobs, world, rank, 16)
        mine = torch.from_numpy(np.concatenate([cent.view(np.int32).astype(np.int64).ravel(), sums.ravel()])).to(dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        same = all(bool(torch.equal(allr[0], a)) for a in allr)
        equal = None
        if rank == 0:
            # a second context without a communicator: the whole image on this GPU alone
            solo = solo or K.ImageProcessor(dev.index)
            c1, s1 = parity_job(K, D, torch, solo, dev, w, h_total, k, blobs, 1, 0, 16)
            equal = bool(np.array_equal(c1.view(np.uint32), cent.view(np.uint32)) and np.array_equal(s1, sums))
        flag = torch.tensor([1 if (equal is None or equal) else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        equal = bool(int(flag[0]))
        out["ranks_identical"] &= same
        out["equals_single_gpu"] &= equal
        out["jobs"].append({"job": name, "ranks_identical": same, "equals_single_gpu": equal, "passes": 16,
                            "compared": "k x 4 centroids (bit patterns) and k x 4 int64 sums, all-gathered"})
    if solo is not None:
        solo.close()
    return out


def bind_to_gpu_numa_node(local: int, world: int) -> dict:
    """Host side of the end-to-end numbers: page-locked staging buffers are placed by first touch, so a
    rank whose threads run on another socket than its GPU moves every frame across the inter-socket
    link twice, and eight ranks that all start on node 0 share one memory controller.  Bind this rank
    (and the OpenMP / copy threads it spawns) to the cores of the NUMA node its GPU hangs off; when the
    topology is not visible (containers often hide it) fall back to an even split of the visible cores."""
    info = {"policy": "none"}
    try:
        cores = sorted(os.sched_getaffinity(0))
        node = -1
        try:
            q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                               capture_output=True, text=True, timeout=20).stdout.strip().lower()
            bdf = q[-12:] if len(q) >= 12 else q  # 00000000:1B:00.0 -> 0000:1b:00.0
            node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text().strip())
        except Exception:
            node = -1
        chosen = None
        if node >= 0:
            try:
                txt = Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip()
                ids = set()
                for part in txt.split(","):
                    a, _, b = part.partition("-")
                    ids.update(range(int(a), int(b or a) + 1))
                chosen = [c for c in cores if c in ids]
                info = {"policy": "gpu numa node", "node": node}
            except Exception:
                chosen = None
        if not chosen and world > 1:
            per = max(1, len(cores) // world)
            chosen = cores[local * per:(local + 1) * per] or cores
            info = {"policy": "even split of the visible cores (GPU NUMA node not visible)"}
        if chosen:
            os.sched_setaffinity(0, chosen)
            info["cores"] = len(chosen)
    except Exception as e:  # never let topology probing break the bench
        info = {"policy": "none", "error": str(e)[:80]}
    return info


def run_ours(args):
    import torch
    import torch.distributed as dist

    import kmeans_gpu_b200 as K
    import kmeans_gpu_b200.device as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: kmeans_gpu_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # (at N = 1 the process keeps every core: the CPU-baseline leg of the same run uses them all)
    numa = (bind_to_gpu_numa_node(local, world) if world > 1 and not os.environ.get("KMG_BENCH_NO_BIND")
            else {"policy": "none (single GPU)" if world == 1 else "none (KMG_BENCH_NO_BIND)"})
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    proc = K.ImageProcessor(local)
    comm_mode = 0

    n = W * H
    # rank r owns rows [r*H, (r+1)*H) of the 8192 x (8192*world) image
    img = D.synth(proc, n, first_pixel=rank * n, seed=SEED, blobs=BLOBS, device=dev).view(H, W, 4)
    work = D.convert(proc, img)
    opts = K.Opts(max_dim=0, max_iter=1 << 30, check_every=0)
    job = D.Job(proc, work, W, H, K_CLUSTERS, opts=opts)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.frombuffer(bytearray(D.comm_unique_id(proc)), dtype=torch.uint8).to(dev)
        dist.broadcast(uid, 0)
        D.comm_init(proc, bytes(uid.cpu().numpy().tobytes()), world, rank)
        comm_mode = D.comm_mode(proc)
        job.set_shard(W, H * world, rank * H)
    job.init()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        # nvidia-smi needs a few hundred ms to deliver its first line: wait for it with the GPU idle (the
        # previous version kept the pass running meanwhile, i.e. 0.3 .. 1.5 s of extra, untimed load that
        # pushed the part into its 1 kW power cap before the first timed step)
        t_wait = time.perf_counter()
        while sampler.proc is not None and sampler.count() < 1 and time.perf_counter() - t_wait < 3.0:
            time.sleep(0.02)
    barrier()
    t_load = time.perf_counter()
    for _ in range(max(args.warmup, 3)):
        job.step(1)
    barrier()
    launches0 = proc.launch_count()
    # one event between consecutive steps (K + 1 in all): the end of step i is the start of step i + 1
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t_wall0 = time.perf_counter()
    evs[0].record()
    for i in range(args.steps):
        job.step(1)
        evs[i + 1].record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = proc.launch_count() - launches0
    total_ms = evs[0].elapsed_time(evs[-1])
    per_step = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)])
    step_ms = float(np.mean(per_step))  # average launch duration of the pass kernel (event to event, one launch in between)
    clocks = sampler.stop(t_wall0, t_wall1, t_load) if rank == 0 else None
    if world > 1:
        t = torch.tensor([total_ms, step_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, step_ms = float(t[0]), float(t[1])

    # ---- the same pass, sustained: 3000 more steps back to back (the part reaches its power cap) ----
    sus_sampler = ClockSampler(local)
    if rank == 0:
        sus_sampler.start()
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_s0 = time.perf_counter()
    s0.record()
    job.step(3000)
    s1.record()
    barrier()
    t_s1 = time.perf_counter()
    sus_ms = s0.elapsed_time(s1) / 3000
    sus_clocks = sus_sampler.stop(t_s0 + 0.5 * (t_s1 - t_s0), t_s1, t_s0) if rank == 0 else None
    if world > 1:
        t = torch.tensor([sus_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sus_ms = float(t[0])

    # ---- end to end through the host C ABI (pinned host image) --------------------------------
    host = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
    host.copy_(img)
    torch.cuda.synchronize()
    stats = job.stats()
    e2e_opts = K.Opts(max_dim=0, max_iter=E2E_PASSES, check_every=0)
    host_np = host.numpy()
    e2e_steps = max(1, min(args.steps, 3))
    proc.kmeans_centroids(K_CLUSTERS, host_np, opts=e2e_opts)  # warm-up (workspace allocation)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        cent, passes = proc.kmeans_centroids(K_CLUSTERS, host_np, opts=e2e_opts)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_mpix = world * n * E2E_PASSES * e2e_steps / e2e_s / 1e6

    # config 4 on this GPU's share: k = 256 iteration over the same plane
    job.close()
    del work, host, img
    img = D.synth(proc, n, first_pixel=rank * n, seed=SEED, blobs=512, device=dev).view(H, W, 4)
    work = D.convert(proc, img)
    job256 = D.Job(proc, work, W, H, 256, opts=opts)
    job256.init()
    for _ in range(2):
        job256.step(1)
    torch.cuda.synchronize()
    a256, b256 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a256.record()
    job256.step(3)
    b256.record()
    torch.cuda.synchronize()
    ms256 = a256.elapsed_time(b256) / 3
    st256 = job256.stats()
    job256.close()
    del work, img
    torch.cuda.empty_cache()

    # BASELINE config 4 as stated: ONE 8192x8192 image, k = 256, rows sharded over the N GPUs
    # (strong scaling): farthest-point init (255 rounds) + 16 passes, every exchange inside the kernels
    rows4 = K.row_shards(H, world)[rank]
    n4 = W * (rows4[1] - rows4[0])
    img4 = D.synth(proc, n4, first_pixel=W * rows4[0], seed=SEED, blobs=512, device=dev)
    work4 = D.convert(proc, img4)
    job4 = D.Job(proc, work4, W, rows4[1] - rows4[0], 256, opts=opts)
    if world > 1:
        job4.set_shard(W, H, rows4[0])
    barrier()
    c4 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    c4[0].record()
    job4.init()
    c4[1].record()
    job4.step(16)
    c4[2].record()
    barrier()
    c4_ms = [c4[0].elapsed_time(c4[1]), c4[1].elapsed_time(c4[2])]
    if world > 1:
        t = torch.tensor(c4_ms, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c4_ms = [float(t[0]), float(t[1])]
    job4.close()
    del work4, img4
    torch.cuda.empty_cache()
    parity = parity_check(K, D, torch, dist, proc, dev, world, rank) if world > 1 else None
    extras = run_extras(proc, K, D, torch, dev, world, rank, dist if world > 1 else None)
    extras["iteration_k8_8192_sustained"] = {
        "ms_per_pass": sus_ms, "mpix_per_s": world * n / sus_ms / 1e3, "hbm_frac_16B_per_px": n * BYTES_PER_PX / (sus_ms * 1e-3) / 1e9 / measured_peak()[0],
        "steps": 3000, "clocks_second_half": sus_clocks,
        "what": "the headline pass, 3000 further steps back to back: under sustained load the part sits at its 1 kW power cap "
                "(sw_power_cap) and the SM clock settles below the 1965 MHz of a short run; the pass is half arithmetic-bound, so it slows with the clock"}
    extras["iteration_k256_8192"] = {"mpix_per_s_per_gpu": n / ms256 / 1e3, "ms_per_pass": ms256,
                                     "exact_path_pixels_per_pass": st256["slow_pixels"] / max(st256["passes"], 1),
                                     "what": "8192x8192 blobs(512) k=256 assign+update pass on one GPU (config 4 without the all-reduce)"}
    extras["config4_8192_k256_sharded"] = {"n_gpus": world, "init_ms": c4_ms[0], "passes16_ms": c4_ms[1],
                                           "mpix_per_s_per_iteration": W * H * 16 / c4_ms[1] / 1e3,
                                           "what": "one 8192x8192 blobs(512) image, k=256, rows sharded over the GPUs (strong scaling): "
                                                   "255 farthest-point rounds, then 16 assign+update passes; arg-max, colour and "
                                                   "k x 4 sums exchanged inside the kernels (peer mailboxes) when N > 1"}

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = n * BYTES_PER_PX / (step_ms * 1e-3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("lloyd_k8_8192_bytes_per_launch")
            except Exception:
                traffic = None
        # FP32 side of the roofline (BASELINE.md sections 2-3): the reference formula costs 12 + 20 k
        # single operations per pixel; the peak is measured on this GPU (FFMA issue rate)
        fma_peak = D.fp32_peak(proc)
        ops_px = 12 + 20 * K_CLUSTERS
        warp_inst = None
        if tp.exists():
            try:
                warp_inst = json.loads(tp.read_text()).get("lloyd_k8_8192_warp_instructions_per_launch")
            except Exception:
                warp_inst = None
        fp32_roof = {"bound": "fp32", "unit": "Top/s (an FMA counts as one operation)", "ops_per_px_reference_formula": ops_px,
                     "achieved": n * ops_px / (step_ms * 1e-3) / 1e12, "peak": fma_peak / 1e12,
                     "frac": n * ops_px / (step_ms * 1e-3) / fma_peak, "peak_source": "measured on this GPU (kmg_dev_fp32_peak)",
                     "note": "above 1 because the pass evaluates the reference's 20 operations per (pixel, centroid) as 5 FMAs "
                             "plus a certificate; the binding resource is instruction issue under register-bank limits "
                             "(DESIGN.md 4.6)",
                     "issue_frac": (warp_inst * 32 / (step_ms * 1e-3) / fma_peak) if warp_inst else None,
                     "issue_frac_what": "static instruction count: executed warp instructions of one launch (ncu capture in "
                                        "profiles/traffic.json, not measured in this run) x 32 lanes / time / peak"}
        if world == 1:
            cpu_mpix, cpu_threads, cpu_sec = cpu_iteration_sample(2048, 3)
            cpu_baseline = {"value": cpu_mpix, "unit": "Mpix/s", "cores": cpu_threads, "kind": "port",
                            "sample": "2048x2048 crop of the same synthetic image, 3 iterations, oracle/oracle.cpp "
                                      "(restated reference on the host CPU; not wgpu/lavapipe)"}
        else:
            cpu_baseline = None  # reported at N=1 only
        line = {
            "metric": METRIC, "value": world * n * args.steps / (total_ms * 1e-3) / 1e6, "unit": "Mpix/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{W}x{H} blobs({BLOBS}) k={K_CLUSTERS} Lab assign+update iteration per GPU"
                                   + (f", rows sharded over {world} GPUs, k x 4 int64 sums exchanged per pass "
                                      + ("inside the pass kernel through peer-mapped mailboxes (NVLink)" if comm_mode == 2
                                         else "by an NCCL all-reduce") if world > 1 else ""),
                       "k": K_CLUSTERS, "bytes_per_px": BYTES_PER_PX, "l2": "work plane 1 GiB per GPU > 126 MB L2",
                       "exact_path_pixels_per_pass": stats["slow_pixels"] / max(stats["passes"], 1)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "static: profiles/traffic.json (ncu --set full capture of this kernel "
                         "on this shape, committed with the profile; not measured in this run)", "peak_source": peak_src, "kernel": "k_lloyd_ring<KT=8, 8 consumer warps + 1 TMA producer warp, 4 px per lane and stage, ring of 4 x 16 KiB, "
                                   "2 blocks/SM, table resident in uniform registers> (kmg_lloyd_ring.cuh)",
                         "kernel_ms": step_ms},
            "roofline_fp32": fp32_roof,
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_mpix, "unit": "Mpix/s", "h2d_bytes_per_step": n * 4, "d2h_bytes_per_step": K_CLUSTERS * 16 + 64,
                    "call": f"kmg_kmeans_palette(max_dim=0, {E2E_PASSES} passes) on a pinned host image", "steps": e2e_steps,
                    "call_ms": e2e_s / e2e_steps * 1e3, "images_per_s": world * e2e_steps / e2e_s,
                    "note": f"one call = one 268 MB upload + conversion + 7 init rounds + {E2E_PASSES} passes + read-back; "
                            "Mpix/s counts every pass, images/s counts calls"},
            "gpu_launches": int(launches),
            "host_binding": numa,
            "parity_check": parity,
            "clocks": clocks,
            "extras": extras,
        }
        print(json.dumps(line))
    proc.close()
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not (parity["ranks_identical"] and parity["equals_single_gpu"]):
        raise SystemExit("bench.py: multi-GPU parity check failed: " + json.dumps(parity))


def run_extras(proc, K, D, torch, dev, world=1, rank=0, dist=None):
    """The other BASELINE configs, bounded to a few seconds each.  Device-side numbers are CUDA-event
    timed on the launching stream; end-to-end numbers are wall clock around the public host call with
    page-locked buffers (H2D + kernels + D2H inside).  With N ranks every rank runs its own share and
    the slowest rank's time counts."""
    import oracle_lib as O
    from PIL import Image as PILImage

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])
        return x

    def ev_ms(fn, reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    out = {}
    # ---- configs 1-2: tokyo reduce -c 8 -m dither + palette -c 8, host image in, host image out ----
    tokyo = np.array(PILImage.open(ROOT / "tests" / "golden" / "tokyo.png").convert("RGBA"))
    tk = K.pinned_empty(tokyo.shape)
    tk[...] = tokyo
    tk_out = K.pinned_empty(tokyo.shape)
    for _ in range(3):
        proc.reduce(8, tk, reduce_mode=K.ReduceMode.Dither, out=tk_out)
        proc.palette(8, tk)
    reps = 50
    t0 = time.perf_counter()
    for _ in range(reps):
        proc.reduce(8, tk, reduce_mode=K.ReduceMode.Dither, out=tk_out)
    t_reduce = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        proc.palette(8, tk)
    t_pal = (time.perf_counter() - t0) / reps
    # the reference's own concurrency pattern: 14 host threads share one ImageProcessor
    # (core/examples/parallel.rs:23,36-51); every thread reduces with its own pinned buffers
    n_thr, per_thr = 14, 30
    bufs = []
    for _ in range(n_thr):
        bi, bo = K.pinned_empty(tokyo.shape), K.pinned_empty(tokyo.shape)
        bi[...] = tokyo
        bufs.append((bi, bo))

    def worker(i):
        bi, bo = bufs[i]
        for _ in range(per_thr):
            proc.reduce(8, bi, reduce_mode=K.ReduceMode.Dither, out=bo)

    for rnd in range(2):  # first round warms the per-thread workspaces up
        ths = [threading.Thread(target=worker, args=(i,)) for i in range(n_thr)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        t_par = time.perf_counter() - t0
    entry = {"images_per_s": 1.0 / (t_reduce + t_pal), "reduce_dither_us": t_reduce * 1e6, "palette_us": t_pal * 1e6,
             "reduce_dither_14_threads_images_per_s": n_thr * per_thr / t_par,
             "h2d_bytes": int(tokyo.nbytes) * 2, "d2h_bytes": int(tokyo.nbytes) + 8 * 16,
             "launches_per_reduce": 2, "what": "ImageProcessor.reduce(8, dither) + .palette(8), pinned host buffers"}
    if rank == 0 and world == 1:
        t0 = time.perf_counter()
        O.reduce(tokyo, 8, "dither")
        O.palette(tokyo, 8)
        entry["cpu_oracle_images_per_s"] = 1.0 / (time.perf_counter() - t0)
        entry["cpu_oracle_threads"] = O.num_threads()
    out["tokyo_reduce_dither_plus_palette_c8"] = entry

    # ---- config 3: find, 64-colour resurrect palette, dither, synthetic 4K (uniform colours) --------
    pal = K.parse_palette(ROOT / "tests" / "golden" / "resurrect_64.png")
    cent64 = K.fixed_centroids(pal)
    w4, h4 = 3840, 2160
    img4 = D.synth(proc, w4 * h4, seed=1, blobs=0, device=dev).view(h4, w4, 4)
    out4 = torch.empty_like(img4)
    tiny = torch.empty((8, 4), dtype=torch.float32, device=dev)
    job64 = D.Job(proc, tiny, 8, 1, 64)
    job64.set_centroids(cent64)
    for _ in range(3):
        job64.remap(img4, K.ReduceMode.Dither, out=out4)
    ms = ev_ms(lambda: job64.remap(img4, K.ReduceMode.Dither, out=out4), 20)
    h4in = K.pinned_empty((h4, w4, 4))
    h4in[...] = img4.cpu().numpy()
    h4out = K.pinned_empty((h4, w4, 4))
    proc.find(h4in, pal, K.ReduceMode.Dither, out=h4out)
    t0 = time.perf_counter()
    for _ in range(5):
        proc.find(h4in, pal, K.ReduceMode.Dither, out=h4out)
    t_find = (time.perf_counter() - t0) / 5
    out["find_dither_resurrect64_4k"] = {"mpix_per_s_resident": w4 * h4 / ms / 1e3, "kernel_ms": ms,
                                         "gb_per_s_8B_per_px": 8 * w4 * h4 / ms / 1e6,
                                         "e2e_images_per_s": 1.0 / t_find, "e2e_mpix_per_s": w4 * h4 / t_find / 1e6}
    job64.close()

    # ---- config 5 as BASELINE states it: 4096 resident 1920x1080 frames, k=16 reduce + dither, the frames
    # sharded over the N GPUs (strong scaling; no collective) -----------------------------------------
    total_frames = 4096
    lo, hi = K.frame_shards(total_frames, world)[rank]
    nf = hi - lo
    frames = torch.empty((nf, 1080, 1920, 4), dtype=torch.uint8, device=dev)
    for f in range(nf):
        D.synth(proc, 1920 * 1080, frame=lo + f, seed=3, blobs=32, out=frames[f].view(-1, 4))
    outb = torch.empty_like(frames)
    D.reduce_batch(proc, frames, 16, K.ReduceMode.Dither, out=outb)  # warm-up at full size (workspace allocation, first cluster launch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    _, _, passes = D.reduce_batch(proc, frames, 16, K.ReduceMode.Dither, out=outb)
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    entry = {"images_per_s": total_frames / dt, "frames_total": total_frames, "frames_per_gpu": nf, "n_gpus": world,
             "scaling": "strong", "seconds": dt, "mean_passes": float(passes.mean()), "launches": 2,
             "resident_bytes_per_gpu": int(frames.numel()) * 2,
             "what": "BASELINE config 5: 4096 synthetic 1080p frames resident in HBM, contiguous frame ranges per GPU "
                     "(frame_shards), kmg_dev_reduce_batch: one cluster launch + one remap launch for the rank's frames"}
    hn = 256
    hin = K.pinned_empty((hn, 1080, 1920, 4))
    hin[...] = frames[:hn].cpu().numpy()
    hout = K.pinned_empty((hn, 1080, 1920, 4))
    del frames, outb
    proc.reduce_batch(16, hin[:32], K.ReduceMode.Dither, out=hout[:32])
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    proc.reduce_batch(16, hin, K.ReduceMode.Dither, out=hout)
    dt = max_over_ranks(time.perf_counter() - t0)
    entry["e2e_images_per_s"] = world * hn / dt
    entry["e2e_bytes_per_frame_each_way"] = 1920 * 1080 * 4
    entry["e2e_gb_per_s_each_way_per_gpu"] = hn * 1920 * 1080 * 4 / dt / 1e9
    entry["e2e_what"] = f"kmg_reduce_batch on {hn} pinned host frames per GPU: chunked H2D / kernels / D2H pipeline"
    out["frames_1080p_k16_reduce_dither"] = entry

    # ---- what the host side can deliver: plain page-locked copies, all ranks at once, no kernels ------
    # (the ceiling of every end-to-end number above: kmg_reduce_batch moves 8.3 MB per frame each way)
    nbytes = 256 << 20
    hp_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    hp_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dv_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dv_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def copies(up, down, reps=4):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s_up):
                    dv_in.copy_(hp_in, non_blocking=True)
            if down:
                with torch.cuda.stream(s_dn):
                    hp_out.copy_(dv_out, non_blocking=True)
        torch.cuda.synchronize()
        return reps * nbytes / max_over_ranks(time.perf_counter() - t0) / 1e9

    copies(True, True, 1)
    probe = {"h2d_alone_gb_per_s_per_gpu": copies(True, False), "d2h_alone_gb_per_s_per_gpu": copies(False, True),
             "both_ways_gb_per_s_each_way_per_gpu": copies(True, True), "n_gpus": world,
             "what": "256 MiB page-locked copies on every rank at the same time, slowest rank counts; multiply by n_gpus "
                     "for what the host's memory and PCIe fabric deliver in total"}
    out["host_copy_ceiling"] = probe
    del hp_in, hp_out, dv_in, dv_out
    return out


def main():
    # the CPU legs use every host core, also under torchrun (which exports OMP_NUM_THREADS=1);
    # libgomp reads the variable when the oracle library is loaded
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1
    # (NCCL prints its version banner there when NCCL_DEBUG is set on the box) go to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(json_fd, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
