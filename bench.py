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
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
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
                "reasons": sorted(reasons), "samples": len(sm)}


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
        "config": {"workload": f"{W}x{H} blobs({BLOBS}) k={K_CLUSTERS} Lab assign+update iteration", "k": K_CLUSTERS},
        "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": threads, "kind": "port",
                         "sample": f"{side}x{side} crop of the same synthetic image, {steps} iterations, "
                                   "oracle/oracle.cpp (restated reference; Rust+wgpu reference not buildable here)"},
        "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    proc = K.ImageProcessor(local)

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
        job.set_shard(W, H * world, rank * H)
    job.init()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        job.step(1)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = proc.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for a, b in evs:
        a.record()
        job.step(1)
        b.record()
    t_end.record()
    barrier()
    launches = proc.launch_count() - launches0
    total_ms = t_start.elapsed_time(t_end)
    step_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([total_ms, step_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, step_ms = float(t[0]), float(t[1])
    stats = job.stats()

    # ---- end to end through the host C ABI (pinned host image) --------------------------------
    host = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
    host.copy_(img)
    torch.cuda.synchronize()
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
        extras = run_extras(proc, K, D, torch, dev)
        cpu_mpix, cpu_threads, cpu_sec = cpu_iteration_sample(2048, 3)
        line = {
            "metric": METRIC, "value": world * n * args.steps / (total_ms * 1e-3) / 1e6, "unit": "Mpix/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{W}x{H} blobs({BLOBS}) k={K_CLUSTERS} Lab assign+update iteration per GPU"
                                   + (f", rows sharded over {world} GPUs with per-pass NCCL all-reduce of k x 4 int64 sums" if world > 1 else ""),
                       "k": K_CLUSTERS, "bytes_per_px": BYTES_PER_PX, "l2": "work plane 1 GiB per GPU > 126 MB L2",
                       "exact_path_pixels_per_pass": stats["slow_pixels"] / max(stats["passes"], 1)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "k_lloyd<8,8,256,4,true,2>",
                         "kernel_ms": step_ms},
            "cpu_baseline": {"value": cpu_mpix, "unit": "Mpix/s", "cores": cpu_threads, "kind": "port",
                             "sample": "2048x2048 crop of the same synthetic image, 3 iterations, oracle/oracle.cpp "
                                       "(restated reference on the host CPU; not wgpu/lavapipe)"},
            "e2e": {"value": e2e_mpix, "unit": "Mpix/s", "h2d_bytes_per_step": n * 4, "d2h_bytes_per_step": K_CLUSTERS * 16 + 64,
                    "call": f"kmg_kmeans_palette(max_dim=0, {E2E_PASSES} passes) on a pinned host image", "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "extras": extras,
        }
        print(json.dumps(line))
    job.close()
    proc.close()
    if world > 1:
        dist.destroy_process_group()


def run_extras(proc, K, D, torch, dev):
    """End-to-end images/s through the host API for the small configs (bounded, a few seconds)."""
    import oracle_lib as O
    from PIL import Image as PILImage

    out = {}
    tokyo = np.array(PILImage.open(ROOT / "tests" / "golden" / "tokyo.png").convert("RGBA"))
    tk = torch.from_numpy(tokyo).pin_memory().numpy()
    for _ in range(2):
        proc.reduce(8, tk, reduce_mode=K.ReduceMode.Dither)
        proc.palette(8, tk)
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        proc.reduce(8, tk, reduce_mode=K.ReduceMode.Dither)
        proc.palette(8, tk)
    dt = (time.perf_counter() - t0) / reps
    out["tokyo_reduce_dither_plus_palette_c8"] = {"images_per_s": 1.0 / dt, "ms": dt * 1e3, "h2d_bytes": int(tokyo.nbytes) * 2}
    t0 = time.perf_counter()
    O.reduce(tokyo, 8, "dither")
    O.palette(tokyo, 8)
    out["tokyo_reduce_dither_plus_palette_c8"]["cpu_oracle_images_per_s"] = 1.0 / (time.perf_counter() - t0)
    # config 5: 1080p frames, k=16 reduce+dither, frames resident in HBM
    nf = 16
    frames = torch.empty((nf, 1080, 1920, 4), dtype=torch.uint8, device=dev)
    for f in range(nf):
        D.synth(proc, 1920 * 1080, frame=f, seed=3, blobs=32, out=frames[f].view(-1, 4))
    outb = torch.empty_like(frames)
    D.reduce_batch(proc, frames[:2], 16, K.ReduceMode.Dither, out=outb[:2])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    D.reduce_batch(proc, frames, 16, K.ReduceMode.Dither, out=outb)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["frames_1080p_k16_reduce_dither_resident"] = {"images_per_s": nf / dt, "frames": nf}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
