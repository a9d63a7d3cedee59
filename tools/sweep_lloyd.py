"""Time the variants of the thread-private Lloyd pass (KMG_LLOYD8_VARIANT / KMG_LLOYD16_VARIANT /
KMG_LLOYD32_VARIANT, see LLOYD_VARIANTS in kmg_api.cu) on the bench workload and check that they all
produce the same integer sums.  One subprocess per variant.
usage: sweep_lloyd.py <k: 8|16|32> [variants, e.g. 0,1,2] [side] [passes]"""
import hashlib
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def child(k: int, side: int, passes: int):
    sys.path.insert(0, str(ROOT))
    import torch
    import kmeans_gpu_b200 as K
    import kmeans_gpu_b200.device as D
    proc = K.ImageProcessor(0)
    img = D.synth(proc, side * side, seed=2, blobs=2 * k).view(side, side, 4)
    work = D.convert(proc, img)
    job = D.Job(proc, work, side, side, k, opts=K.Opts(max_dim=0, max_iter=1 << 30, check_every=0))
    job.init()
    job.step(5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    job.step(passes)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / passes
    sums = job.sums()
    print(json.dumps({"k": k, "variant": int(os.environ.get("KMG_LLOYD%d_VARIANT" % k, "-1")), "ms_per_pass": ms,
                      "gbps_16B": side * side * 16 / ms / 1e6, "sums_sha": hashlib.sha1(sums.tobytes()).hexdigest()[:12],
                      "stats": job.stats()}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    else:
        k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
        variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(range(4))
        side = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
        passes = int(sys.argv[4]) if len(sys.argv) > 4 else 50
        for v in variants:
            env = dict(os.environ)
            env["KMG_LLOYD%d_VARIANT" % k] = str(v)
            r = subprocess.run([sys.executable, __file__, "--child", str(k), str(side), str(passes)], env=env,
                               capture_output=True, text=True)
            print(r.stdout.strip() or ("variant %d failed: " % v + r.stderr.strip()[-400:]), flush=True)
