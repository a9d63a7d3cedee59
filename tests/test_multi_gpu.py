"""Row-sharded k-means over several GPUs (BASELINE config 4) against the single-GPU result, both
through the in-kernel peer exchange and through the NCCL all-reduce.  Needs >= 2 GPUs (skipped on
the single-GPU box); the CPU-side logic of the sharding is covered by test_sharding_gloo.py."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _gpus() -> int:
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("no_p2p", [False, True])
def test_sharded_job_matches_single_gpu(no_p2p):
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ)
    env.pop("KMG_NO_P2P", None)
    if no_p2p:
        env["KMG_NO_P2P"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533" if no_p2p else "29532", str(ROOT / "tools" / "check_multi_gpu.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MISMATCH" not in r.stdout
