import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import kmeans_gpu_b200 as K, kmeans_gpu_b200.device as D, oracle_lib as O
proc = K.ImageProcessor(0)
v = np.arange(1 << 24, dtype=np.uint32)
px = np.stack([v & 255, (v >> 8) & 255, (v >> 16) & 255, np.full_like(v, 255)], axis=1).astype(np.uint8)
work = D.convert(proc, torch.from_numpy(px).cuda()).cpu().numpy()
want = O.convert(px)
diff = work[:, :3].view(np.uint32) != want[:, :3].view(np.uint32)
rows = np.nonzero(diff.any(axis=1))[0]
print("mismatching colours", len(rows), "components", diff.sum(), "per comp", diff.sum(axis=0))
np.save(ROOT / "gpurun_out" / "mismatch_rows.npy", rows)
np.save(ROOT / "gpurun_out" / "mismatch_gpu.npy", work[rows])
np.save(ROOT / "gpurun_out" / "mismatch_cpu.npy", want[rows])
for r in rows[:12]:
    print(px[r], work[r, :3], want[r, :3], (work[r, :3].view(np.int32) - want[r, :3].view(np.int32)))
