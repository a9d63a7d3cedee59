"""How much of the certificate's error bound is ever used?  All 2^24 colours x palettes of several
shapes; prints, per role, the number of pixels whose fast arg-min is not the reference label and the
largest (score gap / eps) among them.  usage: audit_margin.py [palettes per k]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import kmeans_gpu_b200 as K, kmeans_gpu_b200.device as D
from test_gpu_parity import _audit_palettes
n_pal = int(sys.argv[1]) if len(sys.argv) > 1 else 8
proc = K.ImageProcessor(0)
v = torch.arange(1 << 24, dtype=torch.int32, device="cuda")
img = (v | (255 << 24)).view(torch.uint8).view(4096, 4096, 4)
work = D.convert(proc, img)
rng = np.random.default_rng(77)
worst = {0: 0.0, 1: 0.0, 2: 0.0}
for k in (2, 8, 16, 64, 256, 700):
    for cent in _audit_palettes(K, rng, k, n_pal):
        for mode in (0, 1, 2):
            differ, ratio = D.audit(proc, cent, 4, mode, work=work if mode == 0 else None, rgba=img if mode else None, w=4096, h=4096)
            worst[mode] = max(worst[mode], ratio / 1e6)
    print(f"k={k}: largest gap/eps so far: lloyd {worst[0]:.4f}  replace {worst[1]:.4f}  dither {worst[2]:.4f}", flush=True)
