import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import kmeans_gpu_b200 as K, kmeans_gpu_b200.device as D
proc = K.ImageProcessor(0)
v = torch.arange(1 << 24, dtype=torch.int32, device="cuda")
img = (v | (255 << 24)).view(torch.uint8).view(4096, 4096, 4)
work = D.convert(proc, img)
rng = np.random.default_rng(1)
for k in (2, 3, 5, 8, 12, 16, 40):
    cols = rng.integers(0, 256, (k, 4), dtype=np.uint8); cols[:, 3] = 255
    cent = K.fixed_centroids(cols, K.ColorSpace.Lab)
    for search in ([0, 3] if k <= 8 else []) + ([1] if k <= 16 else []) + [2]:
        print(k, search, [D.audit(proc, cent, search, m, work=work if m == 0 else None, rgba=img if m else None, w=4096, h=4096) for m in ((0, 1, 2) if search != 3 else (0,))])
    job = D.Job(proc, work, 4096, 4096, k, opts=K.Opts(max_dim=0, max_iter=100, check_every=0))
    job.set_centroids(cent); job.step(1); print("   production slow pixels", job.stats()); job.close()
