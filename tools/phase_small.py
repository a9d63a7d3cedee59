"""Phase split of k_kmeans_small on tokyo: vary k and max_iter (run under ncu --metrics gpu__time_duration.sum)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import kmeans_gpu_b200 as K
from PIL import Image as PILImage
proc = K.ImageProcessor(0)
tokyo = np.array(PILImage.open(ROOT / "tests" / "golden" / "tokyo.png").convert("RGBA"))
small = proc.resize(tokyo, 256).rgba
for img, name in ((tokyo, "tokyo"), (small, "preshrunk")):
    for k, mi in ((1, 1), (1, 17), (8, 1), (8, 17), (8, 33), (16, 1), (16, 17)):
        cent, passes = proc.kmeans_centroids(k, img, opts=K.Opts(max_iter=mi, check_every=0))
        print(name, "k", k, "max_iter", mi, "passes", passes)
