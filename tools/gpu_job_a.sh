set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gpu_tests.log
cat gpurun_out/gpu_tests.log
python tools/prof_remap.py 64 1 3840 2160 > gpurun_out/remap.log 2>&1
python tools/prof_remap.py 64 0 3840 2160 >> gpurun_out/remap.log 2>&1
python tools/prof_remap.py 16 1 3840 2160 32 >> gpurun_out/remap.log 2>&1
python tools/prof_remap.py 8 1 3840 2160 16 >> gpurun_out/remap.log 2>&1
cat gpurun_out/remap.log
python - <<'PY'
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, kmeans_gpu_b200 as K, oracle_lib
proc = K.ImageProcessor(0)
w, h = 3840, 2160
img = K.pinned_empty((h, w, 4)); img[...] = oracle_lib.synth(w * h, seed=1).reshape(h, w, 4)
out = K.pinned_empty((h, w, 4))
cols = K.parse_palette('tests/golden/resurrect_64.png')
for mode in (K.ReduceMode.Dither, K.ReduceMode.Replace):
    for _ in range(3): proc.find(img, cols, mode, out=out)
    t0 = time.perf_counter()
    for _ in range(20): proc.find(img, cols, mode, out=out)
    dt = (time.perf_counter() - t0) / 20
    print(mode, "find 4K k=64 e2e: %.3f ms  %.1f images/s" % (dt * 1e3, 1 / dt))
PY
