set -x
mkdir -p gpurun_out
for c in 0 1; do
KMG_LLOYDG_BLOCKACC=$c python tools/prof_lloyd.py 256 8192 6
KMG_LLOYDG_BLOCKACC=$c python tools/prof_lloyd.py 64 8192 6
KMG_LLOYDG_BLOCKACC=$c python tools/prof_lloyd.py 1024 4096 4
done > gpurun_out/lloydg.log 2>&1
cat gpurun_out/lloydg.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gpu_tests.log
cat gpurun_out/gpu_tests.log
