set -x
mkdir -p gpurun_out
(python tools/sweep_lloyd.py 8 0,6,7,8 8192 50; python tools/sweep_lloyd.py 16 0,4,5 8192 30; python tools/sweep_lloyd.py 32 0,2 8192 20) > gpurun_out/sweep_lloyd.log 2>&1
cat gpurun_out/sweep_lloyd.log
for c in 0 1; do
KMG_LLOYDG_RING=$c python tools/prof_lloyd.py 256 8192 6
KMG_LLOYDG_RING=$c python tools/prof_lloyd.py 64 8192 6
done > gpurun_out/lloydg.log 2>&1
cat gpurun_out/lloydg.log
