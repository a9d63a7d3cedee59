set -x
mkdir -p gpurun_out
python tools/sweep_lloyd8.py ${SWEEP:-0,8,9,10} 8192 50 > gpurun_out/sweep_lloyd8.log 2>&1
cat gpurun_out/sweep_lloyd8.log
KMG_LLOYD8_VARIANT=${TESTV:-8} timeout 600 python -m pytest tests -m gpu -x -q -k "not 16m and not all_16m" 2>&1 | tail -5 > gpurun_out/gpu_tests_v.log
cat gpurun_out/gpu_tests_v.log
