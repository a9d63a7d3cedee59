# round 2, first GPU call: parity of the new fixed-point unit + ring kernel, sweep of the k<=8 variants
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_gpu_tests.log
python tools/sweep_lloyd.py 8 0,1,7,8,9,10 8192 50 > gpurun_out/r2a_sweep8.log 2>&1
cat gpurun_out/r2a_gpu_tests.log gpurun_out/r2a_sweep8.log
