set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gpu_tests.log
cat gpurun_out/gpu_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/check_multi_gpu.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -6
KMG_NO_P2P=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 tools/check_multi_gpu.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -6
