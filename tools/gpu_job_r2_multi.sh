# multi-GPU evidence: parity of row-sharded jobs against one GPU (both exchange modes), the multi-GPU
# pytest file, and the bench line with its parity_check.  usage: gpurun --gpus N -- bash tools/gpu_job_r2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
$TR tools/check_multi_gpu.py > gpurun_out/r02_check_multi_gpu_n${N}_p2p.log 2>&1; echo "check p2p rc=$?" >> gpurun_out/r02_check_multi_gpu_n${N}_p2p.log
KMG_NO_P2P=1 $TR tools/check_multi_gpu.py > gpurun_out/r02_check_multi_gpu_n${N}_nccl.log 2>&1; echo "check nccl rc=$?" >> gpurun_out/r02_check_multi_gpu_n${N}_nccl.log
python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_test_multi_gpu_n${N}.log
$TR bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err; echo "bench rc=$?" >> gpurun_out/r02_bench_n${N}.err
grep -h "OK\|MISMATCH\|rc=" gpurun_out/r02_check_multi_gpu_n${N}_p2p.log gpurun_out/r02_check_multi_gpu_n${N}_nccl.log | tail -20
cat gpurun_out/r02_test_multi_gpu_n${N}.log
tail -2 gpurun_out/r02_bench_n${N}.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_n${N}.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "parity_check")})
print(d["roofline"]["frac"], d["e2e"]["value"], d["extras"]["config4_8192_k256_sharded"], d["extras"]["frames_1080p_k16_reduce_dither"]["images_per_s"], d["extras"]["frames_1080p_k16_reduce_dither"]["e2e_images_per_s"])
PY
