set -x
mkdir -p gpurun_out
for v in 0 1; do
KMG_LLOYD8_VARIANT=$v ncu --set full --clock-control none --import-source on -k regex:k_lloyd -s 3 -c 1 -f -o gpurun_out/prof_lloyd_k8_v$v python tools/prof_lloyd.py 8 > gpurun_out/ncu_v$v.log 2>&1
tail -5 gpurun_out/ncu_v$v.log
done
ls -la gpurun_out
