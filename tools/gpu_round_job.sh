# One gpurun call that refreshes everything profiles/ is built from (see tools/make_profile_summary.py).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/gpu_tests.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lloyd -s 3 -c 1 -f -o gpurun_out/prof_lloyd_k8 python tools/prof_lloyd.py 8 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lloyd -s 2 -c 1 -f -o gpurun_out/prof_lloyd_k256 python tools/prof_lloyd.py 256 8192 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lloyd -s 2 -c 1 -f -o gpurun_out/prof_lloyd_k64 python tools/prof_lloyd.py 64 8192 4 > /dev/null 2>&1
if [ "$FULL" = "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:k_kmeans_small -s 2 -c 1 -f -o gpurun_out/prof_small_tokyo python tools/prof_small.py tokyo > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_kmeans_small -s 1 -c 1 -f -o gpurun_out/prof_small_batch python tools/prof_small.py batch 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_remap -s 1 -c 1 -f -o gpurun_out/prof_remap_batch python tools/prof_small.py batch 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_convert -s 1 -c 1 -f -o gpurun_out/prof_convert python tools/time_convert.py > /dev/null 2>&1
fi
cat gpurun_out/gpu_tests.log; cat gpurun_out/bench_n1.json; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_n1.err
