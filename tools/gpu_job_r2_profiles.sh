# One gpurun call that refreshes what profiles/r02_* is built from (tools/make_profile_summary.py turns the
# reports into the tracked summaries on the CPU box).
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lloyd -s 3 -c 1 -f -o gpurun_out/r02_prof_lloyd_k8 python tools/prof_lloyd.py 8 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lloyd -s 2 -c 1 -f -o gpurun_out/r02_prof_lloyd_k256 python tools/prof_lloyd.py 256 8192 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_init_lazy -c 1 -f -o gpurun_out/r02_prof_init_lazy python tools/time_init.py 4096 256 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_convert -s 1 -c 1 -f -o gpurun_out/r02_prof_convert python tools/time_convert.py > /dev/null 2>&1
python tools/time_init.py > gpurun_out/r02_time_init.log 2>&1
python tools/audit_margin.py 8 > gpurun_out/r02_audit_margin.log 2>&1
cat gpurun_out/r02_bench_n1.json | head -c 1500; tail -3 gpurun_out/r02_bench_n1.err; cat gpurun_out/r02_bench_reference_arm.json | head -c 600
