# One gpurun call that refreshes what profiles/r02_* is built from.  The ncu reports (--set full with sources:
# 15-20 MB each) are summarised on the GPU box (tools/make_profile_summary.py) and deleted there: gpurun only
# copies back 64 MiB.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1
C8="ncu --set full --clock-control none --import-source on -k regex:k_lloyd -s 3 -c 1 python tools/prof_lloyd.py 8"
ncu --set full --clock-control none --import-source on -k regex:k_lloyd -s 3 -c 1 -f -o gpurun_out/p8 python tools/prof_lloyd.py 8 > /dev/null 2>&1
python tools/make_profile_summary.py gpurun_out/p8.ncu-rep gpurun_out/r02_lloyd_k8_ncu.md "Round 2 — k_lloyd_ring (TMA ring, table in uniform registers), 8192x8192, k=8: ncu --set full" "$C8"
rm -f gpurun_out/p8.ncu-rep
C256="ncu --set full --clock-control none --import-source on -k regex:k_lloyd -s 2 -c 1 python tools/prof_lloyd.py 256 8192 4"
ncu --set full --clock-control none --import-source on -k regex:k_lloyd -s 2 -c 1 -f -o gpurun_out/p256 python tools/prof_lloyd.py 256 8192 4 > /dev/null 2>&1
python tools/make_profile_summary.py gpurun_out/p256.ncu-rep gpurun_out/r02_lloyd_k256_ncu.md "Round 2 — k_lloyd (chunked search fed from the constant bank, block accumulators), 8192x8192, k=256: ncu --set full" "$C256"
rm -f gpurun_out/p256.ncu-rep
python tools/time_init.py > gpurun_out/r02_time_init.log 2>&1
python tools/audit_margin.py 8 > gpurun_out/r02_audit_margin.log 2>&1
cat gpurun_out/r02_bench_n1.json | head -c 1500; tail -3 gpurun_out/r02_bench_n1.err; cat gpurun_out/r02_bench_reference_arm.json | head -c 600
ls -la gpurun_out
