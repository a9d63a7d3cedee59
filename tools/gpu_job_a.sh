set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_kmeans_small -s 2 -c 1 -f -o gpurun_out/prof_small_tokyo python tools/prof_small.py tokyo > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_kmeans_small -s 1 -c 1 -f -o gpurun_out/prof_small_batch python tools/prof_small.py batch 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_remap -s 1 -c 1 -f -o gpurun_out/prof_remap_batch python tools/prof_small.py batch 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_convert -s 1 -c 1 -f -o gpurun_out/prof_convert python tools/time_convert.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_remap -s 3 -c 1 -f -o gpurun_out/prof_remap_k64 python tools/prof_remap.py 64 1 3840 2160 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_init_round -s 40 -c 1 -f -o gpurun_out/prof_init python tools/prof_init.py 64 8192 > /dev/null 2>&1
ls -la gpurun_out
