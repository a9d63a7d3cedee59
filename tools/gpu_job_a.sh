set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 300 --warmup 5 > gpurun_out/b1.json 2> gpurun_out/b1.err; python -c "
import json; d=json.load(open('gpurun_out/b1.json')); print(d['value'], d['roofline']['frac'], d['e2e'], d['extras']['config4_8192_k256_sharded']['init_ms'])"; tail -2 gpurun_out/b1.err
