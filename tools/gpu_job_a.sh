set -x
mkdir -p gpurun_out
python tools/prof_init.py 64 8192
python tools/prof_init.py 256 8192
python tools/prof_init.py 8 8192
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
