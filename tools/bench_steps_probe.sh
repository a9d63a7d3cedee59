for K in 300 1000 3000; do python bench.py --steps $K --warmup 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print($K, d['roofline']['kernel_ms'], round(d['roofline']['frac'],4), d['clocks'])"; sleep 5; done
