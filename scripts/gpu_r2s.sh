#!/bin/bash
# bounded mbarrier wait: parity + bench + sweep rows; EVERY command under its own timeout
timeout 600 python -m pytest tests/test_gpu_waterfall.py tests/test_gpu_bench_shapes.py -m gpu -q -x -k "not demod" 2>&1 | tail -2
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-demod --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('bench', d['ms_per_step'], d['roofline']['frac'])
"
timeout 120 python scripts/sweep.py --sizes 512,1024 --batches 65536 --n-avg 10 --max-bytes 9e9 | cut -c1-150
timeout 120 python scripts/sweep.py --sizes 2048,4096,8192 --batches 4096 --n-avg 10 --max-bytes 9e9 | cut -c1-150
