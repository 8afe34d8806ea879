#!/bin/bash
# final check of the row-stage selection rule: parity + the affected rows; EVERY command under its own timeout
timeout 600 python -m pytest tests/test_gpu_waterfall.py tests/test_gpu_dropin.py tests/test_display.py tests/test_gpu_bench_shapes.py -m gpu -q -x -k "not demod" 2>&1 | tail -2
timeout 120 python scripts/sweep.py --sizes 1024 --batches 65536 --n-avg 4 --max-bytes 9e9 | cut -c1-150
timeout 120 python scripts/sweep.py --sizes 2048,4096 --batches 16384 --n-avg 10 --max-bytes 9e9 | cut -c1-150
timeout 120 python scripts/sweep.py --sizes 1024,2048,4096 --batches 4096 --n-avg 10 --max-bytes 9e9 | cut -c1-150
timeout 120 python scripts/colorrow_bw.py --shapes 16384x4096x10,8192x8192x10,4096x16384x10,2048x32768x10,1024x65536x10,1024x65536x4,1024x65536x1
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-demod --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('bench', d['ms_per_step'], d['roofline']['frac'])
"
