#!/bin/bash
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-demod --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], d['roofline']['frac'])
    else: print(l.rstrip()[:300])
"; }
timeout 900 python -m pytest tests/test_gpu_waterfall.py tests/test_gpu_dropin.py tests/test_display.py -m gpu -q -x 2>&1 | tail -3
echo -n "bench: "; run
python scripts/colorrow_bw.py
SSDR_B200_LIB=$PWD/build/exp/libssdr_trace.so python scripts/wf_trace.py 2>&1 | tail -8
