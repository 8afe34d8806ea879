#!/bin/bash
# A/B of quantiser group sizes (device-resident, config 2), then the waterfall parity tests on the product library
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-demod --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], d['roofline']['frac'])
    else: print(l.rstrip()[:300])
"; }
echo -n "groups of 16 (product): "; run
echo -n "groups of 8: "; SSDR_B200_LIB=$PWD/build/exp/libssdr_q8.so run
echo -n "groups of 32: "; SSDR_B200_LIB=$PWD/build/exp/libssdr_q32.so run
echo -n "direct-load kernel, groups of 16: "; SSDR_WF_STAGED=0 run
timeout 900 python -m pytest tests/test_gpu_waterfall.py -m gpu -q -x 2>&1 | tail -2
