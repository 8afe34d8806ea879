#!/bin/bash
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-demod --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], d['roofline']['frac'])
    else: print(l.rstrip()[:300])
"; }
for s in 300 500 700; do echo -n "exp512 stagger $s: "; SSDR_B200_LIB=$PWD/build/exp/libssdr_exp512a.so SSDR_WF_STAGGER=$s run; done
echo -n "product: "; run
SSDR_B200_LIB=$PWD/build/exp/libssdr_trace512.so timeout 120 python scripts/wf_trace.py 2>&1 | head -28
