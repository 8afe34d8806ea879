#!/bin/bash
# A/B of small variants of the staged 16384-point kernel (device-resident, config 2)
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-demod --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], d['roofline']['frac'])
    else: print(l.rstrip()[:300])
"; }
export SSDR_WF_STAGGER=300
echo -n "early TMEM twiddle loads: "; run
echo -n "   ... without L2 prefetch: "; SSDR_WF_STAGED_NOPF=1 run
echo -n "late TMEM twiddle loads (as committed): "; SSDR_B200_LIB=$PWD/build/exp/libssdr_tm0.so run
echo -n "   ... without L2 prefetch: "; SSDR_WF_STAGED_NOPF=1 SSDR_B200_LIB=$PWD/build/exp/libssdr_tm0.so run
timeout 600 python -m pytest tests/test_gpu_waterfall.py -m gpu -q -x 2>&1 | tail -2
