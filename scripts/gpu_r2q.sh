#!/bin/bash
# TMA-staged 16384-point kernel: parity, bench against the direct-load kernel, stagger sweep, phase timeline.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_waterfall.py tests/test_gpu_bench_shapes.py -m gpu -q -x -k "not demod" 2>&1 | tail -3
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-demod --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], d['roofline']['frac'])
    else: print(l.rstrip()[:300])
"; }
for s in ${STAGGERS:-1 250 500 750}; do echo -n "staged stagger $s: "; SSDR_WF_STAGGER=$s run; done
echo -n "direct: "; SSDR_WF_STAGED=0 run
SSDR_B200_LIB=$PWD/build/exp/libssdr_trace.so timeout 120 python scripts/wf_trace.py
