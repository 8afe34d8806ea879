#!/bin/bash
# SASS reuse-peephole experiment: parity tests + bench on the patched library against the ptxas-scheduled one.
mkdir -p gpurun_out
for lib in libssdr_b200.so libssdr_b200_reuse.so; do
  echo "== $lib"
  SSDR_B200_LIB=$PWD/supersdr_b200/$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-demod --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d.get('clocks'))
    else: print(l.rstrip()[:300])
"
done
SSDR_B200_LIB=$PWD/supersdr_b200/libssdr_b200_reuse.so timeout 900 python -m pytest tests/test_gpu_waterfall.py -m gpu -q -x 2>&1 | tail -5
