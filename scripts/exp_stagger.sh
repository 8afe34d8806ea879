#!/bin/bash
# developer experiment: stagger of the warp-local passes of the 16384-point kernel (cycles per warp-group level)
mkdir -p gpurun_out
for s in 1 200 300 400 500 600 700 900; do
  echo -n "stagger $s: " >> gpurun_out/exp_stagger.txt
  SSDR_WF_STAGGER=$s timeout 120 python bench.py --no-demod --no-e2e --no-cpu-baseline --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print(round(d['ms_per_step'],4), 'ms', round(d['roofline']['frac'],4))" >> gpurun_out/exp_stagger.txt
done
cat gpurun_out/exp_stagger.txt
