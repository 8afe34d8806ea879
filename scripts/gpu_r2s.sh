#!/bin/bash
# staged kernels for 2048 / 4096 points: parity, then sweep rows; EVERY command under its own timeout
timeout 300 python -m pytest tests/test_gpu_waterfall.py -m gpu -q -x -k "staged" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_waterfall.py tests/test_gpu_bench_shapes.py tests/test_gpu_dropin.py -m gpu -q -x -k "not demod" 2>&1 | tail -3
for st in 1 0; do
  echo "SSDR_WF_STAGED=$st"
  SSDR_WF_STAGED=$st timeout 120 python scripts/sweep.py --sizes 2048,4096 --batches 4096,32768 --n-avg 10 | cut -c1-150
done
for s in 1 200 300; do echo -n "staged, stagger $s: "; SSDR_WF_STAGGER=$s timeout 120 python scripts/sweep.py --sizes 4096 --batches 4096 --n-avg 10 | cut -c1-150; done
