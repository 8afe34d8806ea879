#!/bin/bash
# Developer tool (GPU box): time the demodulator engines of every build/exp variant (configs 3 and 4).
mkdir -p gpurun_out
for so in build/exp/libssdr_exp*.so; do
  echo -n "$(basename $so) " >> gpurun_out/exp_tc.log
  SSDR_B200_LIB=$PWD/$so timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
l = [x for x in sys.stdin if x.startswith('{')]
d = json.loads(l[-1])['demod'] if l else {}
print({k: {e: round(x['value'] / 1e3, 1) for e, x in v['engines'].items()} for k, v in d.items()} or 'FAILED')" >> gpurun_out/exp_tc.log
done
cat gpurun_out/exp_tc.log
