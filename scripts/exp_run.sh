#!/bin/bash
# Developer tool (GPU box): time every build/exp variant on config 2.
mkdir -p gpurun_out
for so in build/exp/libssdr_exp*.so; do
  echo -n "$(basename $so) " >> gpurun_out/exp.log
  SSDR_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-demod --no-e2e 2>&1 | python -c "import sys,json; l=[x for x in sys.stdin if x.startswith('{')]; print(json.loads(l[-1])['ms_per_step'] if l else 'FAILED')" >> gpurun_out/exp.log
done
cat gpurun_out/exp.log
