#!/bin/bash
# ncu --set full capture of the waterfall kernel on config 2.  Usage: scripts/gpu_ncu_wf.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wf_fft -s 3 -c 1 \
    -o gpurun_out/prof_wf_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-demod --no-e2e > gpurun_out/ncu_wf_$TAG.log 2>&1
tail -3 gpurun_out/ncu_wf_$TAG.log
