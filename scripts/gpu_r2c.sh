#!/bin/bash
# round 2, call C: modifier-cost microbenchmark + ncu full capture of the v2 waterfall kernel
mkdir -p gpurun_out
timeout 120 scripts/ubench/ffma2_modifiers > gpurun_out/r2c_ffma2_modifiers.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wf_fft2 -s 3 -c 1 -o gpurun_out/r2c_wf2 python bench.py --no-demod --no-e2e --no-cpu-baseline --steps 2 > gpurun_out/r2c_ncu.log 2>&1
cat gpurun_out/r2c_ffma2_modifiers.txt; tail -3 gpurun_out/r2c_ncu.log
