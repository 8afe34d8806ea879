#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wf_fft_kernel -s 3 -c 1 \
    -o gpurun_out/prof_wf_r1f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-demod > gpurun_out/ncu_wf.log 2>&1
tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log
