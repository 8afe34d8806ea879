#!/bin/bash
# round 2, call A: full GPU test suite (new bench-shape parity tests), TMEM scratch probe, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_gpu.log
timeout 120 scripts/ubench/tmem_scratch_probe > gpurun_out/r2a_tmem_probe.txt 2>&1
timeout 600 python bench.py > gpurun_out/r2a_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2a_bench.log
tail -5 gpurun_out/r2a_pytest_gpu.log; cat gpurun_out/r2a_tmem_probe.txt; tail -c 3000 gpurun_out/r2a_bench.log
