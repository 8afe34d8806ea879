#!/bin/bash
# Quick GPU check: parity tests + bench without the CPU baseline.  Usage: scripts/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ${2:+-k "$2"} ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_$TAG.log 2>&1
tail -15 gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/bench_$TAG.log
