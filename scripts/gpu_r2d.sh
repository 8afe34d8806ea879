#!/bin/bash
# round 2, call D: full GPU suite after the delegation refactor
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_all.log
tail -30 gpurun_out/r2d_pytest_all.log
