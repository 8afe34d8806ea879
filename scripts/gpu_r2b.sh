#!/bin/bash
# round 2, call B: v2 waterfall kernel -- parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_waterfall.py tests/test_gpu_bench_shapes.py -m gpu -q -x --deselect tests/test_gpu_bench_shapes.py::test_demod_config3_shape_usb > gpurun_out/r2b_pytest_wf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_wf.log
timeout 300 python bench.py --no-demod --no-e2e --no-cpu-baseline > gpurun_out/r2b_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2b_bench.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_all.log
tail -15 gpurun_out/r2b_pytest_wf.log; tail -c 1500 gpurun_out/r2b_bench.log; tail -15 gpurun_out/r2b_pytest_all.log
