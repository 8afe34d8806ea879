#!/bin/bash
# round 2, call J: 16384-point kernel with direct first-pass twiddles from tensor memory -- parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_waterfall.py tests/test_gpu_bench_shapes.py -m gpu -q -k "not demod" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -6 gpurun_out/r2j_pytest.log
timeout 300 python bench.py --no-demod --no-e2e --no-cpu-baseline > gpurun_out/r2j_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2j_bench.log
tail -c 900 gpurun_out/r2j_bench.log
timeout 300 python scripts/sweep.py --sizes 32768,65536 --batches 1024 --n-avg 10 > gpurun_out/r2j_sweep_big.jsonl 2>&1
SSDR_WF_BIG=fused timeout 300 python scripts/sweep.py --sizes 32768,65536 --batches 1024 --n-avg 10 >> gpurun_out/r2j_sweep_big.jsonl 2>&1
cat gpurun_out/r2j_sweep_big.jsonl | cut -c1-220
