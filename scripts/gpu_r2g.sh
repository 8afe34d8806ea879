#!/bin/bash
# round 2, call G: fused large-N kernel with the overlapped front pass (parity + sweep)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_waterfall.py -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -6 gpurun_out/r2g_pytest.log
timeout 300 python scripts/sweep.py --sizes 32768,65536 --batches 1024,2048 --n-avg 10 > gpurun_out/r2g_sweep_big_fused.jsonl 2>&1
timeout 300 python scripts/sweep.py --sizes 32768,65536 --batches 4096 --n-avg 1 >> gpurun_out/r2g_sweep_big_fused.jsonl 2>&1
cat gpurun_out/r2g_sweep_big_fused.jsonl | cut -c1-200
