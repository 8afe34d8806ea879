#!/bin/bash
# round 2, call F: fused large-N kernel (parity + sweep vs the three-kernel path), demodulator after the plane-pitch / wait fixes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_waterfall.py tests/test_gpu_bench_shapes.py tests/test_gpu_audio.py -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -6 gpurun_out/r2f_pytest.log
timeout 300 python scripts/sweep.py --sizes 32768,65536 --batches 1024,2048 --n-avg 10 > gpurun_out/r2f_sweep_big_fused.jsonl 2>&1
SSDR_WF_BIG=3k timeout 300 python scripts/sweep.py --sizes 32768,65536 --batches 1024,2048 --n-avg 10 > gpurun_out/r2f_sweep_big_3k.jsonl 2>&1
timeout 300 python scripts/sweep.py --sizes 32768,65536 --batches 4096 --n-avg 1 >> gpurun_out/r2f_sweep_big_fused.jsonl 2>&1
SSDR_WF_BIG=3k timeout 300 python scripts/sweep.py --sizes 32768,65536 --batches 4096 --n-avg 1 >> gpurun_out/r2f_sweep_big_3k.jsonl 2>&1
cat gpurun_out/r2f_sweep_big_fused.jsonl gpurun_out/r2f_sweep_big_3k.jsonl | cut -c1-200
timeout 300 python scripts/demod_modes.py --modes usb --frames 64 > gpurun_out/r2f_demod_modes.jsonl 2>&1
timeout 300 python scripts/demod_modes.py --batch 8192 --frames 32 >> gpurun_out/r2f_demod_modes.jsonl 2>&1
cat gpurun_out/r2f_demod_modes.jsonl
