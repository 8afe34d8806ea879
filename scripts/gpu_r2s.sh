#!/bin/bash
# key -> colour table for warp-sized groups: parity, then the n_avg = 1 / 2 rows; EVERY command under its own timeout
timeout 600 python -m pytest tests/test_gpu_waterfall.py tests/test_gpu_dropin.py tests/test_display.py -m gpu -q -x 2>&1 | tail -3
timeout 120 python scripts/sweep.py --sizes 256,512,1024,2048 --batches 65536 --n-avg 1 --max-bytes 9e9 | cut -c1-150
timeout 120 python scripts/sweep.py --sizes 1024 --batches 65536 --n-avg 2 --max-bytes 9e9 | cut -c1-150
timeout 120 python scripts/colorrow_bw.py --shapes 1024x65536x1,1024x65536x10
