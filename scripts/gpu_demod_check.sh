#!/bin/bash
# Demodulator quick check (one gpurun call): parity tests of both FIR engines, per-mode throughput, BASELINE configs 3 / 4
# with the checksums bench.py prints (for re-pinning bench.DEMOD_CHECKSUMS after an arithmetic change).  Usage: scripts/gpu_demod_check.sh <tag>
TAG=${1:-dq}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_audio.py tests/test_gpu_bench_shapes.py tests/test_gpu_dropin.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/demod_check_$TAG.log
python scripts/demod_modes.py --modes usb,am,nbfm 2>&1 | tee -a gpurun_out/demod_check_$TAG.log
python scripts/demod_modes.py --modes usb --hang 1 2>&1 | tee -a gpurun_out/demod_check_$TAG.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        for k, v in (d.get('demod') or {}).items():
            print(' ', k, {e: (round(x['value'] / 1e3, 1), x.get('pcm_checksum'), x.get('pcm_checksum_ok')) for e, x in v['engines'].items()})
" | tee -a gpurun_out/demod_check_$TAG.log
