#!/bin/bash
# fast atan2 of the NBFM detector: parity (both engines, float64 oracle at 1e-5) + per-mode throughput + bench lines; all under timeouts
timeout 600 python -m pytest tests/test_gpu_audio.py tests/test_gpu_bench_shapes.py -m gpu -q 2>&1 | tail -6
timeout 300 python scripts/demod_modes.py --modes usb,am,nbfm 2>&1 | cut -c1-130
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        for k, v in (d.get('demod') or {}).items():
            print(' ', k, {e: (round(x['value'] / 1e3, 1), x.get('pcm_checksum'), x.get('pcm_checksum_ok')) for e, x in v['engines'].items()})
"
