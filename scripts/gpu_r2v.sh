#!/bin/bash
# tcgen05 demodulator: tile stagger sweep on BASELINE configs 3 and 4 (device-resident), then the parity tests
for s in ${STAGGERS:-0 1000 2000 3000}; do
  echo "== SSDR_TC_STAGGER=$s"
  SSDR_TC_STAGGER=$s timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        for k, v in (d.get('demod') or {}).items():
            print(' ', k, {e: (round(x['value'] / 1e3, 1), x.get('pcm_checksum_ok')) for e, x in v['engines'].items()})
"
done
timeout 900 python -m pytest tests/test_gpu_audio.py tests/test_gpu_bench_shapes.py -m gpu -q -x 2>&1 | tail -3
