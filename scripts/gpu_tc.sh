#!/bin/bash
# GPU check of the tcgen05 demodulator engine: parity tests of both engines (bounded), then the demodulator bench lines.
# Usage: scripts/gpu_tc.sh <tag>
TAG=${1:-tc}
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_audio.py -m gpu -q -x -k demod ) > gpurun_out/pytest_tc_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_tc_$TAG.log
tail -25 gpurun_out/pytest_tc_$TAG.log
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/bench_tc_$TAG.log 2>&1
python - <<'PY' gpurun_out/bench_tc_$TAG.log
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print("wf ms", d["ms_per_step"])
        for k, v in (d.get("demod") or {}).items():
            print(k, v["engine"], {e: round(x["value"] / 1e3, 1) for e, x in v["engines"].items()}, "Gsamples/s")
PY
tail -3 gpurun_out/bench_tc_$TAG.log | cut -c1-300
