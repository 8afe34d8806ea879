#!/bin/bash
# round 2, call E: full GPU suite, bench (demod lines), ncu of the tcgen05 demodulator after the plane-pitch fix
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest_all.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2e_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2e_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_tc -s 3 -c 1 -o gpurun_out/r2e_demod_tc python scripts/demod_modes.py --modes usb --frames 64 --iters 1 > gpurun_out/r2e_ncu.log 2>&1
tail -8 gpurun_out/r2e_pytest_all.log
python - <<'PY'
import json
for l in open("gpurun_out/r2e_bench.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("wf ms", d["ms_per_step"], "e2e", d["e2e"]["value"] if d.get("e2e") else None)
        for k, v in (d.get("demod") or {}).items():
            print(k, v.get("engine"), {e: (round(x["value"] / 1e3, 1), x.get("pcm_checksum")) for e, x in v.get("engines", {}).items()}, v.get("error"))
PY
tail -3 gpurun_out/r2e_ncu.log
