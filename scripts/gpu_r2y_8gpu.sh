#!/bin/bash
# round 2, multi-GPU refresh with every staged size (one 8-GPU box): bench at N = 8, config-5 sweep on 8 GPUs and on 1 GPU
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-demod > gpurun_out/r2y_bench_8gpu.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2y_bench_8gpu.log
timeout 300 python scripts/sweep.py --gpus 8 --sizes 256,512,1024,2048,4096,8192,16384,65536 --batches 4096 --n-avg 10 --max-bytes 3e10 > gpurun_out/r2y_sweep_8gpu.jsonl 2>&1
timeout 200 python scripts/sweep.py --gpus 8 --sizes 256,512,1024 --batches 65536 --n-avg 10 --max-bytes 3e10 >> gpurun_out/r2y_sweep_8gpu.jsonl 2>&1
timeout 200 python scripts/sweep.py --sizes 256,512,1024,2048,4096,8192,16384 --batches 4096,65536 --n-avg 10 --max-bytes 9e9 > gpurun_out/r2y_sweep_1gpu_navg10.jsonl 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2y_bench_8gpu.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["n_gpus"], round(d["value"]), d["ms_per_step"], "e2e", d["e2e"]["value"], "scatter", d.get("scatter_from_root", {}).get("value"), "peer", d.get("peer_ingest", {}).get("value"))
PY
cut -c1-200 gpurun_out/r2y_sweep_8gpu.jsonl; cut -c1-150 gpurun_out/r2y_sweep_1gpu_navg10.jsonl
