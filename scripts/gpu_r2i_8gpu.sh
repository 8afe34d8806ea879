#!/bin/bash
# round 2, call I (8 GPUs): bench at N = 8 and N = 4 under torchrun (e2e scaling with the bare-copy ceilings), config-5 sweep on 8 / 4 GPUs
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2i_topo.txt 2>&1; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" >> gpurun_out/r2i_topo.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-demod > gpurun_out/r2i_bench_8gpu.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2i_bench_8gpu.log
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 10 --warmup 3 --no-demod > gpurun_out/r2i_bench_4gpu.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2i_bench_4gpu.log
timeout 600 python scripts/sweep.py --gpus 8 --sizes 256,1024,4096,16384,65536 --batches 4096 --n-avg 10 --max-bytes 3e10 > gpurun_out/r2i_sweep_8gpu.jsonl 2>&1
timeout 600 python scripts/sweep.py --gpus 4 --sizes 256,1024,4096,16384,65536 --batches 4096 --n-avg 10 --max-bytes 3e10 > gpurun_out/r2i_sweep_4gpu.jsonl 2>&1
timeout 300 python scripts/sweep.py --gpus 8 --sizes 256,1024,4096,16384 --batches 65536 --n-avg 1 --max-bytes 3e10 >> gpurun_out/r2i_sweep_8gpu.jsonl 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2i_bench_8gpu.log", "gpurun_out/r2i_bench_4gpu.log"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["n_gpus"], round(d["value"]), d["ms_per_step"], "e2e", d["e2e"], "scatter", d.get("scatter_from_root"), "peer", d.get("peer_ingest"))
PY
cat gpurun_out/r2i_sweep_8gpu.jsonl | cut -c1-200; cat gpurun_out/r2i_topo.txt | head -30
