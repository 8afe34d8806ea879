#!/bin/bash
# round 2, final multi-GPU evidence on ONE 8-GPU box: two-GPU tests, bench at N = 8 / 4 / 2 under torchrun, config-5 sweep on 8 GPUs
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2x_topo.txt 2>&1; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" >> gpurun_out/r2x_topo.txt
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -q > gpurun_out/r2x_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2x_bench_8gpu.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2x_bench_8gpu.log
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 10 --warmup 3 --no-demod > gpurun_out/r2x_bench_4gpu.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2x_bench_4gpu.log
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-demod > gpurun_out/r2x_bench_2gpu.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2x_bench_2gpu.log
timeout 600 python scripts/sweep.py --gpus 8 --sizes 256,1024,4096,16384,65536 --batches 4096 --n-avg 10 --max-bytes 3e10 > gpurun_out/r2x_sweep_8gpu.jsonl 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2x_bench_8gpu.log", "gpurun_out/r2x_bench_4gpu.log", "gpurun_out/r2x_bench_2gpu.log"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["n_gpus"], round(d["value"]), d["ms_per_step"], "e2e", {k: d["e2e"].get(k) for k in ("value", "h2d_gbs_per_rank", "h2d_ceiling_all_ranks_gbs")}, "scatter", d.get("scatter_from_root"), "peer", d.get("peer_ingest"))
            for k, v in (d.get("demod") or {}).items():
                print("   ", k, round(v["value"]), v.get("e2e", {}).get("value"))
PY
tail -3 gpurun_out/r2x_pytest_2gpu.log; cat gpurun_out/r2x_sweep_8gpu.jsonl | cut -c1-230
