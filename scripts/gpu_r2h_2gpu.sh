#!/bin/bash
# round 2, call H (2 GPUs): the tests that need a second GPU (NCCL scatter/gather through the C ABI, handles on device 1 used
# from other threads), bench at N = 2 under torchrun, config-5 sweep on 2 GPUs
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2h_smi.txt; nvidia-smi topo -m >> gpurun_out/r2h_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -q > gpurun_out/r2h_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest_2gpu.log
tail -5 gpurun_out/r2h_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-demod > gpurun_out/r2h_bench_2gpu.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2h_bench_2gpu.log
tail -c 2500 gpurun_out/r2h_bench_2gpu.log
timeout 600 python scripts/sweep.py --gpus 2 --sizes 256,1024,4096,16384,65536 --batches 4096 --n-avg 10 --max-bytes 3e10 > gpurun_out/r2h_sweep_2gpu.jsonl 2>&1
cat gpurun_out/r2h_sweep_2gpu.jsonl | cut -c1-260
