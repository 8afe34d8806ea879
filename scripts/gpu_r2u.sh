#!/bin/bash
# Round-2 final evidence (one gpurun call): full GPU suite, smoke, bench both arms, launch list, ncu of the three hot kernels, memcheck.
TAG=${1:-r2z}
bash scripts/gpu_full.sh $TAG > gpurun_out/full_$TAG.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py > gpurun_out/sanitize_memcheck_$TAG.log 2>&1
echo "== memcheck rc=$?"; grep -E "ERROR SUMMARY|Error" gpurun_out/sanitize_memcheck_$TAG.log | head -5
python scripts/sweep.py --sizes 256,1024,4096,16384,65536 --batches 4096 --n-avg 10 > gpurun_out/sweep_navg10_$TAG.jsonl 2>&1
python scripts/sweep.py --sizes 256,512,1024,2048,4096,8192,16384,32768,65536 --batches 1,16,256,4096,65536 --n-avg 1 > gpurun_out/sweep_navg1_$TAG.jsonl 2>&1
tail -30 gpurun_out/full_$TAG.log | cut -c1-400
cat gpurun_out/sweep_navg10_$TAG.jsonl
