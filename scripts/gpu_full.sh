#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (ours + reference arm), micro-benchmark, ncu launch list,
# ncu --set full captures of the hot kernels (waterfall, demodulator on both FIR engines).  Usage: scripts/gpu_full.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_$TAG.log 2>&1
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_$TAG.log 2>&1
[ -x scripts/ubench/fp32_pipes ] && timeout 120 ./scripts/ubench/fp32_pipes > gpurun_out/ubench_fp32_$TAG.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wf_fft_kernel -s 3 -c 1 \
    -o gpurun_out/prof_wf_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-demod > gpurun_out/ncu_wf_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:demod_kernel -s 3 -c 1 \
    -o gpurun_out/prof_demod_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_demod_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:demod_tc_kernel -s 3 -c 1 \
    -o gpurun_out/prof_demod_tc_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_demod_tc_$TAG.log 2>&1
tail -8 gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/smoke_$TAG.log; cat gpurun_out/bench_$TAG.log; cat gpurun_out/bench_ref_$TAG.log; cat gpurun_out/ubench_fp32_$TAG.txt
