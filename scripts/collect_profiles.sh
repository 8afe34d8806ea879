#!/bin/bash
# After a scripts/gpu_r2u.sh <tag> call: copy the evidence from gpurun_out/ (scratch) into profiles/ (tracked) and
# summarise the ncu captures.  Usage: scripts/collect_profiles.sh <tag>     (runs here, no GPU)
set -e
TAG=${1:?tag}
cd "$(dirname "$0")/.."
cp gpurun_out/bench_$TAG.log profiles/${TAG}_bench.log
cp gpurun_out/bench_ref_$TAG.log profiles/${TAG}_bench_ref.log
cp gpurun_out/launches_$TAG.csv profiles/${TAG}_launches.csv
cp gpurun_out/pytest_gpu_$TAG.log profiles/${TAG}_pytest_gpu.log
cp gpurun_out/smoke_$TAG.log profiles/${TAG}_smoke.log
cp gpurun_out/sanitize_memcheck_$TAG.log profiles/${TAG}_sanitize_memcheck.log
cp gpurun_out/sweep_navg10_$TAG.jsonl profiles/${TAG}_sweep_navg10.jsonl
cp gpurun_out/sweep_navg1_$TAG.jsonl profiles/${TAG}_sweep_navg1.jsonl
[ -f gpurun_out/demod_modes_$TAG.jsonl ] && cp gpurun_out/demod_modes_$TAG.jsonl profiles/${TAG}_demod_modes.jsonl
python scripts/ncu_summary.py gpurun_out/prof_wf_$TAG.ncu-rep > profiles/${TAG}_wf_ncu_summary.txt
python scripts/ncu_summary.py gpurun_out/prof_demod_$TAG.ncu-rep > profiles/${TAG}_demod_ncu_summary.txt
python scripts/ncu_summary.py gpurun_out/prof_demod_tc_$TAG.ncu-rep > profiles/${TAG}_demod_tc_ncu_summary.txt
python scripts/sass_opcodes.py > profiles/${TAG}_sass_opcodes.txt
ls -la profiles/${TAG}_*
