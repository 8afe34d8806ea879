#!/bin/bash
# compute-sanitizer over small invocations of every kernel.  Usage: scripts/gpu_sanitize.sh <tag>
TAG=${1:-s}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/gpu_sanitize.py > gpurun_out/sanitize_${tool}_$TAG.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error" gpurun_out/sanitize_${tool}_$TAG.log | head -8
done
