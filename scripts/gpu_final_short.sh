mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_r1zzz.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r1zzz.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_r1zzz.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r1zzz.log
( timeout 400 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_r1zzz.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:demod_tc_kernel -s 3 -c 1 -o gpurun_out/prof_demod_tc_r1zzz -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_demod_tc_r1zzz.log 2>&1
timeout 200 compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py > gpurun_out/sanitize_memcheck_r1zzz.log 2>&1; grep "ERROR SUMMARY" gpurun_out/sanitize_memcheck_r1zzz.log
tail -3 gpurun_out/pytest_gpu_r1zzz.log; tail -2 gpurun_out/smoke_r1zzz.log
