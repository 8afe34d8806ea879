#!/bin/bash
# developer helper: product library + the -DSSDR_TRACE variant (build/exp/libssdr_trace.so), in parallel
set -e
cd "$(dirname "$0")/../supersdr_b200/csrc"
mkdir -p ../../build/exp
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -fmad=false"
( nvcc $F -DSSDR_TRACE $EXTRA -c wf_kernels.cu -o ../../build/exp/wf_trace.o 2>&1 | grep -iE "error" || true ) &
make 2>&1 | grep -iE "error|warning: v" || true
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/exp/libssdr_trace.so ../../build/csrc/capi.o ../../build/exp/wf_trace.o ../../build/csrc/demod_kernels.o ../../build/csrc/demod_tc_kernels.o ../../build/csrc/misc_kernels.o ../../build/csrc/nccl_comm.o -ldl
ls -la ../libssdr_b200.so ../../build/exp/libssdr_trace.so
