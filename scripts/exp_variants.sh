#!/bin/bash
# Developer tool: build libssdr_b200 variants with SSDR_EXP=<mask> into build/exp/ (run here, on the CPU box).
set -e
cd "$(dirname "$0")/../supersdr_b200/csrc"
mkdir -p ../../build/exp
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC"
for m in "$@"; do
  nvcc $FLAGS -fmad=false -DSSDR_EXP=$m $EXTRA -c wf_kernels.cu -o ../../build/exp/wf_$m.o &
done
wait
for m in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/exp/libssdr_exp$m.so ../../build/csrc/capi.o ../../build/exp/wf_$m.o ../../build/csrc/demod_kernels.o ../../build/csrc/demod_tc_kernels.o ../../build/csrc/misc_kernels.o ../../build/csrc/nccl_comm.o -ldl
done
ls -la ../../build/exp/*.so
