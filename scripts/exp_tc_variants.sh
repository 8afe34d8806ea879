#!/bin/bash
# Developer tool: build libssdr_b200 variants of the tcgen05 demodulator kernel into build/exp/ (run here, on the CPU box).
# Usage: scripts/exp_tc_variants.sh "name:-DFLAG=.. -DFLAG=.." ...
set -e
cd "$(dirname "$0")/../supersdr_b200/csrc"
mkdir -p ../../build/exp
rm -f ../../build/exp/libssdr_exp*.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC"
for v in "$@"; do
  n=${v%%:*}; f=${v#*:}
  ( nvcc $FLAGS $f -c demod_tc_kernels.cu -o ../../build/exp/tc_$n.o &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/exp/libssdr_exp$n.so ../../build/csrc/capi.o ../../build/csrc/wf_kernels.o \
         ../../build/csrc/demod_kernels.o ../../build/exp/tc_$n.o ../../build/csrc/misc_kernels.o ../../build/csrc/nccl_comm.o -ldl ) &
done
wait
ls -la ../../build/exp/*.so
