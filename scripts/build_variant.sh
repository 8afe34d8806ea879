#!/bin/bash
# developer helper: build/exp/libssdr_<name>.so with extra nvcc flags for wf_kernels.cu.  usage: build_variant.sh <name> <flags...>
set -e
name=$1; shift
cd "$(dirname "$0")/../supersdr_b200/csrc"
mkdir -p ../../build/exp
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -fmad=false "$@" -c wf_kernels.cu -o ../../build/exp/wf_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/exp/libssdr_$name.so ../../build/csrc/capi.o ../../build/exp/wf_$name.o ../../build/csrc/demod_kernels.o ../../build/csrc/demod_tc_kernels.o ../../build/csrc/misc_kernels.o ../../build/csrc/nccl_comm.o -ldl
