#!/bin/bash
# build_variant.sh NAME [-DFLAG=VALUE ...] — library variant with rc_trace.cu compiled under extra defines (the other objects come from the
# default build): build/variants/NAME/libraycore_cuda.so, for tools/exp_variant.py (RAYCORE_CUDA_LIB=...).  Prints the registers / spills
# of the multi-instance closest_hit kernel.
set -e
cd "$(dirname "$0")/../raycore.jl_b200/csrc"
name=$1; shift
out=../../build/variants/$name
mkdir -p $out
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-Wall -Xptxas -v --expt-relaxed-constexpr "$@" -c rc_trace.cu -o $out/rc_trace.o 2> $out/rc_trace.ptxas.log || (cat $out/rc_trace.ptxas.log; false)
$NVCC -shared $ARCH -o $out/libraycore_cuda.so rc_api.o rc_build.o $out/rc_trace.o rc_analysis.o rc_collide.o rc_wavefront.o rc_multi.o
echo "$name: $(grep -A2 'k_trace_wideILb0ELb0E10RcIoArraysLb0' $out/rc_trace.ptxas.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')"
