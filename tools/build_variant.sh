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
src=${VARIANT_SRC:-rc_trace}   # VARIANT_SRC=rc_build: vary the builder instead of the traversal kernels
$NVCC -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-Wall -Xptxas -v --expt-relaxed-constexpr "$@" -c $src.cu -o $out/$src.o 2> $out/$src.ptxas.log || (cat $out/$src.ptxas.log; false)
objs=""
for o in rc_api rc_build rc_trace rc_analysis rc_collide rc_wavefront rc_multi; do
  if [ $o = $src ]; then objs="$objs $out/$o.o"; else objs="$objs $o.o"; fi
done
$NVCC -shared $ARCH -o $out/libraycore_cuda.so $objs
[ $src != rc_trace ] || echo "$name: $(grep -A2 'k_trace_wideILb0ELb0E10RcIoArraysLb0' $out/rc_trace.ptxas.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')"
