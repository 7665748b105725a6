#!/bin/bash
# run_variants.sh OUT NAME... — time every build/variants/NAME library with tools/exp_variant.py on C3 (2^24 box rays) and C2 (2^23 interior
# rays); one JSON line per run into OUT (hit-record CRCs included, so variants can be checked for bit-identical results).
out=$1; shift
: > $out
for v in "$@"; do
  RAYCORE_CUDA_LIB=build/variants/$v/libraycore_cuda.so python tools/exp_variant.py --log2rays 24 >> $out 2>> $out.err
  [ -n "$SKIP_C2" ] || RAYCORE_CUDA_LIB=build/variants/$v/libraycore_cuda.so python tools/exp_variant.py --c2 --log2rays 23 >> $out 2>> $out.err
done
