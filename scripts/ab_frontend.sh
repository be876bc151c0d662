#!/bin/bash
# A/B of compile-time variants of the fast front-end on the GPU box: each argument is a set of -D flags.
cd "$(dirname "$0")/.."
B=kaldi_tflite_b200
for FLAGS in "$@"; do
  echo "=== variant: $FLAGS"
  nvcc $FLAGS -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
      -c $B/csrc/frontend_r16.cu -o $B/build/frontend_r16.o > /dev/null 2>&1 || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $B/lib/libktf_b200.so $B/build/*.o || exit 1
  [ -z "$NOPREC" ] && python scripts/fe_precision.py 2>&1 | tail -2
  [ -z "$NOTIME" ] && python scripts/quick_time.py 1024 2>&1 | tail -1
done
