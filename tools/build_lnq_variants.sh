#!/bin/bash
# Build libeffocr_b200 variants that differ only in the fused norm1 + QKV kernel's compile-time knobs (lnqkv_sm100.cuh).
set -e
cd "$(dirname "$0")/.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -DEFFOCR_BUILDING=1"
OBJS=$(ls effocr_b200/build/*.o | grep -v "/gemm.o")
build() { # name, defines...
  name=$1; shift
  nvcc $FLAGS "$@" -I include -c effocr_b200/csrc/gemm.cu -o /tmp/gemm_$name.o
  nvcc -shared -o effocr_b200/_variants/lib_$name.so /tmp/gemm_$name.o $OBJS -gencode arch=compute_100a,code=sm_100a -cudart static -Xlinker --no-undefined
  echo built $name
}
for v in "$@"; do
  case $v in
    A) build A -DLNQ_BN=192 -DLNQ_X_SLOTS=4 -DLNQ_ILP=2 & ;;
    B) build B -DLNQ_BN=192 -DLNQ_X_SLOTS=3 -DLNQ_ILP=2 & ;;
    D) build D -DLNQ_BN=128 -DLNQ_X_SLOTS=3 -DLNQ_ILP=2 & ;;
    F) build F -DLNQ_BN=128 -DLNQ_X_SLOTS=4 -DLNQ_ILP=2 & ;;
  esac
done
wait
