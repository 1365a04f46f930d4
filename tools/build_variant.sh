#!/bin/bash
# Development aid: builds build/variants/libsipoc_<tag>.so with extra nvcc flags on
# riccati_fast.cu (e.g. -DSIPOC_FUSED_PF=3) so that several settings of one kernel can be
# timed back to back in one GPU session: SIPOC_LIB_PATH=build/variants/libsipoc_<tag>.so python bench.py ...
set -e
TAG=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT"
mkdir -p build/variants build/obj
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC -std=c++17 -O3 -lineinfo $ARCH -ccbin /usr/bin/g++ -Xcompiler -fPIC -cudart static --expt-relaxed-constexpr \
  "$@" -c sip_optimal_control_b200/csrc/riccati_fast.cu -o build/variants/riccati_fast_$TAG.o
OBJS=$(ls build/obj/*.o | grep -v "/riccati_fast.cu.o")
$NVCC $ARCH -ccbin /usr/bin/g++ -shared -cudart static -o build/variants/libsipoc_$TAG.so $OBJS build/variants/riccati_fast_$TAG.o -ldl
echo build/variants/libsipoc_$TAG.so
