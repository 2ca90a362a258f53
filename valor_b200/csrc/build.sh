#!/bin/bash
# Builds libvalor_b200.so in-tree for sm_100a.  Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --use_fast_math -Xptxas -v"
FLAGS="${FLAGS//--use_fast_math/}"
mkdir -p build
pids=()
for f in api gemm_sm100 gemm_simt norm attention_ref attention_mma window_attn window_attn_sm100 elementwise dropout loss optim; do
  [ -f $f.cu ] || continue
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ common.cuh -nt build/$f.o ] || [ attention.cuh -nt build/$f.o ] || [ mma_utils.cuh -nt build/$f.o ] || [ ../../include/valor_b200.h -nt build/$f.o ]; then
    ( $NVCC $FLAGS "$@" -c $f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libvalor_b200.so build/*.o -lcudart_static -ldl -lrt -lpthread
echo "built $(pwd)/libvalor_b200.so"
