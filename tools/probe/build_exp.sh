#!/bin/bash
# TEST TOOL: experiment builds of the window backward (VALOR_EXP bit 0: no dK/dV/dQ MMAs, bit 1: no element math)
set -e
cd "$(dirname "$0")/../../valor_b200/csrc"
for v in 3 7 8; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -DVALOR_EXP=$v -c window_attn_sm100.cu -o /tmp/w100_exp$v.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/probe/libvalor_exp$v.so /tmp/w100_exp$v.o $(ls build/*.o | grep -v window_attn_sm100) -lcudart_static -ldl -lrt -lpthread
done
echo built
