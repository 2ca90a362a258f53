#!/bin/bash
# test tool only: builds tools/probe/libumma_probe.so (sm_100a)
set -e
cd "$(dirname "$0")"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lineinfo -Xcompiler -fPIC -shared umma_probe.cu -o libumma_probe.so -lcudart_static -ldl -lrt -lpthread
echo built $(pwd)/libumma_probe.so
