#!/bin/bash
# build libs4mpi from the in-tree sources into $1 (A/B kernel builds; extra nvcc flags after the output path)
out=$1; shift
exec nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared "$@" \
    -o "$out" /root/repo/streets4mpi_b200/csrc/s4_capi.cu
