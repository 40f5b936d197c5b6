#!/bin/bash
# [SRC=routing_bwd] mkvariant.sh NAME [nvcc -D flags...]: one translation unit (default routing_fused.cu) rebuilt with the flags, linked with the other objects of
# the last full build -> gpurun_variants/lib_NAME.so (load with SRF_B200_LIB=...)
set -e
name=$1; shift
src=${SRC:-routing_fused}
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c -o /tmp/rf_$name.o srf_b200/csrc/$src.cu
objs=$(ls srf_b200/_build/*.o | grep -v $src.o)
mkdir -p gpurun_variants
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_variants/lib_$name.so /tmp/rf_$name.o $objs
echo gpurun_variants/lib_$name.so
