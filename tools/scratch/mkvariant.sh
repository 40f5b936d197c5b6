#!/bin/bash
# mkvariant.sh NAME [nvcc -D flags...]: routing_fused.cu rebuilt with the flags, linked with the other objects of
# the last full build -> tools/scratch/lib_NAME.so (load with SRF_B200_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/../.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c -o /tmp/rf_$name.o srf_b200/csrc/routing_fused.cu
objs=$(ls srf_b200/_build/*.o | grep -v routing_fused.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/scratch/lib_$name.so /tmp/rf_$name.o $objs
echo tools/scratch/lib_$name.so
