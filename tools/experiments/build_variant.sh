#!/bin/bash
# developer helper: build a variant of the library with extra -D flags for hp_sweep4.cu
#   tools/experiments/build_variant.sh <tag> [-DHP4_AHEAD=6 ...]   ->  tools/experiments/_libs/lib_<tag>.so
set -e
cd "$(dirname "$0")/../.."
tag=$1; shift
C=helmholtz_preconditioner_b200/csrc
mkdir -p tools/experiments/_libs
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --fmad=true "$@" -c $C/hp_sweep4.cu -o /tmp/hp_sweep4_$tag.o
objs=""
for f in hp_api hp_assembly hp_blas hp_setup hp_front hp_front_coupled hp_sweep hp_sweep2 hp_sweep4m hp_sweep4d; do objs="$objs $C/$f.o"; done
nvcc -shared -o tools/experiments/_libs/lib_$tag.so $objs /tmp/hp_sweep4_$tag.o -gencode arch=compute_100a,code=sm_100a -lcudart_static -ldl -lrt -lpthread
echo built lib_$tag.so
