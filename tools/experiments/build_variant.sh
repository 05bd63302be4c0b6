#!/bin/bash
# developer helper: build a variant of the library with extra -D flags for one source (SRC, default hp_sweep4)
#   SRC=hp_sweep4d tools/experiments/build_variant.sh <tag> [-DHP4D_GATE=0 ...]   ->  tools/experiments/_libs/lib_<tag>.so
set -e
cd "$(dirname "$0")/../.."
tag=$1; shift
SRC=${SRC:-hp_sweep4}
C=helmholtz_preconditioner_b200/csrc
mkdir -p tools/experiments/_libs
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --fmad=true "$@" -c $C/$SRC.cu -o /tmp/${SRC}_$tag.o
objs=""
for f in hp_api hp_assembly hp_blas hp_setup hp_front hp_front_coupled hp_sweep hp_sweep2 hp_sweep4 hp_sweep4m hp_sweep4d; do
  if [ $f != $SRC ]; then objs="$objs $C/$f.o"; fi
done
nvcc -shared -o tools/experiments/_libs/lib_$tag.so $objs /tmp/${SRC}_$tag.o -gencode arch=compute_100a,code=sm_100a -lcudart_static -ldl -lrt -lpthread
echo built lib_$tag.so
