#!/bin/bash
# GPU box: correctness of the first variant against the oracle, then the 4096^2 sweep timing and phase sums of every variant
mkdir -p gpurun_out
L=tools/experiments/_libs
first=$1
HELMHOLTZ_B200_LIB=$L/lib_$first.so timeout 300 python tools/dbg_cluster.py small > gpurun_out/dbg_small_$first.log 2>&1
for tag in "$@"; do
  HP_TAG=_$tag HELMHOLTZ_B200_LIB=$L/lib_$tag.so timeout 200 python tools/dbg_cluster.py big 4096 cluster > gpurun_out/dbg_$tag.log 2>&1
  echo "== $tag"; grep -E "forward sweep|precond apply|status" gpurun_out/dbg_$tag.log | head -4
done
grep -E "err" gpurun_out/dbg_small_$first.log | awk '{print $NF, $0}' | sort -g | tail -3
