#!/bin/bash
# GPU box: correctness of the strip setup against the oracle (small cases), then setup time at 4096^2 with the
# shared-memory chain kernel (HP_CHAIN_SMEM=1) and the register-resident one
mkdir -p gpurun_out
timeout 300 python tools/dbg_cluster.py small > gpurun_out/setup_small.log 2>&1
grep -E "err" gpurun_out/setup_small.log | awk '{print $NF, $0}' | sort -g | tail -3
grep -E "err" gpurun_out/setup_small.log | awk '{for(i=1;i<=NF;i++) if($i=="err") print $(i+1)}' | sort -g | tail -2
for v in smem reg; do
  if [ $v = smem ]; then export HP_CHAIN_SMEM=1; else unset HP_CHAIN_SMEM; fi
  timeout 200 python tools/dbg_cluster.py big 4096 cluster > gpurun_out/setup_$v.log 2>&1
  echo "== $v"; grep -E "setup ms|forward sweep|status" gpurun_out/setup_$v.log | head -3
done
unset HP_CHAIN_SMEM
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
