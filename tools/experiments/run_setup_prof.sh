#!/bin/bash
# GPU box: launch list of the strip setup at 4096^2 (all strips) and one full ncu capture of the leaf and chain kernels
mkdir -p gpurun_out
HP_NO_COOP=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/setup_launches_r1b.csv \
   -k regex:'hp_(chain|leaf|corner|sep|mleaf|rsep|wg)' python tools/ncu_sweep.py 4096 12 > gpurun_out/setup_ncu1.log 2>&1
HP_NO_COOP=1 ncu --set full --clock-control none --import-source on -k regex:'hp_(leaf_fast|chain_reg|corner|sep_chain)' -c 4 \
   -o gpurun_out/setup_r1b python tools/ncu_sweep.py 4096 12 300 > gpurun_out/setup_ncu2.log 2>&1
ls -la gpurun_out/setup_r1b.ncu-rep
