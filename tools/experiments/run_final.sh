#!/bin/bash
# GPU box: the round's captures - GPU tests, bench line, ncu launch list of the same command, one full ncu capture of the
# sweep kernel and of the setup kernels
mkdir -p gpurun_out
T=${1:-r1h}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 600 gpurun_out/bench_$T.err
HP_NO_COOP=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-tts > gpurun_out/ncu_bench_$T.log 2>&1
HP_NO_COOP=1 ncu --set full --clock-control none --import-source on -k regex:hp_sweep4_kernel -s 1 -c 1 -o gpurun_out/sweep4_$T \
    python tools/ncu_sweep.py 4096 12 600 > gpurun_out/ncu4_$T.log 2>&1
HP_NO_COOP=1 ncu --set full --clock-control none --import-source on -k regex:'hp_(leaf_warp|chain_reg|corner_warp|sep_chain_half)' -c 4 \
    -o gpurun_out/setup_$T python tools/ncu_sweep.py 4096 12 300 > gpurun_out/ncu_setup_$T.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
cat gpurun_out/bench_$T.json | head -c 3000
