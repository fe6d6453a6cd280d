#!/bin/bash
# Round-2 evidence (run under gpurun from the repo root): every ncu command is preceded by the same command without ncu.
set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_launch.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_launches_full_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
python bench.py --quick --no-cpu --steps 20 --warmup 3 > gpurun_out/plain_bs.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bs_kernel -s 5 -c 3 -f -o gpurun_out/r2_prof_bs \
    python bench.py --quick --no-cpu --steps 20 --warmup 3 > gpurun_out/ncu_bs.log 2>&1
ls -la gpurun_out/*.ncu-rep
