#!/bin/bash
# Run under gpurun from the repo root: regenerates the evidence under gpurun_out/ that tools/ncu_summary.py and
# tools/launch_agg.py condense into profiles/.  Every ncu command is preceded by the same command without ncu.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference_arm.json 2> gpurun_out/r1_bench_ref.err
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_launch.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r1_launches_full_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
python bench.py --quick --no-cpu --steps 20 --warmup 3 > gpurun_out/plain_bs.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bs_kernel -s 5 -c 2 -o gpurun_out/prof_bs \
    python bench.py --quick --no-cpu --steps 20 --warmup 3 > gpurun_out/ncu_bs.log 2>&1
python tools/jit_probe.py chain 5 > gpurun_out/plain_jit.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fadb_stage_jit -s 3 -c 2 -o gpurun_out/prof_jit_chain \
    python tools/jit_probe.py chain 5 > gpurun_out/ncu_jit.log 2>&1
python bench.py --workloads mlp3 --no-cpu --steps 5 --warmup 3 > gpurun_out/plain_mlp3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp3_kernel -s 12 -c 2 -o gpurun_out/prof_mlp3 \
    python bench.py --workloads mlp3 --no-cpu --steps 5 --warmup 3 > gpurun_out/ncu_mlp3.log 2>&1
python bench.py --workloads linreg,expr --no-cpu --steps 5 --warmup 3 > gpurun_out/plain_lin.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:linreg_tma -s 3 -c 2 -o gpurun_out/prof_linreg \
    python bench.py --workloads linreg,expr --no-cpu --steps 5 --warmup 3 > gpurun_out/ncu_lin.log 2>&1
ls -la gpurun_out/*.ncu-rep
