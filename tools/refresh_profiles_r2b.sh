#!/bin/bash
# Round-2 evidence of the final build (run under gpurun from the repo root, one GPU).  Every ncu command is preceded by the
# same command without ncu; the reports stay small (-c 1) and are condensed on the box with tools/ncu_stalls.py / ncu_blocks.py.
set -x
O=gpurun_out
mkdir -p $O
(time python -m pytest tests -x -q -m gpu) > $O/r2_gpu_tests.log 2>&1
python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
python bench.py --impl reference > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_smoke.log 2>&1
# launch list of the headline command
python bench.py --quick --no-cpu --no-gate --steps 20 --warmup 5 > $O/plain_headline.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_headline.csv \
    python bench.py --quick --no-cpu --no-gate --steps 20 --warmup 5 > $O/ncu_headline.log 2>&1
# the specialised stage kernel of the vectorised Black-Scholes call expression: plain form, then the operand ring
for v in plain ring; do
    if [ $v = plain ]; then export FADB_JIT_NO_RING=1; else unset FADB_JIT_NO_RING; fi
    python tools/jit_probe.py bs 20 > $O/jit_bs_$v.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:fadb_stage_jit -s 3 -c 1 -f -o $O/prof_jit_bs_$v \
        python tools/jit_probe.py bs 5 > $O/ncu_jit_bs_$v.log 2>&1
done
unset FADB_JIT_NO_RING
{
    echo "== fadb_stage_jit of the vectorised example/black_scholes.cpp call expression, 2^20 options (tools/jit_probe.py bs)"
    echo "== timing without ncu:"; tail -1 $O/jit_bs_plain.log; tail -1 $O/jit_bs_ring.log
    echo; echo "== PLAIN one-element kernel (FADB_JIT_NO_RING=1): ncu --set full --import-source on, tools/ncu_stalls.py"
    python tools/ncu_stalls.py $O/prof_jit_bs_plain.ncu-rep 24
    echo; echo "== OPERAND RING (default): tools/ncu_stalls.py"
    python tools/ncu_stalls.py $O/prof_jit_bs_ring.ncu-rep 12
    echo; echo "== OPERAND RING: executed warp instructions by SASS block (tools/ncu_blocks.py)"
    python tools/ncu_blocks.py $O/prof_jit_bs_ring.ncu-rep
} > $O/r2_ncu_stage_jit_bs_ring.txt 2>&1
python tools/jit_probe.py chain 20 > $O/jit_chain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fadb_stage_jit -s 3 -c 1 -f -o $O/prof_jit_chain \
    python tools/jit_probe.py chain 5 > $O/ncu_jit_chain.log 2>&1
{
    echo "== fadb_stage_jit of the 10-node sum chain at 2^24 (four-element form; tools/jit_probe.py chain)"; tail -1 $O/jit_chain.log
    python tools/ncu_stalls.py $O/prof_jit_chain.ncu-rep 10
} > $O/r2_ncu_stage_jit_chain.txt 2>&1
python tools/lp_probe.py bernoulli 20 > $O/lp_bernoulli.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lp_kernel -s 2 -c 1 -f -o $O/prof_lp \
    python tools/lp_probe.py bernoulli 3 > $O/ncu_lp.log 2>&1
{
    echo "== lp_kernel<BERNOULLI, vec p, early adjoint load> at 2^25 elements (tools/lp_probe.py bernoulli)"; tail -1 $O/lp_bernoulli.log
    for v in 0 5 3; do echo "FADB_LP_VARIANT=$v"; FADB_LP_VARIANT=$v python tools/lp_probe.py bernoulli 20 | tail -1; done
    python tools/ncu_stalls.py $O/prof_lp.ncu-rep 8
} > $O/r2_ncu_bernoulli.txt 2>&1
rm -f $O/*.ncu-rep
ls -la $O
