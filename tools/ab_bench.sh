#!/bin/bash
# A/B of two builds of libfastad_b200.so on ONE box (run under gpurun from the repo root):
#   make -C fastad_b200/csrc OUT=../lib_exp        # the experimental build next to the shipped fastad_b200/lib
#   gpurun -- 'tools/ab_bench.sh lib lib_exp [bench.py workloads, default none = headline + e2e only]'
# Alternates the builds twice (A B A B) through the FASTAD_B200_LIB override of fastad_b200.load() and prints one line
# per run: headline us per 2^20 options, roofline fraction, e2e ms, then ms_per_step of every extra workload in us.
A=${1:-lib}; B=${2:-lib_exp}; W=${3:-none}
for L in $A $B $A $B; do
    FASTAD_B200_LIB=$PWD/fastad_b200/$L/libfastad_b200.so timeout 180 python bench.py --workloads $W --no-cpu --steps 200 --warmup 10 2>/dev/null |
    L=$L python -c '
import sys, json, os
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])
print(os.environ["L"], "bs_us=%.3f frac=%.4f e2e_ms=%.4f" % (d["ms_per_step"] * 1e3, d["roofline"]["frac"], d["e2e"]["ms_per_step"]),
      " ".join("%s=%.2f" % (k, v["ms_per_step"] * 1e3) for k, v in d.get("workloads", {}).items()))'
done
