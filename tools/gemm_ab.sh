#!/bin/bash
# A/B of the FP64 GEMM tile on the tensor cores (mma.sync m8n8k4.f64, DMMA) against the register-tiled DFMA kernel:
# the 1024^3 DotNode (forward + both adjoints) and the dense-covariance normal at n = 1000 / 2000 / 4000.  Run under gpurun.
for D in 0 1 0 1; do
    echo "== FADB_GEMM_DMMA=$D"
    FADB_GEMM_DMMA=$D python tools/dot_probe.py 14 64 2>&1 | grep "matmul"
    FADB_GEMM_DMMA=$D python bench.py --workloads dense --no-cpu --steps 5 --warmup 3 2>/dev/null | python -c '
import sys, json
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])
print(" ".join("%s=%.3fms(%.1fTF)" % (k[-5:], v["ms_per_step"], v["fp64_gflops"] / 1e3) for k, v in d["workloads"].items() if "dense" in k))'
done
