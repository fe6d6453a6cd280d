"""torchrun --nproc-per-node N tools/p2p_probe.py: the peer-memory all-reduce (csrc/p2p.cu) against NCCL:
bit-identity with the rank-ordered sum, and device time per call for the payload sizes of the sharded kernels."""
import os, sys
import torch, torch.distributed as dist
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import fastad_b200 as F
from fastad_b200 import sharding
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
F.load(); F.check(F.lib().fadb_init(local)); F.use_torch_stream()
pa = sharding.PeerAllReduce(world, rank)
ok = True
for n in (1, 4, 26, 65, 128):
    for it in range(5):
        g = torch.Generator(device="cuda").manual_seed(1000 * rank + 7 * n + it)
        t = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        want = parts[0].clone()
        for r in range(1, world): want += parts[r]
        got = pa(t.clone())
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(got, want))
print(f"rank {rank}: bit-identical to the rank-ordered sum: {ok}, healthy: {pa.healthy()}", flush=True)
def timeit(fn, reps=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for n in (4, 26, 65):
    t = torch.ones(n, dtype=torch.float64, device="cuda")
    us_p = timeit(lambda: pa(t))
    us_n = timeit(lambda: dist.all_reduce(t))
    if rank == 0: print(f"n={n:3d} doubles: peer-memory all-reduce {us_p:6.2f} us/call, NCCL {us_n:6.2f} us/call", flush=True)
# ---- the all-reduce fused into the producing kernel (fadb_sum_unary_allreduce_async / fadb_mlp3_allreduce_async):
# identical bits to plain call + mailbox all-reduce (both sum in rank order), and device time per evaluation
import ctypes as C
import numpy as np
sys.path.insert(0, os.path.join(R, "tests"))
import kats, synth
L = F.lib()
P = lambda t: C.c_void_p(t.data_ptr())
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
ok2 = True
for op, n in (("exp", 70001 + 8 * rank), ("log", 1 << 22)):
    x = dev(synth.sum_inputs(n, seed=11 + rank))
    res = []
    for fused in (False, True):
        adj, v = torch.zeros(n, dtype=torch.float64, device="cuda"), torch.zeros(1, dtype=torch.float64, device="cuda")
        call = L.fadb_sum_unary_allreduce_async if fused else L.fadb_sum_unary_async
        for _ in range(3):
            F.check(call(F._op_code(op), n, P(x), P(adj), 1.0, 0, P(v)))
            if not fused: pa(v)
        torch.cuda.synchronize()
        res.append((v.clone(), adj))
    ok2 = ok2 and torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
sms = torch.cuda.get_device_properties(local).multi_processor_count
for b in (257 + rank, 1 << 16, sms * 1024 + 5):
    Xn, yn = synth.mlp3_inputs(b)
    gX, gy, gp = torch.from_numpy(np.ascontiguousarray(Xn.T)).cuda(), dev(yn), dev(kats.NN_PARM)
    res = []
    for fused in (False, True):
        gl = torch.zeros(26, dtype=torch.float64, device="cuda")
        call = L.fadb_mlp3_allreduce_async if fused else L.fadb_mlp3_async
        for _ in range(3):
            F.check(call(b, P(gX), P(gy), P(gp), P(gl), F.OVERWRITE_ADJ, C.c_void_p(gl.data_ptr() + 200)))
            if not fused: pa(gl)
        torch.cuda.synchronize()
        res.append(gl.clone())
    ok2 = ok2 and torch.equal(res[0], res[1])
    us_s = timeit(lambda: (F.check(L.fadb_mlp3_async(b, P(gX), P(gy), P(gp), P(gl), F.OVERWRITE_ADJ, C.c_void_p(gl.data_ptr() + 200))), pa(gl)))
    us_f = timeit(lambda: F.check(L.fadb_mlp3_allreduce_async(b, P(gX), P(gy), P(gp), P(gl), F.OVERWRITE_ADJ, C.c_void_p(gl.data_ptr() + 200))))
    if rank == 0: print(f"network, {b} rows per rank: kernel + mailbox all-reduce {us_s:6.2f} us, fused {us_f:6.2f} us per evaluation", flush=True)
print(f"rank {rank}: fused sum / network all-reduce bit-identical to call + mailbox all-reduce: {ok2}, healthy: {pa.healthy()}", flush=True)
pa.close()
dist.destroy_process_group()
