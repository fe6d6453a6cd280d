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
pa.close()
dist.destroy_process_group()
