"""torchrun --nproc-per-node N tools/sharded_gate_run.py [--soak K] [--fault-rank R] [--collective p2p|nccl]
Runs tests/sharded_gate.py (exactly-representable sharded evaluations, bit-identity across ranks, exchange health) on N
GPUs, through the fused peer-memory exchange AND through partial -> mailbox all-reduce -> finish, then a soak of K
mailbox all-reduces with alternating payload sizes, each checked bit for bit against the rank-ordered sum.
--fault-rank R makes rank R leave out one exchange: every rank must then FAIL (non-zero exit), never print OK."""
import argparse, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch, torch.distributed as dist
import fastad_b200 as F
from fastad_b200 import sharding
import sharded_gate

ap = argparse.ArgumentParser()
ap.add_argument("--soak", type=int, default=10000)
ap.add_argument("--fault-rank", type=int, default=-1)
ap.add_argument("--collective", default="p2p")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
F.load(); F.check(F.lib().fadb_init(local)); F.use_torch_stream()
if a.collective == "p2p":
    pa = sharding.PeerAllReduce(world, rank)
    modes = ((pa, True, "fused exchange"), (pa, False, "partial + mailbox all-reduce + finish"))
else:
    pa = None
    modes = ((lambda t: dist.all_reduce(t), False, "partial + NCCL all-reduce + finish"),)
fault = a.fault_rank if a.fault_rank >= 0 else None
for ar, fused, name in modes:
    done = sharded_gate.run(F, world, rank, ar, fused, skip_exchange_on_rank=fault)
    if rank == 0:
        print(f"gate ok ({name}): {done}", flush=True)
if pa is not None and a.soak > 0:
    # soak: alternating payload sizes, bit-identity with the rank-ordered sum on every call (checked in batches)
    sizes = (1, 4, 26, 65, 128, 3)
    g = torch.Generator(device="cuda").manual_seed(99 + rank)
    pool = torch.randn(64, 128, dtype=torch.float64, device="cuda", generator=g)
    allp = [torch.empty_like(pool) for _ in range(world)]
    dist.all_gather(allp, pool)
    want = allp[0].clone()
    for r in range(1, world):
        want += allp[r]                          # rank-ordered sum, the order the mailbox kernel uses
    bad = 0
    for it in range(a.soak):
        n = sizes[it % len(sizes)]
        row = it % 64
        buf = pool[row, :n].clone()
        pa(buf)
        if not torch.equal(buf, want[row, :n]):
            bad += 1
    torch.cuda.synchronize()
    assert bad == 0 and pa.healthy(), f"soak: {bad} of {a.soak} exchanges differ from the rank-ordered sum"
    if rank == 0:
        print(f"soak ok: {a.soak} mailbox all-reduces, alternating sizes {sizes}, all bit-identical, healthy", flush=True)
dist.barrier()
if pa is not None:
    pa.close()
dist.destroy_process_group()
if rank == 0:
    print("SHARDED GATE OK", flush=True)
