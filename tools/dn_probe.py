import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import fastad_b200 as F, ctypes as C
F.load(); F.check(F.lib().fadb_init(0)); F.use_torch_stream()
nd = int(sys.argv[1])
rng = np.random.default_rng(nd)
Lo = np.tril(rng.uniform(-1, 1, (nd, nd)), -1) + 10 * np.eye(nd)
Sg = torch.from_numpy(Lo @ Lo.T).cuda()
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
z = lambda n: torch.zeros(n, dtype=torch.float64, device="cuda")
xd, md, xa, ma, Sa = dev(rng.uniform(-1, 1, nd)), dev(rng.uniform(-1, 1, nd)), z(nd), z(nd), z(nd * nd)
for _ in range(2):
    print(F.normal_dense(xd, md, Sg, xa, ma, Sa))
