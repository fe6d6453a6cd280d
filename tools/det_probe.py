"""det / log_det (FullPivLU and LDLT policies: the Gauss-Jordan path of csrc/dense_normal.cu) at the given sizes: ms per value +
adjoint and a digest of the results.  FADB_GJ_SMALL=0 keeps the launch chain for n <= 448 (A/B and bit-identity runs)."""
import hashlib, os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import fastad_b200 as F
F.load(); L = F.lib(); F.check(L.fadb_init(0)); F.use_torch_stream()
sizes = [int(a) for a in sys.argv[1:]] or [4, 64, 300]
for n in sizes:
    rng = np.random.default_rng(n)
    A = rng.uniform(-1, 1, (n, n)) + 0.5 * np.eye(n)
    Lo = np.tril(rng.uniform(-1, 1, (n, n)), -1) / np.sqrt(n) + 1.1 * np.eye(n)
    for policy, M in (("fullpivlu", A), ("ldlt", Lo @ Lo.T)):
        dA = torch.from_numpy(np.ascontiguousarray(M.T)).cuda()
        adj = torch.zeros(n * n, dtype=torch.float64, device="cuda")
        v = F.det(dA, adj, seed=0.7, policy=policy, log=True)
        torch.cuda.synchronize()
        print("digest", policy, n, repr(v), hashlib.sha256(adj.cpu().numpy().tobytes()).hexdigest()[:16])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): F.det(dA, adj, seed=0.7, policy=policy, log=True)
        e1.record(); torch.cuda.synchronize()
        print("time", policy, n, "ms/call %.4f" % (e0.elapsed_time(e1) / 5))
