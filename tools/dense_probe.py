"""Dense-covariance normal at size n (default 4000): ms per value + gradient, and with FADB_DENSE_TIMING=1 the device time of
the phases of the factorisation (steps / syrk per outer panel, trtri, Y^T Y) summed per phase."""
import os, sys, re, subprocess
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
if len(sys.argv) > 2 and sys.argv[2] == "phases":
    env = dict(os.environ, FADB_DENSE_TIMING="1")
    out = subprocess.run([sys.executable, __file__, sys.argv[1]], env=env, capture_output=True, text=True)
    lines = [l for l in out.stderr.splitlines() if l.startswith("[fadb dense]")]
    tot = {}
    for name, ms in re.findall(r" ([A-Za-z()+_]+) ([0-9.]+) ms", lines[-1]):
        tot[name] = tot.get(name, 0.0) + float(ms)
    print(out.stdout.strip()); print({k: round(v, 3) for k, v in tot.items()}, "sum", round(sum(tot.values()), 3))
    sys.exit(0)
import torch, fastad_b200 as F
F.load(); L = F.lib(); F.check(L.fadb_init(0)); F.use_torch_stream()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
rng = np.random.default_rng(1)
Lo = np.tril(rng.uniform(-1, 1, (n, n)), -1) / np.sqrt(n) + 1.2 * np.eye(n)
Sig = torch.from_numpy(np.ascontiguousarray((Lo @ Lo.T).T)).cuda()
x = torch.from_numpy(rng.uniform(-1, 1, n)).cuda(); mu = torch.from_numpy(rng.uniform(-1, 1, n)).cuda()
xa, ma, Sa = (torch.zeros(k, dtype=torch.float64, device="cuda") for k in (n, n, n * n))
step = lambda: F.normal_dense(x, mu, Sig, xa, ma, Sa)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
print("dense vvm n =", n, "ms/step", e0.elapsed_time(e1) / 5)
