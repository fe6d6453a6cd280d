"""Probe: Black-Scholes end to end from pinned host buffers.
(a) fadb_black_scholes_host at several chunk sizes (staged copies, 3 streams);
(b) the resident kernel launched directly on pinned host pointers (zero-copy over PCIe, UVA)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import fastad_b200 as F
import synth

n = 1 << 20
F.load(); L = F.lib(); F.check(L.fadb_init(0)); F.use_torch_stream()

keys = ["S", "K", "sigma", "tau", "r"]
rng = np.random.default_rng(1)
arr = dict(S=rng.uniform(50, 150, n), K=rng.uniform(50, 150, n), sigma=rng.uniform(.05, .6, n), tau=rng.uniform(.05, 2, n), r=rng.uniform(0, .08, n))
h_in = [torch.from_numpy(arr[k]).pin_memory() for k in keys]
h_out = [torch.zeros(n, dtype=torch.float64).pin_memory() for _ in range(8)]
d_in = [t.cuda() for t in h_in]
d_out = [torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(8)]

def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

import ctypes
def raw(ins, outs):
    tab = (ctypes.c_void_p * 8)(*[t.data_ptr() for t in outs])
    def f():
        F.check(L.fadb_black_scholes(ctypes.c_size_t(n), *[ctypes.c_void_p(t.data_ptr()) for t in ins], tab, ctypes.c_uint(F.OVERWRITE_ADJ)))
    return f

np_in = [t.numpy() for t in h_in]; np_out = [t.numpy() for t in h_out]
if os.environ.get("PROBE_ONLY_HOST"):
    raw(d_in, d_out)(); torch.cuda.synchronize()
    ms = timeit(lambda: F.black_scholes_host(*np_in, np_out))
    same = all(torch.equal(a, b.cpu()) for a, b in zip(h_out, d_out))
    pg_in = [a.copy() for a in np_in]; pg_out = [np.zeros(n) for _ in range(8)]
    ms_pg = timeit(lambda: F.black_scholes_host(*pg_in, pg_out), reps=5)
    same_pg = all(np.array_equal(a, b.cpu().numpy()) for a, b in zip(pg_out, d_out))
    print("mode", os.environ.get("FADB_BS_HOST_MODE"), "chunk", os.environ.get("FADB_BS_HOST_CHUNK_LOG2"), "pinned ms %.4f" % ms, "bits", same, "| pageable ms %.4f" % ms_pg, "bits", same_pg)
    sys.exit(0)
print("resident            ms", timeit(raw(d_in, d_out)))
print("zero-copy in+out    ms", timeit(raw(h_in, h_out)))
print("zero-copy in only   ms", timeit(raw(h_in, d_out)))
print("zero-copy out only  ms", timeit(raw(d_in, h_out)))
ref = [t.clone() for t in h_out]
raw(d_in, d_out)(); torch.cuda.synchronize()
print("zero-copy == resident bits:", all(torch.equal(a, b.cpu()) for a, b in zip(ref, d_out)))
np_in = [t.numpy() for t in h_in]; np_out = [t.numpy() for t in h_out]
print("staged host call    ms", timeit(lambda: F.black_scholes_host(*np_in, np_out)))
# plain copies for reference
big_h = torch.empty(8 * n, dtype=torch.float64).pin_memory(); big_d = torch.empty(8 * n, dtype=torch.float64, device="cuda")
print("D2H 64 MB copy      ms", timeit(lambda: big_h.copy_(big_d, non_blocking=True)))
print("H2D 40 MB copy      ms", timeit(lambda: big_d[:5 * n].copy_(big_h[:5 * n], non_blocking=True)))
