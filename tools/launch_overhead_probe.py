"""Is a bench step CPU-issue-bound?  Wall-clock cost of ISSUING n calls (no sync) against the CUDA-event time."""
import os, sys, time, ctypes as C
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import fastad_b200 as F, synth
F.load(); L = F.lib(); F.check(L.fadb_init(0)); F.use_torch_stream()
n = 1 << 20
inp = synth.black_scholes_inputs(n)
d_in = [torch.from_numpy(inp[k]).cuda() for k in ("S", "K", "sigma", "tau", "r")]
d_out = [torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(8)]
def wrapped(): F.black_scholes(*d_in, out=d_out, flags=F.OVERWRITE_ADJ)
arr = (C.c_void_p * 8)(*[o.data_ptr() for o in d_out])
args = [C.c_size_t(n)] + [C.c_void_p(t.data_ptr()) for t in d_in] + [arr, C.c_uint(F.OVERWRITE_ADJ)]
fn = L.fadb_black_scholes
def raw(): fn(*args)
for name, f in (("python wrapper", wrapped), ("prebuilt ctypes args", raw)):
    for _ in range(20): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 400
    t0 = time.perf_counter(); e0.record()
    for _ in range(k): f()
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("%-22s issue %.2f us/call   gpu %.2f us/call" % (name, (t1 - t0) / k * 1e6, e0.elapsed_time(e1) / k * 1e3))
# mlp3 2^16
b = 1 << 16
import kats
X, y = synth.mlp3_inputs(b); parm = kats.NN_PARM
Xd, yd, pd = torch.from_numpy(np.asfortranarray(X).reshape(-1, order="F")).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(parm).cuda()
g, dl = torch.zeros(25, dtype=torch.float64, device="cuda"), torch.zeros(1, dtype=torch.float64, device="cuda")
margs = [C.c_size_t(b)] + [C.c_void_p(t.data_ptr()) for t in (Xd, yd, pd, g)] + [C.c_uint(0), C.c_void_p(dl.data_ptr())]
mf = L.fadb_mlp3_async
def mraw(): mf(*margs)
for _ in range(20): mraw()
torch.cuda.synchronize()
k = 1000
t0 = time.perf_counter(); e0.record()
for _ in range(k): mraw()
e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
print("mlp3 2^16 prebuilt     issue %.2f us/call   gpu %.2f us/call" % ((t1 - t0) / k * 1e6, e0.elapsed_time(e1) / k * 1e3))
# the same launches replayed from a CUDA graph (no per-launch CPU work at all)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    F.use_torch_stream()
    for _ in range(3): mraw()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        F.use_torch_stream()
        for _ in range(100): mraw()
    torch.cuda.synchronize()
    for _ in range(3): gr.replay()
    e0.record(s)
    for _ in range(10): gr.replay()
    e1.record(s); torch.cuda.synchronize()
    print("mlp3 2^16 CUDA graph   gpu %.2f us/call" % (e0.elapsed_time(e1) / 1000 * 1e3))
    gr2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr2, stream=s):
        F.use_torch_stream()
        for _ in range(100): raw()
    torch.cuda.synchronize()
    for _ in range(3): gr2.replay()
    e0.record(s)
    for _ in range(10): gr2.replay()
    e1.record(s); torch.cuda.synchronize()
    print("black-scholes CUDA graph gpu %.2f us/call" % (e0.elapsed_time(e1) / 1000 * 1e3))
