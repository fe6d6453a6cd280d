"""Per-launch time of the network kernel against batch size (CUDA graph replay: no CPU launch cost)."""
import os, sys, ctypes as C
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import fastad_b200 as F, synth, kats
F.load(); L = F.lib(); F.check(L.fadb_init(0))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    F.use_torch_stream()
    for b in (32, 256, 4096, 37888, 65536, 1 << 18, 1 << 20):
        X, y = synth.mlp3_inputs(b)
        Xd, yd, pd = torch.from_numpy(np.asfortranarray(X).reshape(-1, order="F")).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(kats.NN_PARM).cuda()
        g, dl = torch.zeros(25, dtype=torch.float64, device="cuda"), torch.zeros(1, dtype=torch.float64, device="cuda")
        margs = [C.c_size_t(b)] + [C.c_void_p(t.data_ptr()) for t in (Xd, yd, pd, g)] + [C.c_uint(0), C.c_void_p(dl.data_ptr())]
        for _ in range(3): L.fadb_mlp3_async(*margs)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            F.use_torch_stream()
            for _ in range(100): L.fadb_mlp3_async(*margs)
        torch.cuda.synchronize()
        for _ in range(3): gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10): gr.replay()
        e1.record(s); torch.cuda.synchronize()
        print("batch %8d  %.2f us/launch" % (b, e0.elapsed_time(e1)))
