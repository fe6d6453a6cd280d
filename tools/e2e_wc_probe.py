"""Does write-combined page-locked memory for the five INPUT arrays (cudaHostAllocWriteCombined: no CPU cache snooping on the
device's reads) move the host-buffer Black-Scholes call?  Inputs in ordinary page-locked memory vs write-combined, outputs
ordinary page-locked either way; direct and staged scheme.  Prints ms per 2^20-option call and a digest of the results."""
import ctypes as C, hashlib, os, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(wc, mode):
    sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
    import numpy as np, torch
    import fastad_b200 as F, synth
    F.load(); F.check(F.lib().fadb_init(0)); F.use_torch_stream()
    rt = C.CDLL("libcudart.so.12")
    n = 1 << 20
    keys = ("S", "K", "sigma", "tau", "r")
    inp = synth.black_scholes_inputs(n)

    def host_array(src, flags):
        p = C.c_void_p()
        assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(8 * n), C.c_uint(flags)) == 0
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n,))
        if src is not None: a[:] = src
        return a
    h_in = [host_array(inp[k], 4 if wc else 0) for k in keys]           # cudaHostAllocWriteCombined = 4
    h_out = [host_array(None, 0) for _ in range(8)]
    for _ in range(4): F.black_scholes_host(*h_in, h_out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    te = []
    for _ in range(15):
        e0.record(); F.black_scholes_host(*h_in, h_out); e1.record(); torch.cuda.synchronize()
        te.append(e0.elapsed_time(e1))
    te.sort()
    h = hashlib.sha256()
    for a in h_out: h.update(a.tobytes())
    print("inputs %-14s scheme %-6s  e2e ms min %.4f med %.4f max %.4f | %s" % ("write-combined" if wc else "page-locked", "direct" if mode == "1" else "staged",
          te[0], te[len(te) // 2], te[-1], h.hexdigest()[:16]), flush=True)


if __name__ == "__main__":
    if os.environ.get("E2E_CHILD"):
        child(os.environ["E2E_WC"] == "1", os.environ["FADB_BS_HOST_MODE"])
    else:
        for wc in ("0", "1", "0", "1"):
            for mode in ("1", "0"):
                env = dict(os.environ, E2E_CHILD="1", E2E_WC=wc, FADB_BS_HOST_MODE=mode)
                subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, timeout=300)
