"""A/B of the host-buffer Black-Scholes call with 1 / 2 / 4 options per thread on the PCIe path (FADB_BS_HOST_WIDE), each in its
own process (the setting is read once); direct scheme pinned (FADB_BS_HOST_MODE=1).  Prints ms per 2^20-option call and a
digest of the results (must agree)."""
import hashlib, os, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
    import numpy as np, torch
    import fastad_b200 as F, synth
    F.load(); F.check(F.lib().fadb_init(0)); F.use_torch_stream()
    n = 1 << 20
    keys = ("S", "K", "sigma", "tau", "r")
    inp = synth.black_scholes_inputs(n)
    h_in = [torch.from_numpy(inp[k]).pin_memory().numpy() for k in keys]
    h_out = [torch.zeros(n, dtype=torch.float64).pin_memory().numpy() for _ in range(8)]
    for _ in range(4): F.black_scholes_host(*h_in, h_out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    te = []
    for _ in range(15):
        e0.record(); F.black_scholes_host(*h_in, h_out); e1.record(); torch.cuda.synchronize()
        te.append(e0.elapsed_time(e1))
    te.sort()
    h = hashlib.sha256()
    for a in h_out: h.update(a.tobytes())
    print("FADB_BS_HOST_WIDE=%s grid=%s  e2e ms min %.4f med %.4f max %.4f | %s" % (os.environ.get("FADB_BS_HOST_WIDE", "0"),
          os.environ.get("FADB_BS_HOST_GRID", "32"), te[0], te[len(te) // 2], te[-1], h.hexdigest()[:16]), flush=True)


if __name__ == "__main__":
    if os.environ.get("E2E_CHILD"):
        child()
    else:
        for w, g in (("0", "32"), ("2", "32"), ("4", "32"), ("4", "16"), ("4", "64"), ("2", "64"), ("0", "32")):
            env = dict(os.environ, E2E_CHILD="1", FADB_BS_HOST_WIDE=w, FADB_BS_HOST_GRID=g, FADB_BS_HOST_MODE="1")
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, timeout=300)
