"""A/B of the Black-Scholes kernel variants (FADB_BS_VARIANT) on one box: device time per 2^20-option launch through
fadb_black_scholes_batches (K launches per host call, rotating buffer sets > L2), the host-buffer (e2e) call, and a digest of
the results (all variants share bs_eval_fast, so the digests must agree).
  python tools/bs_variants.py 23 30 31 ...        (each variant runs in its own process: the choice is read once)"""
import hashlib, os, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(variant):
    sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
    import numpy as np, torch
    import fastad_b200 as F, synth
    F.load(); F.check(F.lib().fadb_init(0)); F.use_torch_stream()
    n = int(os.environ.get("BS_N", 1 << 20))
    keys = ("S", "K", "sigma", "tau", "r")
    sets = []
    for s in range(4):
        inp = synth.black_scholes_inputs(n, seed=synth.SEED + s)
        sets.append(([torch.from_numpy(inp[k]).cuda() for k in keys],
                     [torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(8)]))
    B = F.BlackScholesBatches(sets)
    B.run(40)
    torch.cuda.synchronize()
    h = hashlib.sha256()
    for t in sets[0][1]:
        h.update(t.cpu().numpy().tobytes())
    # accumulate mode on set 1 (outputs hold one overwrite pass): prices stored, adjoints doubled
    B2 = F.BlackScholesBatches(sets[1:2], flags=0)
    before = [t.clone() for t in sets[1][1]]
    B2.run(1)
    torch.cuda.synchronize()
    acc_ok = all(torch.equal(sets[1][1][k], before[k] * (1 if k < 2 else 2)) for k in range(8))
    K, reps = 200, 9
    times = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps):
        e0.record(); B.run(K); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / K * 1e3)
    times.sort()
    # 20-launch windows (the driver's --steps 20)
    t20 = []
    for _ in range(25):
        e0.record(); B.run(20); e1.record(); torch.cuda.synchronize()
        t20.append(e0.elapsed_time(e1) / 20 * 1e3)
    t20.sort()
    # the same windows enqueued behind the host-released gate (tools/probes/gate.cu): device time of the 20 launches
    import ctypes as C
    G = C.CDLL(os.path.join(R, "tools", "_build", "libfadb_probes.so"))
    G.fadb_probe_gate_arm.argtypes = [C.c_void_p]
    g20 = []
    for _ in range(25):
        torch.cuda.synchronize()
        G.fadb_probe_gate_arm(None)
        e0.record(); B.run(20); e1.record()
        G.fadb_probe_gate_release(); torch.cuda.synchronize()
        g20.append(e0.elapsed_time(e1) / 20 * 1e3)
    g20.sort()
    # host-buffer call
    inp = synth.black_scholes_inputs(n)
    h_in = [torch.from_numpy(inp[k]).pin_memory().numpy() for k in keys]
    h_out = [torch.zeros(n, dtype=torch.float64).pin_memory().numpy() for _ in range(8)]
    for _ in range(3): F.black_scholes_host(*h_in, h_out)
    te = []
    for _ in range(10):
        e0.record(); F.black_scholes_host(*h_in, h_out); e1.record(); torch.cuda.synchronize()
        te.append(e0.elapsed_time(e1))
    te.sort()
    print("variant %3d  us/launch min %.2f med %.2f | 20-step windows min %.2f med %.2f max %.2f | gated min %.2f med %.2f max %.2f | e2e ms min %.3f med %.3f | acc %s | %s"
          % (variant, times[0], times[len(times) // 2], t20[0], t20[len(t20) // 2], t20[-1], g20[0], g20[len(g20) // 2], g20[-1], te[0], te[len(te) // 2],
             "ok" if acc_ok else "MISMATCH", h.hexdigest()[:16]), flush=True)


if __name__ == "__main__":
    if os.environ.get("BS_CHILD"):
        child(int(os.environ["FADB_BS_VARIANT"]))
    else:
        for v in sys.argv[1:] or ["23", "30"]:
            env = dict(os.environ, BS_CHILD="1", FADB_BS_VARIANT=v)
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, timeout=300)
