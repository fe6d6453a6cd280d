#!/usr/bin/env python
"""bench.py — headline benchmark of the FastAD reverse-mode hot path on B200.

Headline (BASELINE.json `metric`: "gradient evals/sec (batched Black-Scholes; ...)"):
one step = one pass of the hot path (price + reverse-mode greeks of call AND put,
example/black_scholes.cpp) over one batch of 2^20 synthetic options; one gradient eval = one
option.  `value` = whole-job options/s with inputs resident in HBM; `e2e` = the same metric
through the host-buffer C-ABI call (H2D + D2H inside the timed region).  The other BASELINE
configs (sum map-reduce, normal log-pdf, regression, 3-layer net) are timed the same way and
reported under `workloads`, each with its own HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--quick]
N > 1: launched by torchrun, one rank per GPU; the batch shards across ranks with no
data-path collective (weak scaling); the normal log-pdf workload shards by element range with
one 4-double NCCL all-reduce per evaluation.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

BS_N = 1 << 20            # BASELINE config 0/1: batch of 2^20 options
# one string for both arms (the driver compares `config.workload` of the two lines)
BS_WORKLOAD = ("black_scholes call+put price + delta/vega/rho (example/black_scholes.cpp), "
               "batch 2^20 options per GPU per step")
BS_BYTES = 104            # algorithmic bytes per option: 5 in + 8 out doubles
L2_BYTES = 126 * (1 << 20)
NCCL_LOG_DIR = os.environ.get("FADB_NCCL_LOG_DIR", "/tmp/fadb_nccl_logs")
# committed `ncu --set full` summary of the headline kernel (profiles/): this round's capture, else round 1's
NCU_BS = next((f for f in ("r2_ncu_full_bs.csv", "r1_ncu_full_bs.csv")
               if os.path.exists(os.path.join(ROOT, "profiles", f))), "r2_ncu_full_bs.csv")
BS_KERNEL = ("bs_kernel_fast_pf<false, true, 3, 256, PDL_EARLY> (batch 0 of each K-step call: the PDL_HEAD instantiation, "
             "same code with the dependency wait at its start)")


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the headline kernel, from the
    committed `ncu --set full` summary (profiles/); None when no capture is on file."""
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        return None
    tot, scale = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    import csv
    for row in csv.reader(open(p)):
        if row and row[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(row[2]) * scale.get(row[1], 1.0)
    return tot or None


def ncu_metric(name, metric):
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        return None
    import csv
    for row in csv.reader(open(p)):
        if row and row[0] == metric:
            return float(row[2])
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        # NCCL's INFO log is NOT silenced: it goes to a file per process (stdout stays the one JSON
        # line); rank 0 copies the communicator lines (rank / nranks) to stderr at the end of the run.
        os.environ["NCCL_DEBUG"] = "INFO"
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        os.environ["NCCL_DEBUG_FILE"] = os.path.join(NCCL_LOG_DIR, "nccl_%h_%p.log")
        os.makedirs(NCCL_LOG_DIR, exist_ok=True)
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, world):
    import torch
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed_loop(fn, steps, warmup, world, windows=3):
    """W untimed calls, then `windows` timed regions of exactly K calls of fn(i) each (CUDA events on the launching
    stream, barrier + synchronize on both sides, max over ranks); returns the MEDIAN window in ms."""
    import torch
    for i in range(warmup):
        fn(i)
    out = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for w in range(windows):
        barrier(world)
        e0.record()
        for i in range(steps):
            fn(warmup + w * steps + i)
        e1.record()
        barrier(world)
        out.append(max_over_ranks(e0.elapsed_time(e1), world))
    return sorted(out)[len(out) // 2]


def timed_windows(run_k, steps, world, windows, gate=None):
    """`windows` repeats of the K-step timed region (run_k(w) issues exactly K steps), each bracketed by barrier +
    synchronize and reduced by MAX over ranks; returns the sorted per-window times in ms.  The median is what is reported:
    a single 0.45 ms window (20 x 22 us) swings by ~8 % with clock ramp and launch jitter.
    gate (tools/probes/gate.cu): the region is enqueued behind a host-released gate, so the K launches run from a full
    queue and the window is the DEVICE time of exactly K steps, not K steps + the CPU's enqueue latency of the first."""
    import torch
    out = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for w in range(windows):
        barrier(world)
        gated = gate is not None and gate.fadb_probe_gate_arm(torch.cuda.current_stream().cuda_stream) == 0
        e0.record()
        run_k(w)
        e1.record()
        if gated:
            gate.fadb_probe_gate_release()
        barrier(world)
        out.append(max_over_ranks(e0.elapsed_time(e1), world))
    return sorted(out)


def load_probes():
    """tools/_build/libfadb_probes.so: the DFMA and host-link micro-benchmarks that anchor the FP64 and e2e rooflines."""
    import ctypes as C
    p = os.path.join(ROOT, "tools", "_build", "libfadb_probes.so")
    if not os.path.exists(p):
        return None
    L = C.CDLL(p)
    L.fadb_probe_dfma_peak.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.fadb_probe_pcie.argtypes = [C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.fadb_probe_gate_arm.argtypes = [C.c_void_p]
    return L


# --------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path on this box's host cores.
    FastAD needs Eigen 3.3, absent from this image, so the real reference cannot be built
    (oracle/Makefile `ref-probe`); the Eigen-free node-by-node port in oracle/ stands in
    (kind "port"), built -O3 -march=native, batch split over all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_py as O
    import synth
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    cores = os.cpu_count() or 1
    n = 1 << 18 if args.quick else BS_N
    inp = synth.black_scholes_inputs(n)
    keys = ("S", "K", "sigma", "tau", "r")
    call = lambda: O.black_scholes(*(inp[k] for k in keys), nthreads=cores, fast=True)
    for _ in range(max(1, min(args.warmup, 10))):
        call()
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    dt = time.perf_counter() - t0
    v = n * steps / dt
    line = {
        "impl": "reference", "metric": "gradient evals/sec (batched Black-Scholes)", "value": v,
        "unit": "options/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": BS_WORKLOAD if n == BS_N else BS_WORKLOAD + f" (quick: {n} options)",
                   "host": "CPU port of the reference path (FastAD+Eigen is not buildable here: no Eigen)"},
        "cpu_baseline": {"value": v, "unit": "options/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {n} options, {cores} threads"},
        "e2e": {"value": v, "unit": "options/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import fastad_b200 as F
    import synth

    rank, world, local = dist_setup(args.gpus)
    F.load()
    F.check(F.lib().fadb_init(local))
    F.use_torch_stream()
    L = F.lib()
    peak_gbs, peak_src = load_peaks()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    zeros = lambda n: torch.zeros(n, dtype=torch.float64, device="cuda")
    keys = ("S", "K", "sigma", "tau", "r")

    # ---------------- headline: Black-Scholes, inputs resident, rotating buffer sets > L2
    n = BS_N
    nsets = 4                                     # 4 x 104 MiB = 416 MiB > 126 MiB L2
    sets = []
    for s in range(nsets):
        inp = synth.black_scholes_inputs(n, seed=synth.SEED + 100 * rank + s)
        sets.append(([dev(inp[k]) for k in keys], [zeros(n) for _ in range(8)], inp))

    # parity gate before timing (oracle = test infrastructure, used only as the checker): the full 2^20 batch of set 0
    # through the SAME batched entry point that is timed below
    batches = F.BlackScholesBatches([(d_in, d_out) for d_in, d_out, _ in sets])
    if rank == 0:
        import oracle_py as O
        m = n if not args.quick else 1 << 14
        batches.run(1, 0)
        ref = O.black_scholes(*(sets[0][2][k][:m] for k in keys), nthreads=os.cpu_count() or 1)
        import bs_tol                                  # tests/bs_tol.py: eps x the magnitude of each output's terms
        sub = {k: sets[0][2][k][:m] for k in keys}
        for nm, g, o, sc in zip(bs_tol.NAMES, sets[0][1], ref, bs_tol.scales(sub)):
            err = np.abs(g[:m].cpu().numpy() - o)
            assert np.all(err <= 8.0 * sc), f"parity gate failed on {nm}: {float(np.max(err / sc)):.2f} x scale"

    # ---------------- N > 1: the sharded evaluations must be CORRECT before anything is timed (tests/sharded_gate.py)
    # The one exchange of a sharded evaluation (1-66 doubles): the peer-memory mailbox all-reduce of
    # csrc/p2p.cu (one tiny kernel per rank over NVLink, measured 8-10 us against NCCL's 18-35 us at N=2);
    # NCCL if the mailboxes cannot be set up (no CUDA IPC) or --collective nccl.
    collective = "none (1 GPU)"
    small_all_reduce = lambda t: None
    peer = None
    if world > 1:
        small_all_reduce, collective = (lambda t: torch.distributed.all_reduce(t)), "NCCL all-reduce"
        if args.collective == "p2p":
            try:
                from fastad_b200 import sharding as _sh
                peer = _sh.PeerAllReduce(world, rank)
                small_all_reduce, collective = peer, "peer-memory mailbox all-reduce over NVLink (csrc/p2p.cu)"
            except Exception as ex:        # every rank fails or succeeds together (same node, same driver)
                collective = "NCCL all-reduce (peer-memory setup failed: %s)" % str(ex)[:80]
    gate = None
    if world > 1 and not args.quick:
        import sharded_gate
        fused_gate = collective.startswith("peer-memory")
        gate = sharded_gate.run(F, world, rank, small_all_reduce, fused_gate)     # raises -> non-zero exit
        gate = {"passed": True, "fused_exchange": fused_gate,
                "checked": "closed-form value <= 1e-12, vector + scalar adjoints bit-exact, reduced doubles bit-identical "
                           "on every rank, fadb_p2p_status == 0", "sizes": gate}

    # K steps = K launches from ONE host call (fadb_black_scholes_batches): no Python between launches
    clk = ClockSampler(local)
    clk.__enter__()
    launches0 = L.fadb_launch_count()
    batches.run(args.warmup, 0)
    windows = 3 if args.quick else args.windows
    probes = load_probes()
    gate_lib = None if args.no_gate else probes
    ws = timed_windows(lambda w: batches.run(args.steps, args.warmup + w * args.steps), args.steps, world, windows, gate_lib)
    ms = ws[len(ws) // 2]                          # median window
    launches = (L.fadb_launch_count() - launches0 - args.warmup) // windows
    value = world * n * args.steps / (ms * 1e-3)
    kernel_ms = ms / args.steps                    # one kernel launch per step, back to back
    achieved = BS_BYTES * n / (kernel_ms * 1e-3) / 1e9

    # ---------------- measured FP64 issue peak (register-resident DFMA chains, all SMs) for the second roofline
    fp64_peak = None
    if probes is not None:
        import ctypes as C
        o3 = (C.c_double * 3)()
        if probes.fadb_probe_dfma_peak(8, 8, 2048, o3) == 0:
            fp64_peak = o3[0]

    # ---------------- e2e: host buffers through the C ABI, H2D + D2H inside the timed region
    h_in = [torch.from_numpy(sets[0][2][k]).pin_memory() for k in keys]
    h_out = [torch.zeros(n, dtype=torch.float64).pin_memory() for _ in range(8)]
    np_in = [t.numpy() for t in h_in]
    np_out = [t.numpy() for t in h_out]
    e2e_steps = max(3, min(args.steps, 30))
    for _ in range(max(8, args.warmup)):           # includes the six calls in which the library picks its transfer scheme
        F.black_scholes_host(*np_in, np_out)
    import ctypes as C
    tm = (C.c_double * 2)()
    host_choice = L.fadb_black_scholes_host_mode(tm)
    host_scheme = {"choice": {1: "direct (one launch over PCIe)", 0: "staged (copy engines)"}.get(host_choice, "measuring"),
                   "staged_ms": tm[0], "direct_ms": tm[1]}

    def e2e_k(w):
        for _ in range(e2e_steps):
            F.black_scholes_host(*np_in, np_out)
    ws_e2e = timed_windows(e2e_k, e2e_steps, world, 3 if args.quick else 5)
    ms_e2e = ws_e2e[len(ws_e2e) // 2]
    e2e_val = world * n * e2e_steps / (ms_e2e * 1e-3)
    batches.run(1, 0)
    torch.cuda.synchronize()
    for k in range(8):   # the host-buffer call returns the same bits as the resident call
        np.testing.assert_array_equal(np_out[k], sets[0][1][k].cpu().numpy())
    # the host-link ceiling the e2e call competes against: plain cudaMemcpyAsync of the same bytes, both directions at
    # once, on ALL ranks at the same time (aggregate host ceiling at N GPUs)
    link = None
    if probes is not None:
        import ctypes as C
        ms1 = C.c_double()
        h2d_b, d2h_b = 40 * n, 64 * n
        res = {}
        for mode, nm in ((0, "both"), (1, "h2d_only"), (2, "d2h_only")):
            probes.fadb_probe_pcie(h2d_b, d2h_b, 2, mode, C.byref(ms1))        # warm
            barrier(world)
            probes.fadb_probe_pcie(h2d_b, d2h_b, 10, mode, C.byref(ms1))
            res[nm] = max_over_ranks(ms1.value, world)
            barrier(world)
        link = {"copy_engine_ms_per_step": res["both"], "h2d_only_ms": res["h2d_only"], "d2h_only_ms": res["d2h_only"],
                "aggregate_gbs_both": world * (h2d_b + d2h_b) / (res["both"] * 1e-3) / 1e9,
                "aggregate_gbs_h2d": world * h2d_b / (res["h2d_only"] * 1e-3) / 1e9,
                "aggregate_gbs_d2h": world * d2h_b / (res["d2h_only"] * 1e-3) / 1e9,
                "note": "plain cudaMemcpyAsync of one step's bytes from / to page-locked buffers on every rank at once (max over ranks)"}

    # ---------------- the other BASELINE configs (rank-local unless stated)
    workloads = {}

    def add(name, bytes_per_launch, ms_total, steps, units, unit_name, extra=None, sharded=False):
        # sharded=True: ONE evaluation spans all ranks (element/row-range shards + all-reduce);
        # otherwise every rank evaluates its own independent instance.
        k_ms = ms_total / steps
        gbs = bytes_per_launch / (k_ms * 1e-3) / 1e9
        workloads[name] = {"ms_per_step": k_ms, "evals_per_s": (1 if sharded else world) * steps / (ms_total * 1e-3),
                           unit_name + "_per_s": world * units * steps / (ms_total * 1e-3),
                           "hbm_gbs": gbs, "roofline_frac": gbs / peak_gbs,
                           "algorithmic_bytes": bytes_per_launch}
        if extra:
            workloads[name].update(extra)

    wl = set() if args.quick else set(args.workloads.split(","))
    steps2 = max(5, min(args.steps, 50))
    import ctypes as C
    P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    if "sum" in wl:
        # C2: sum(exp(x)), sum(log(x)), sum(pow<2>(x)); N = 2^24; 24 B/elem (x read + adj RMW)
        n2 = 1 << 24
        xs = [dev(synth.sum_inputs(n2, seed=synth.SEED + s)) for s in range(2)]   # 2 x 128 MiB
        adjs = [zeros(n2) for _ in range(2)]
        for op in ("exp", "log", ("pow", 2)):
            code = F._op_code(op)
            dval = zeros(1)

            fused_sum = world > 1 and collective.startswith("peer-memory")

            def step(i, code=code):
                # N > 1: every rank owns 2^24 elements of ONE sum; the seed is known up front, so the
                # adjoints are shard-local and the exchange is the 1-double value (sharding.sharded_sum_autodiff);
                # with the mailboxes connected the all-reduce is fused into the pass (the last CTA pushes, waits, sums)
                call = L.fadb_sum_unary_allreduce_async if fused_sum else L.fadb_sum_unary_async
                F.check(call(code, n2, C.c_void_p(xs[i % 2].data_ptr()), C.c_void_p(adjs[i % 2].data_ptr()), 1.0, 0,
                             C.c_void_p(dval.data_ptr())))
                if world > 1 and not fused_sum:
                    small_all_reduce(dval)
            t = timed_loop(step, steps2, 3, world)
            nm = op if isinstance(op, str) else f"pow{op[1]}"
            add(f"sum_{nm}_2^24", 24 * n2, t, steps2, n2, "elements",
                {"exchange": "all-reduce fused into the pass (fadb_sum_unary_allreduce_async)"} if fused_sum else None,
                sharded=world > 1)
        # prod_benchmark.cpp shape: ad::prod over a vector (prod.hpp:172-212): reduce pass (8 B/elem) + adjoint pass (24 B/elem)
        def prod_step(i):
            F.check(L.fadb_prod(n2, C.c_void_p(xp.data_ptr()), C.c_void_p(adjs[i % 2].data_ptr()), 1.0, 0, None))
        xp = dev(np.exp(np.random.default_rng(3).uniform(-1e-3, 1e-3, n2)))       # product stays O(1)
        t = timed_loop(prod_step, steps2, 3, world)
        add("prod_2^24", 32 * n2, t, steps2, n2, "elements", {"note": "two passes: (prod of non-zeros, #zeros), then adj_i += seed * prod / x_i"})
        del xs, adjs, xp

    if "normal" in wl:
        # C3: normal log-pdf, N = 2^26 per rank.  vvv 72 B/elem; vss (x var) 24; vss (x const) 8
        n3 = 1 << 26
        x, mu, sg = (dev(a) for a in synth.normal_inputs(n3, seed=synth.SEED + rank))
        xa, ma, sa = zeros(n3), zeros(n3), zeros(n3)
        part = zeros(4); inval = zeros(1)

        from fastad_b200 import sharding
        all_reduce = small_all_reduce

        def normal_step(shape, x_adj, mu_t, sg_t, mu_adj, sg_adj, flags=0, n3=n3, x=x):
            # rank-local shard of n3 elements; N > 1: one 4-double all-reduce per evaluation
            class Ops:
                def sigma_check(self):
                    F.check(L.fadb_sigma_check(n3, P(sg_t), P(inval)))
                    return inval

                def partial(self, cnt, seed):
                    F.check(L.fadb_normal_lpdf_partial(shape, n3, P(x), P(mu_t), P(sg_t), P(x_adj), P(mu_adj),
                                                       P(sg_adj), seed, flags, P(cnt), P(part)))
                    return part

                def finish(self, prt, n_total, seed):
                    F.check(L.fadb_normal_lpdf_finish(shape, n_total, P(prt), P(x), P(mu_t), P(sg_t), None,
                                                      P(mu_adj), P(sg_adj), seed, flags, None))

                def undo(self, prt, seed):      # returns at once unless the global count prt[3] is non-zero
                    F.check(L.fadb_normal_lpdf_undo(shape, n3, P(x), P(mu_t), P(sg_t), P(x_adj), P(mu_adj), P(sg_adj), seed, flags,
                                                    C.c_void_p(prt.data_ptr() + 24)))
            ops = Ops()

            fused = world > 1 and collective.startswith("peer-memory")
            guarded = bool(shape & 2) and not (flags & F.TRUST_SIGMA)
            strict = bool(flags & F.STRICT_SIGMA)

            def step(i):
                if fused:
                    # exchange fused into the two kernels (fadb_normal_lpdf_partial_push / _finish_pull).  The sigma
                    # validity count travels with it (optimistic pass + undo); only the strict form needs its own
                    # 1-double mailbox exchange ahead of the pass
                    cnt = None
                    if guarded and strict:
                        cnt = ops.sigma_check()
                        all_reduce(cnt)
                    F.check(L.fadb_normal_lpdf_partial_push(shape, n3, P(x), P(mu_t), P(sg_t), P(x_adj), P(mu_adj),
                                                            P(sg_adj), 1.0, flags, P(cnt), P(part)))
                    F.check(L.fadb_normal_lpdf_finish_pull(shape, n3 * world, P(x), P(mu_t), P(sg_t), None,
                                                           P(mu_adj), P(sg_adj), 1.0, flags, None))
                    if guarded and not strict:
                        F.check(L.fadb_normal_lpdf_undo(shape, n3, P(x), P(mu_t), P(sg_t), P(x_adj), P(mu_adj), P(sg_adj), 1.0,
                                                        flags, None))
                else:
                    sharding.sharded_normal_autodiff(ops, all_reduce, shape, n3 * world, 1.0, guarded, strict)
            return step

        mu1, sg1, mua1, sga1 = dev([0.1]), dev([1.3]), zeros(1), zeros(1)
        t = timed_loop(normal_step(3, xa, mu, sg, ma, sa), steps2, 3, world)
        add("normal_vvv_2^26", 72 * n3, t, steps2, n3, "elements",
            {"note": "exact semantics (default): optimistic single pass, the sigma validity count rides on the partial sums, "
                     "undo launch behind the finish (returns at once when every sigma_i > 0); per-GPU shard"},
            sharded=True)
        t = timed_loop(normal_step(3, xa, mu, sg, ma, sa, F.STRICT_SIGMA), steps2, 3, world)
        add("normal_vvv_2^26_strict_sigma", 72 * n3, t, steps2, n3, "elements",
            {"note": "FADB_STRICT_SIGMA: +8 B/elem sigma validity pre-pass (80 B/elem traffic); per-GPU shard"},
            sharded=True)
        t = timed_loop(normal_step(3, xa, mu, sg, ma, sa, F.TRUST_SIGMA), steps2, 3, world)
        add("normal_vvv_2^26_trust_sigma", 72 * n3, t, steps2, n3, "elements", sharded=True)
        t = timed_loop(normal_step(0, xa, mu1, sg1, mua1, sga1), steps2, 3, world)
        add("normal_vss_2^26", 24 * n3, t, steps2, n3, "elements", sharded=True)
        t = timed_loop(normal_step(0, None, mu1, sg1, mua1, sga1), steps2, 3, world)
        add("normal_vss_xconst_2^26", 8 * n3, t, steps2, n3, "elements", sharded=True)
        del x, mu, sg, xa, ma, sa
        # the upper size of BASELINE config 3: 2^28 elements per GPU, vvv (6 x 2 GiB of operands and adjoints); inputs
        # drawn on the device (same distributions; parity at this size is tests/test_gpu_kernels.py's job)
        n5 = 1 << 28
        gen = torch.Generator(device="cuda").manual_seed(synth.SEED + rank)
        xL = torch.randn(n5, dtype=torch.float64, device="cuda", generator=gen)
        muL = torch.randn(n5, dtype=torch.float64, device="cuda", generator=gen)
        sgL = torch.rand(n5, dtype=torch.float64, device="cuda", generator=gen) * 1.5 + 0.5
        xaL, maL, saL = zeros(n5), zeros(n5), zeros(n5)
        steps5 = max(5, min(steps2, 20))
        t = timed_loop(normal_step(3, xaL, muL, sgL, maL, saL, 0, n5, xL), steps5, 3, world)
        add("normal_vvv_2^28", 72 * n5, t, steps5, n5, "elements",
            {"note": "exact semantics (default): optimistic single pass + undo launch; per-GPU shard"},
            sharded=True)
        t = timed_loop(normal_step(3, xaL, muL, sgL, maL, saL, F.TRUST_SIGMA, n5, xL), steps5, 3, world)
        add("normal_vvv_2^28_trust_sigma", 72 * n5, t, steps5, n5, "elements", sharded=True)
        del xL, muL, sgL, xaL, maL, saL
        torch.cuda.empty_cache()

    if "logpdf" in wl:
        # the other diagonal log-pdfs, all-vector operands, N = 2^25, exact semantics
        # (optimistic pass + undo launch) and FADB_TRUST_SIGMA (the pass alone)
        n4 = 1 << 25
        rng = np.random.default_rng(synth.SEED + 77)
        x = dev(rng.uniform(-1, 1, n4))
        loc, sc = dev(rng.uniform(-1, 1, n4)), dev(rng.uniform(0.2, 2, n4))
        lo, hi = dev(-1.5 - rng.uniform(0, 1, n4)), dev(1.5 + rng.uniform(0, 1, n4))
        xb, pb = dev(rng.integers(0, 2, n4).astype(np.float64)), dev(rng.uniform(0.05, 0.95, n4))
        a1, a2, a3, out4 = zeros(n4), zeros(n4), zeros(n4), zeros(4)

        def lp_step(dist, X, B, Cc, XA, BA, CA, flags):
            def step(i):
                F.check(L.fadb_logpdf_async(dist, 1 | (2 if Cc is not None else 0), n4, P(X), P(B), P(Cc), P(XA), P(BA), P(CA),
                                            1.0, flags, P(out4)))
            return step
        for nm, dist, ops_, bytes_per, pre in (("cauchy_vvv", 0, (x, loc, sc, a1, a2, a3), 72, 24),
                                               ("uniform_vvv", 1, (x, lo, hi, None, a2, a3), 56, 24),
                                               ("bernoulli_vv", 2, (xb, pb, None, None, a2, None), 32, 16)):
            t = timed_loop(lp_step(dist, *ops_, 0), steps2, 3, world)
            add(f"{nm}_2^25", bytes_per * n4, t, steps2, n4, "elements",
                {"note": "exact semantics (default): optimistic single pass + undo launch (returns at once when every "
                         f"element is in range); FADB_STRICT_SIGMA would add a {pre} B/elem validity pre-pass"})
            t = timed_loop(lp_step(dist, *ops_, F.TRUST_SIGMA), steps2, 3, world)
            add(f"{nm}_2^25_trusted", bytes_per * n4, t, steps2, n4, "elements")
        del x, loc, sc, lo, hi, xb, pb, a1, a2, a3

    if "dense" in wl:
        # SURVEY.md 8f row 3 / BASELINE.md section 1: BM_normal_adj_log_pdf (vvm, dense Sigma), the only shape the
        # reference publishes timings for (benchmark/ref/normal_benchmark_res.py:4-15, ns per autodiff, unknown CPU)
        published_ns = {1000: 2720733, 2000: 16687945, 4000: 104822861}
        for nd in (1000, 2000, 4000):
            rng = np.random.default_rng(nd)
            Lo = np.tril(rng.uniform(-1, 1, (nd, nd)), -1) + 10 * np.eye(nd)     # make_cov, normal_benchmark.cpp:15-33
            Sg = torch.from_numpy(Lo @ Lo.T).cuda()
            xd, md = dev(rng.uniform(-1, 1, nd)), dev(rng.uniform(-1, 1, nd))
            xa, ma, Sa, dv = zeros(nd), zeros(nd), zeros(nd * nd), zeros(1)

            def dn_step(i):
                F.check(L.fadb_normal_dense_async(nd, P(xd), P(md), 1, P(Sg), P(xa), P(ma), P(Sa), 1.0, 0, 1, P(dv)))
            sd = max(3, min(steps2, 10))
            t = timed_loop(dn_step, sd, 2, world)
            flops = nd ** 3 * (1 / 3 + 1 / 3 + 1 / 3)          # potrf + trtri + lauum (lower halves)
            workloads[f"normal_dense_vvm_n{nd}"] = {
                "ms_per_step": t / sd, "evals_per_s": world * sd / (t * 1e-3),
                "fp64_gflops": flops / (t / sd * 1e-3) / 1e9,
                "reference_published_ns": published_ns[nd],
                "speedup_vs_published": published_ns[nd] * 1e-6 / (t / sd),
                "note": "published number: unknown CPU, unstated units (Google-Benchmark default ns), reference README has no hardware"}
            del Sg, Sa

    if "linreg" in wl:
        # C4: regression, X 2^20 x 64 column-major; 520 B/row
        rows, cols = 1 << 20, 64
        Xs = []
        for s in range(2):                          # 2 x 512 MiB > L2
            Xn, yn, wt = synth.linreg_inputs(rows, cols, seed=synth.SEED + s)
            Xs.append((torch.from_numpy(np.ascontiguousarray(Xn.T)).cuda(), dev(yn)))
        w, wa, sg1, sga1 = zeros(cols), zeros(cols), dev([1.0]), zeros(1)
        lpart = zeros(cols + 1)

        from fastad_b200 import sharding
        all_reduce = small_all_reduce

        def lin_step(i):
            Xd, yd = Xs[i % 2]

            class Ops:      # rank-local row shard; N > 1: one (cols+1)-double all-reduce
                def partial(self):
                    F.check(L.fadb_linreg_partial(rows, cols, P(Xd), P(yd), P(w), P(lpart)))
                    return lpart

                def finish(self, prt, rows_total, seed):
                    F.check(L.fadb_linreg_finish(rows_total, cols, P(prt), P(sg1), P(wa), P(sga1), seed, 0, None))
            if fused_exchange:
                # exchange fused into the two kernels: the last CTA of the pass pushes {sum r^2, X^T r} into every
                # rank's mailbox over NVLink, the finish kernel waits, reduces in rank order and finishes
                F.check(L.fadb_linreg_partial_push(rows, cols, P(Xd), P(yd), P(w), P(lpart)))
                F.check(L.fadb_linreg_finish_pull(rows * world, cols, P(sg1), P(wa), P(sga1), 1.0, 0, None))
            else:
                sharding.sharded_linreg_autodiff(Ops(), all_reduce, rows * world, 1.0)
        fused_exchange = world > 1 and collective.startswith("peer-memory")
        t = timed_loop(lin_step, steps2, 3, world)
        add("linreg_2^20x64", 520 * rows, t, steps2, rows, "rows",
            {"exchange": "fused into the pass and the finish kernel (fadb_linreg_partial_push / _finish_pull)"} if fused_exchange else None,
            sharded=True)
        del Xs

    if "mlp3" in wl:
        # C5: 3-layer net, batch 2^16 (configured) and 2^22 (roofline-sized); 24 B/sample
        for b in (1 << 16, 1 << 22):
            Xn, yn = synth.mlp3_inputs(b)
            Xd, yd = torch.from_numpy(np.ascontiguousarray(Xn.T)).cuda(), dev(yn)
            import kats
            parm = dev(kats.NN_PARM)
            gl = zeros(26)                      # {grad[25], loss}: one buffer so that N > 1 exchanges it in one piece
            p_grad, p_loss = C.c_void_p(gl.data_ptr()), C.c_void_p(gl.data_ptr() + 25 * 8)
            mflags = F.OVERWRITE_ADJ if world > 1 else 0

            fused_mlp = world > 1 and collective.startswith("peer-memory")

            def mlp_step(i):
                # N > 1: rows are sharded over the ranks, the 25 gradients + the loss are sums over rows: ONE
                # 26-double all-reduce per evaluation (sharding.sharded_mlp3_autodiff), fused into the kernel's
                # last CTA when the mailboxes are connected (fadb_mlp3_allreduce_async: one launch per evaluation)
                call = L.fadb_mlp3_allreduce_async if fused_mlp else L.fadb_mlp3_async
                F.check(call(b, P(Xd), P(yd), P(parm), p_grad, mflags, p_loss))
                if world > 1 and not fused_mlp:
                    small_all_reduce(gl)
            t = timed_loop(mlp_step, steps2, 3, world)
            extra = {"bound": "fp64 pipe / launch latency, not HBM"}
            if fused_mlp:
                extra["exchange"] = "all-reduce fused into the kernel (fadb_mlp3_allreduce_async)"
            add(f"mlp3_batch_{b}", 24 * b, t, steps2, b, "samples", extra, sharded=world > 1)

    if "expr" in wl:
        # the GENERIC path (ad::bind -> planner -> interpreter kernel), same C-ABI calls the C++
        # host layer makes: (i) the 3-deep chain of SURVEY.md 8d C2, sum(-0.5*pow<2>((x - c*w0)/w1)),
        # (ii) the batched Black-Scholes call expression with vec leaves and 3 vec placeholders
        import exprs
        n2 = 1 << 24
        e = F.Expr()
        x = e.var(synth.sum_inputs(n2))
        w0, w1 = e.var(2.0), e.var(1.3)
        cc = e.const(synth.sum_inputs(n2, seed=5))
        z = e.binary("div", e.binary("sub", x, e.binary("mul", w0, cc)), w1)
        e.bind(e.sum(e.binary("mul", e.const(-0.5), e.pow(2, z))))

        def chain_step(i):
            F.check(L.fadb_expr_autodiff(e.h, 1.0, None, None))
        # default engine: the stage kernel specialised at bind time by NVRTC (csrc/jit.cu), then the
        # same stage on the tape interpreter (fadb_jit_enable(0))
        j0 = F.jit_stats()
        t = timed_loop(chain_step, steps2, 3, world)
        j1 = F.jit_stats()
        add("expr_sum_chain_2^24_specialised", 32 * n2, t, steps2, n2, "elements",
            {"note": "g++-side C-ABI path; x read + c read + x_adj RMW = 32 B/elem; 2 scalar adjoints by reduction channels",
             "specialised_kernel_launches": j1["launched"] - j0["launched"]})
        prev_jit = F.jit_enable(False)
        t = timed_loop(chain_step, steps2, 3, world)
        F.jit_enable(prev_jit)
        add("expr_sum_chain_2^24_interpreter", 32 * n2, t, steps2, n2, "elements",
            {"note": "same expression, specialisation off"})
        del e
        nb = BS_N
        exprs.RNG = np.random.default_rng(1)
        e = F.Expr()
        root, leaves, _ = exprs.black_scholes_vec(nb)(e)
        e.bind(root)
        ones = torch.ones(nb, dtype=torch.float64, device="cuda")

        def bs_expr_step(i):
            F.check(L.fadb_expr_autodiff(e.h, 0.0, P(ones), None))
        j0 = F.jit_stats()
        t = timed_loop(bs_expr_step, steps2, 3, world)
        j1 = F.jit_stats()
        # S,K,sigma,tau,sqrt_tau,r reads (48) + seed read (8) + price store (8) + 3 adj RMW (48)
        # + 3 placeholders: value store + adjoint RMW (72) = 184 B/option
        add("expr_black_scholes_call_2^20_specialised", 184 * nb, t, steps2, nb, "options",
            {"note": "one autodiff of the CALL expression only, reference placeholder semantics kept",
             "specialised_kernel_launches": j1["launched"] - j0["launched"]})
        prev_jit = F.jit_enable(False)
        t = timed_loop(bs_expr_step, steps2, 3, world)
        F.jit_enable(prev_jit)
        add("expr_black_scholes_call_2^20_interpreter", 184 * nb, t, steps2, nb, "options",
            {"note": "same expression, specialisation off"})
        del e

        # least squares sum(pow<2>(y - dot(X, w))), the batched form of the reference's README regression
        # example (README.md:640-662): recognised at bind time, one pass over X; then the same expression
        # with an extra map between dot and pow (not recognised): tall GEMV kernels, two passes over X
        rows, cols = 1 << 20, 64
        Xh, yh, wt = synth.linreg_inputs(rows, cols)
        L.fadb_glm_enable.argtypes, L.fadb_glm_enable.restype = [C.c_int], C.c_int
        # BASELINE config 4 with an intercept: normal(y, dot(X, w) + b, sigma) -- no longer the fixed regression shape (fast=3),
        # but still one pass over X: the mean's map and the log-density run inside the tile kernel (fast=6)
        for tag, glm in (("one_pass", 1), ("three_launch", 0)):
            e = F.Expr()
            Xc, yc = e.const(Xh), e.const(yh)
            w, b, sg = e.var((wt + 0.05).reshape(cols, 1)), e.var(0.1), e.var(1.3)
            e.bind(e.normal(yc, e.binary("add", e.dot(Xc, w), b), sg))
            prev_glm = L.fadb_glm_enable(glm)

            def li_step(i):
                F.check(L.fadb_expr_autodiff(e.h, 1.0, None, None))
            t = timed_loop(li_step, steps2, 3, world)
            L.fadb_glm_enable(prev_glm)
            add(f"expr_linreg_intercept_2^20x64_{tag}", 8 * rows * (cols + 1), t, steps2, rows, "rows",
                {"note": "normal(y, dot(X, w) + b, sigma); algorithmic bytes = one pass over X + y"
                         + ("" if glm else "; one-pass route off: gemv -> normal stage -> gemv^T")})
            del e
        # a posterior: the regression with intercept as log-likelihood plus three prior terms.  With the additive-root execution every
        # term is swept once as a root of its own (the likelihood on the one-pass tile kernel); without it the likelihood runs
        # gemv -> normal stage forward, then normal stage backward -> gemv^T after the root
        for tag, on in (("per_term", 1), ("one_block", 0)):
            e = F.Expr()
            Xc, yc = e.const(Xh), e.const(yh)
            w, b, sg = e.var((wt + 0.05).reshape(cols, 1)), e.var(0.1), e.var(1.3)
            lik = e.normal(yc, e.binary("add", e.dot(Xc, w), b), sg)
            pri = e.binary("add", e.normal(w, e.const(0.0), e.const(2.0)), e.normal(b, e.const(0.0), e.const(10.0)))
            e.bind(e.binary("add", lik, e.binary("sub", pri, e.binary("mul", e.const(0.5), e.binary("mul", sg, sg)))))
            prev_add = L.fadb_additive_root_enable(on)

            def post_step(i):
                F.check(L.fadb_expr_autodiff(e.h, 1.0, None, None))
            t = timed_loop(post_step, steps2, 3, world)
            L.fadb_additive_root_enable(prev_add)
            add(f"expr_posterior_linreg_2^20x64_{tag}", 8 * rows * (cols + 1), t, steps2, rows, "rows",
                {"note": "normal(y, dot(X, w) + b, s) + normal(w, 0, 2) + normal(b, 0, 10) - 0.5 s^2; algorithmic bytes = one pass over X + y"})
            del e
        for tag, inner, glm in (("one_pass", False, 1), ("map_one_pass", True, 1), ("generic_dot", True, 0)):
            e = F.Expr()
            Xc, yc, w = e.const(Xh), e.const(yh), e.var((wt + 0.05).reshape(cols, 1))
            r = e.binary("sub", yc, e.dot(Xc, w))
            if inner:          # a map between dot and the reduction: not the fixed least-squares shape
                r = e.binary("mul", r, e.unary("exp", e.binary("mul", e.const(-0.01), e.binary("mul", r, r))))
            e.bind(e.sum(e.pow(2, r)))
            prev_glm = L.fadb_glm_enable(glm)

            def ls_step(i):
                F.check(L.fadb_expr_autodiff(e.h, 1.0, None, None))
            t = timed_loop(ls_step, steps2, 3, world)
            L.fadb_glm_enable(prev_glm)
            add(f"expr_least_squares_2^20x64_{tag}", 8 * rows * (cols + 1), t, steps2, rows, "rows",
                {"note": "algorithmic bytes = one pass over X + y"
                         + ("; the map (r exp(-0.01 r^2)) runs inside the one-pass tile kernel specialised at bind time (fast=6)" if inner and glm else "")
                         + ("; same expression with the one-pass route off: gemv -> map stage -> gemv^T, two passes over X" if inner and not glm else "")})
            del e

        # the same two expressions through the TYPE-FUSED path (include/fastad_b200/fused.cuh): the
        # expression type instantiates the kernel in an nvcc-compiled TU (tests/cpp/fused_bench.cu)
        fb_path = os.path.join(ROOT, "tests", "_build", "libfused_bench.so")
        if os.path.exists(fb_path):
            FB = C.CDLL(fb_path)
            FB.fused_chain_setup.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
            FB.fused_bs_setup.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
            xs, xa, cs = dev(synth.sum_inputs(n2)), zeros(n2), dev(synth.sum_inputs(n2, seed=5))
            assert FB.fused_chain_setup(n2, P(xs), P(xa), P(cs)) == 0
            t = timed_loop(lambda i: FB.fused_chain_step(), steps2, 3, world)
            add("expr_sum_chain_2^24_type_fused", 32 * n2, t, steps2, n2, "elements",
                {"note": "same expression as the interpreter line; kernel instantiated from the expression type"})
            del xs, xa, cs
            inp = synth.black_scholes_inputs(nb, seed=11)
            bufs = [dev(inp["S"]), zeros(nb), dev(inp["sigma"]), zeros(nb), dev(inp["r"]), zeros(nb), dev(inp["K"]),
                    dev(inp["tau"]), dev(np.sqrt(inp["tau"])), zeros(nb), zeros(nb), zeros(nb), zeros(nb), zeros(nb),
                    zeros(nb), ones]
            arr = (C.c_void_p * 16)(*[b.data_ptr() for b in bufs])
            assert FB.fused_bs_setup(nb, arr) == 0
            t = timed_loop(lambda i: FB.fused_bs_step(), steps2, 3, world)
            add("expr_black_scholes_call_2^20_type_fused", 184 * nb, t, steps2, nb, "options",
                {"note": "CALL expression only, reference placeholder semantics kept (3 vec placeholders)"})
            del bufs

    clk.__exit__()
    clocks = clk.summary()     # sampled over every timed region above (headline first)

    # ---------------- CPU baseline beside it (rank 0, bounded sample)
    cpu = None
    if rank == 0 and not args.no_cpu:
        import oracle_py as O
        cores = os.cpu_count() or 1
        m = 1 << 20 if not args.quick else 1 << 16
        inp = sets[0][2]
        t0 = time.perf_counter()
        reps = 0
        while True:
            O.black_scholes(*(inp[k][:m] for k in keys), nthreads=cores, fast=True)
            reps += 1
            if time.perf_counter() - t0 > 8.0 or reps >= 20:
                break
        dt = time.perf_counter() - t0
        cpu = {"value": m * reps / dt, "unit": "options/s", "cores": cores, "kind": "port",
               "sample": f"{reps} x {m} options, {cores} threads, oracle -O3 -march=native"}

    # exchange health after every timed loop (a timed-out wait would have poisoned the results with NaN)
    p2p_status = None
    if peer is not None:
        assert peer.healthy(), "peer-memory exchange reported a timeout during the timed loops"
        p2p_status = 0

    if rank == 0:
        # the sharded configs (C2-C5) in a field the driver keeps: per-GPU GB/s, ms, and the exchange that was timed
        sharded = {"n_gpus": world, "exchange": collective, "correctness_gate": gate, "p2p_status_after_timing": p2p_status}
        for k, v in workloads.items():
            if k.startswith(("normal_", "sum_", "linreg_", "mlp3_", "prod_")) and "hbm_gbs" in v:
                sharded[k] = {"ms": round(v["ms_per_step"], 5), "gbs_per_gpu": round(v["hbm_gbs"], 1),
                              "frac_hbm": round(v["roofline_frac"], 4), "evals_per_s": round(v["evals_per_s"], 2)}
        fp64 = {"pipe_active_pct_ncu": ncu_metric(NCU_BS, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                "inst_executed_per_launch_ncu": ncu_metric(NCU_BS, "sm__inst_executed_pipe_fp64.sum")}
        if fp64_peak:
            # executed FP64 warp-instructions (ncu, committed capture) x 32 lanes / launch time / measured DFMA peak
            fp64["measured_peak_thread_inst_per_s"] = fp64_peak
            fp64["measured_peak_tflops"] = 2 * fp64_peak / 1e12
            if fp64["inst_executed_per_launch_ncu"]:
                rate = fp64["inst_executed_per_launch_ncu"] * 32 / (kernel_ms * 1e-3)
                fp64["executed_thread_inst_per_s"] = rate
                fp64["frac_of_measured_peak"] = rate / fp64_peak
        e2e = {"value": e2e_val, "unit": "options/s", "h2d_bytes_per_step": 40 * n,
               "d2h_bytes_per_step": 64 * n, "ms_per_step": ms_e2e / e2e_steps,
               "windows_ms_per_step": [round(w / e2e_steps, 4) for w in ws_e2e],
               "api": "fadb_black_scholes_host on page-locked host buffers; transfer scheme picked by the library on its first calls",
               "scheme": host_scheme}
        if link:
            e2e["host_link"] = link
            e2e["roofline"] = {"bound": "host link (PCIe, both directions)", "achieved_gbs": world * 104 * n / (ms_e2e / e2e_steps * 1e-3) / 1e9,
                               "peak_gbs": link["aggregate_gbs_both"],
                               "frac": link["copy_engine_ms_per_step"] / (ms_e2e / e2e_steps)}
        line = {
            "metric": "gradient evals/sec (batched Black-Scholes)", "value": value, "unit": "options/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": kernel_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": BS_WORKLOAD,
                       "l2": f"rotating {nsets} resident buffer sets ({nsets * BS_BYTES * n >> 20} MiB) > 126 MiB L2",
                       "parallelism": f"batch sharded over {world} GPU(s), no data-path collective",
                       "timing": f"median of {windows} windows of exactly {args.steps} steps (one launch each, issued by ONE "
                                 f"fadb_black_scholes_batches call per window"
                                 + (", enqueued behind a host-released gate: device time of the K steps" if gate_lib else "")
                                 + f"); windows_us_per_step min/med/max = "
                                 f"{ws[0] / args.steps * 1e3:.2f}/{kernel_ms * 1e3:.2f}/{ws[-1] / args.steps * 1e3:.2f}",
                       "collective_of_the_sharded_workloads": collective,
                       "sharded": sharded},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                         "frac": achieved / peak_gbs, "traffic": ncu_traffic(NCU_BS),
                         "traffic_note": "ncu dram bytes inside the launch; most of the 64 B/option of stores is still in "
                                         "the 126 MB L2 when the kernel ends",
                         "peak_source": peak_src,
                         "kernel": BS_KERNEL, "algorithmic_bytes_per_launch": BS_BYTES * n,
                         # the second bound of this kernel: FP64-pipe issue, against the MEASURED DFMA peak
                         "fp64": fp64},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "workloads": workloads,
        }
        print(json.dumps(line))
        # the driver keeps the tail of stderr: the sharded configs in <= 1 KB, printed LAST
        tail = {"n_gpus": world, "bs_us": round(kernel_ms * 1e3, 2), "bs_frac": round(achieved / peak_gbs, 3),
                "e2e_ms": round(ms_e2e / e2e_steps, 3), "gate": bool(gate and gate["passed"]) if world > 1 else None,
                "exch": "p2p" if collective.startswith("peer-memory") else ("nccl" if world > 1 else None)}
        for k, v in workloads.items():
            if k in ("normal_vvv_2^26", "normal_vvv_2^26_trust_sigma", "normal_vvv_2^28", "normal_vss_2^26", "normal_vss_xconst_2^26",
                     "sum_exp_2^24", "linreg_2^20x64", "mlp3_batch_65536", "mlp3_batch_4194304"):
                tail[k] = [round(v["ms_per_step"], 4), round(v["hbm_gbs"])]
        sys.stderr.flush()
        stderr_tail = "SHARDED_SUMMARY [ms, GB/s per GPU] " + json.dumps(tail, separators=(",", ":"))
    else:
        stderr_tail = None
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        if peer is not None:
            peer.close()
        dist.destroy_process_group()
    if stderr_tail:
        if world > 1:      # the communicator lines of NCCL's INFO log (every rank's "... rank r nranks N ..." init line)
            import glob
            seen = 0
            for f in sorted(glob.glob(os.path.join(NCCL_LOG_DIR, "nccl_*.log")), key=os.path.getmtime)[-world:]:
                for ln in open(f, errors="replace"):
                    if "nranks" in ln and "Init COMPLETE" in ln and seen < 16:
                        print(ln.rstrip()[:300], file=sys.stderr)
                        seen += 1
            if not seen:
                for f in sorted(glob.glob(os.path.join(NCCL_LOG_DIR, "nccl_*.log")), key=os.path.getmtime)[-world:]:
                    for ln in open(f, errors="replace"):
                        if "nranks" in ln and seen < 16:
                            print(ln.rstrip()[:300], file=sys.stderr)
                            seen += 1
        print(stderr_tail[:1000], file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--windows", type=int, default=15, help="repeats of the K-step timed region; the median is reported")
    ap.add_argument("--quick", action="store_true", help="headline only, small CPU sample")
    ap.add_argument("--workloads", default="sum,normal,logpdf,dense,linreg,mlp3,expr",
                    help="secondary workloads to time besides the headline (comma list)")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: exchange of the sharded reductions (peer-memory mailboxes or NCCL)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-gate", action="store_true", help="time the K-step windows without the host-released gate")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
