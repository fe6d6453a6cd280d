"""Correctness gate of the SHARDED evaluations on real GPUs, one process per GPU.

The reference's contract is one value + gradients per `ad::autodiff` (eval.hpp:82-104).  At N > 1 ranks that value (and
the adjoints of scalar operands) crosses the peer-memory exchange of csrc/p2p.cu, so this gate runs every sharded
workload once on inputs whose per-element quantities are EXACTLY representable and asserts, on every rank:
  * value == closed form (<= 1e-12 relative; bit-exact where the sum itself is exact),
  * vector adjoints of the shard bit-exact, scalar adjoints bit-exact,
  * the reduced doubles bit-identical on every rank (all-gather and compare),
  * fadb_p2p_status == 0 afterwards (no wait timed out).
bench.py calls run() before timing whenever world > 1 (a failure makes the bench exit non-zero); tests/test_multi_gpu.py
launches it under torchrun when the box has >= 2 GPUs.  Nothing here touches oracle/: the expectations are closed forms.
"""
import ctypes as C
import math

import numpy as np

REL = 1e-12


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def run(F, world, rank, small_all_reduce, fused, rows_per_rank_log2=18, n_per_rank_log2=22, skip_exchange_on_rank=None):
    """small_all_reduce(t): in-place sum of a <=128-double CUDA tensor over the ranks (mailbox or NCCL);
    fused: the exchange is fused into the kernels (mailboxes connected).  Returns a dict of what was checked.
    skip_exchange_on_rank: fault injection -- that rank leaves out ONE exchange (the others must time out, not
    return a plausible number)."""
    import torch
    import torch.distributed as dist
    L = F.lib()
    P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    f64 = dict(dtype=torch.float64, device="cuda")
    done = {}

    def same_on_all_ranks(name, t):
        # bit-identical reduced doubles on every rank
        if world == 1:
            return
        bits = t.detach().clone().view(torch.int64)
        allb = [torch.empty_like(bits) for _ in range(world)]
        dist.all_gather(allb, bits)
        for r in range(world):
            assert torch.equal(allb[r], allb[0]), f"{name}: rank {r} holds different bits than rank 0"

    def exact_sum(t):
        # all-reduce of exactly representable partial sums (any order is exact): NCCL is fine as the checker
        if world > 1:
            dist.all_reduce(t)
        return t

    # ------------------------------------------------------------------ normal log-pdf, vvv and vss (x variable)
    n = (1 << n_per_rank_log2) + 8 * rank                 # ragged shards; multiples of the 8-cycle
    n_total = sum((1 << n_per_rank_log2) + 8 * r for r in range(world))
    i = torch.arange(n, device="cuda")
    zc = torch.tensor([0.5, -1.0, 1.5, -2.0, 2.0, -1.5, 1.0, 0.5], **f64)        # sum z = 1, sum z^2 = 15 per cycle
    z = zc[i & 7]
    seed = 0.75
    part, inval = torch.zeros(4, **f64), torch.zeros(1, **f64)

    def normal_eval(shape, x, mu, sg, xa, ma, sa, flags):
        # default: optimistic pass, the count of sigma_i <= 0 rides on the one exchange, undo launch behind the finish;
        # FADB_STRICT_SIGMA: the count is exchanged first; FADB_TRUST_SIGMA: no guard
        val = C.c_double()
        guarded = bool(shape & 2) and not (flags & F.TRUST_SIGMA)
        strict = bool(flags & F.STRICT_SIGMA)
        cnt = None
        if guarded and strict:
            F.check(L.fadb_sigma_check(n, P(sg), P(inval)))
            small_all_reduce(inval)
            cnt = inval
        if fused:
            F.check(L.fadb_normal_lpdf_partial_push(shape, n, P(x), P(mu), P(sg), P(xa), P(ma), P(sa), seed, flags, P(cnt), P(part)))
            F.check(L.fadb_normal_lpdf_finish_pull(shape, n_total, P(x), P(mu), P(sg), None, P(ma), P(sa), seed, flags, C.byref(val)))
            if guarded and not strict:
                F.check(L.fadb_normal_lpdf_undo(shape, n, P(x), P(mu), P(sg), P(xa), P(ma), P(sa), seed, flags, None))
        else:
            F.check(L.fadb_normal_lpdf_partial(shape, n, P(x), P(mu), P(sg), P(xa), P(ma), P(sa), seed, flags, P(cnt), P(part)))
            small_all_reduce(part)
            F.check(L.fadb_normal_lpdf_finish(shape, n_total, P(part), P(x), P(mu), P(sg), None, P(ma), P(sa), seed, flags, C.byref(val)))
            if guarded and not strict:
                F.check(L.fadb_normal_lpdf_undo(shape, n, P(x), P(mu), P(sg), P(xa), P(ma), P(sa), seed, flags,
                                                C.c_void_p(part.data_ptr() + 24)))
        return val.value

    # vvv: sigma cycles through powers of two whose logs cancel per cycle
    sg = torch.tensor([1.0, 2.0, 0.5, 4.0, 0.25, 2.0, 0.5, 1.0], **f64)[i & 7]
    mu = (i & 1023).to(torch.float64) / 1024.0
    x = mu + sg * z
    for flags in (0, F.TRUST_SIGMA, F.STRICT_SIGMA):
        xa, ma, sa = (torch.full((n,), 0.5, **f64) for _ in range(3))
        v = normal_eval(3, x, mu, sg, xa, ma, sa, flags)
        assert _rel(v, -0.5 * 15.0 * n_total / 8.0) <= REL, ("normal vvv value", v, flags)
        assert torch.equal(xa, 0.5 - seed * z / sg) and torch.equal(ma, 0.5 + seed * z / sg) \
            and torch.equal(sa, 0.5 + seed * (z * z - 1.0) / sg), "normal vvv adjoints"
        same_on_all_ranks("normal vvv value", torch.tensor([v], **f64))
    # ONE sigma_i <= 0, in the last rank's shard: every rank must report -inf and hold untouched adjoints
    # (normal.hpp:440-457) -- through the undo of the optimistic pass (every contribution here is exactly representable,
    # so taking it back restores the previous bits) and through the strict pre-pass
    sg_bad = sg.clone()
    if rank == world - 1:
        sg_bad[n - 3] = -1.0
    for flags in (0, F.STRICT_SIGMA):
        xa, ma, sa = (torch.full((n,), 0.5, **f64) for _ in range(3))
        v = normal_eval(3, x, mu, sg_bad, xa, ma, sa, flags)
        torch.cuda.synchronize()
        assert v == -math.inf, ("normal vvv with one invalid sigma", v, flags)
        assert bool((xa == 0.5).all()) and bool((ma == 0.5).all()) and bool((sa == 0.5).all()), \
            ("normal vvv adjoints after an invalid evaluation", flags)
    del sg_bad
    done["normal_vvv"] = n_total
    # vss: scalar mu = 0.25, sigma = 2 as variables; their adjoints are sums over ALL shards
    mu1, sg1 = torch.tensor([0.25], **f64), torch.tensor([2.0], **f64)
    x = 0.25 + 2.0 * z
    xa, ma1, sa1 = torch.full((n,), 0.5, **f64), torch.full((1,), 0.5, **f64), torch.full((1,), 0.5, **f64)
    v = normal_eval(0, x, mu1, sg1, xa, ma1, sa1, 0)
    cyc = n_total / 8.0
    assert _rel(v, -0.5 * 15.0 * cyc - n_total * math.log(2.0)) <= REL, ("normal vss value", v)
    assert torch.equal(xa, 0.5 - seed * z / 2.0), "normal vss x adjoint"
    assert float(ma1[0]) == 0.5 + seed * cyc / 2.0, ("normal vss mu adjoint", float(ma1[0]), 0.5 + seed * cyc / 2.0)
    assert float(sa1[0]) == 0.5 + seed * (15.0 - 8.0) * cyc / 2.0, ("normal vss sigma adjoint", float(sa1[0]))
    same_on_all_ranks("normal vss {value, mu_adj, sigma_adj}", torch.cat([torch.tensor([v], **f64), ma1, sa1]))
    done["normal_vss"] = n_total
    del x, mu, sg, xa, ma, sa, i, z

    # ------------------------------------------------------------------ regression, row shards
    rows, cols = 1 << rows_per_rank_log2, 64
    rows_total = rows * world
    i = torch.arange(rows, device="cuda") + rank * rows
    j = torch.arange(cols, device="cuda")
    X = (((i[None, :] * 7 + j[:, None] * 3) % 5) - 2).to(torch.float64)          # (cols, rows) == rows x cols column-major
    w = ((j % 3) - 1).to(torch.float64)
    r = torch.tensor([0.5, -1.0, 1.5, -2.0, 2.0, -1.5, 1.0, -0.5], **f64)[i & 7]
    y = w @ X + r
    sigma = torch.tensor([2.0], **f64)
    w_adj, s_adj = torch.full((cols,), 0.5, **f64), torch.full((1,), 0.5, **f64)
    lpart = torch.zeros(cols + 1, **f64)
    val = C.c_double()
    if fused:
        F.check(L.fadb_linreg_partial_push(rows, cols, P(X), P(y), P(w), P(lpart)))
        F.check(L.fadb_linreg_finish_pull(rows_total, cols, P(sigma), P(w_adj), P(s_adj), seed, 0, C.byref(val)))
    else:
        F.check(L.fadb_linreg_partial(rows, cols, P(X), P(y), P(w), P(lpart)))
        small_all_reduce(lpart)
        F.check(L.fadb_linreg_finish(rows_total, cols, P(lpart), P(sigma), P(w_adj), P(s_adj), seed, 0, C.byref(val)))
    sum_r2 = 15.0 * rows_total / 8.0
    xtr = exact_sum(X @ r)                                                        # integers x dyadics: exact in any order
    assert _rel(val.value, -0.5 * sum_r2 / 4.0 - rows_total * math.log(2.0)) <= REL, ("linreg value", val.value)
    assert torch.equal(w_adj, 0.5 + seed * 0.25 * xtr), "linreg w adjoint"
    assert float(s_adj[0]) == 0.5 + seed * (sum_r2 / 4.0 - rows_total) / 2.0, "linreg sigma adjoint"
    same_on_all_ranks("linreg {value, w_adj, sigma_adj}", torch.cat([torch.tensor([val.value], **f64), w_adj, s_adj]))
    done["linreg"] = rows_total
    del X, y, r, i

    # ------------------------------------------------------------------ sum(f(x)), element shards
    n2 = (1 << n_per_rank_log2) + 4 * rank
    xi = ((torch.arange(n2, device="cuda") * 5 + rank) % 13 - 6).to(torch.float64)   # integers in [-6, 6]
    adj, dv = torch.full((n2,), 0.5, **f64), torch.zeros(1, **f64)
    code = F._op_code(("pow", 2))
    if fused:
        if skip_exchange_on_rank != rank:
            F.check(L.fadb_sum_unary_allreduce_async(code, n2, P(xi), P(adj), seed, 0, P(dv)))
    else:
        F.check(L.fadb_sum_unary_async(code, n2, P(xi), P(adj), seed, 0, P(dv)))
        if skip_exchange_on_rank != rank:
            small_all_reduce(dv)
    torch.cuda.synchronize()
    want = exact_sum((xi * xi).sum().reshape(1))
    assert float(dv[0]) == float(want[0]), ("sum(pow<2>(x)) value", float(dv[0]), float(want[0]))
    assert torch.equal(adj, 0.5 + seed * 2.0 * xi), "sum(pow<2>(x)) adjoint"
    same_on_all_ranks("sum value", dv)
    # a transcendental map: value against the all-reduced torch sum to 1e-12
    xe = torch.rand(n2, generator=torch.Generator(device="cuda").manual_seed(7 + rank), **f64) + 0.5
    adj.zero_()
    code = F._op_code("exp")
    if fused:
        F.check(L.fadb_sum_unary_allreduce_async(code, n2, P(xe), P(adj), 1.0, 0, P(dv)))
    else:
        F.check(L.fadb_sum_unary_async(code, n2, P(xe), P(adj), 1.0, 0, P(dv)))
        small_all_reduce(dv)
    torch.cuda.synchronize()
    want = torch.exp(xe).sum().reshape(1)
    if world > 1:
        dist.all_reduce(want)
    assert _rel(float(dv[0]), float(want[0])) <= REL, "sum(exp(x)) value"
    same_on_all_ranks("sum(exp) value", dv)
    done["sum"] = n2
    del xi, xe, adj

    # ------------------------------------------------------------------ network, batch shards: against the UNSHARDED evaluation
    import kats
    import synth
    b = 1 << 14
    Xn, yn = synth.mlp3_inputs(b * world)                # every rank draws the whole batch (same seed), owns rows [rank*b, +b)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    parm = dev(kats.NN_PARM)
    Xf, yf = dev(Xn.T), dev(yn)
    full = torch.zeros(26, **f64)
    F.check(L.fadb_mlp3_async(b * world, P(Xf), P(yf), P(parm), P(full), F.OVERWRITE_ADJ, C.c_void_p(full.data_ptr() + 200)))
    Xs, ys = dev(Xn[rank * b:(rank + 1) * b].T), dev(yn[rank * b:(rank + 1) * b])
    gl = torch.zeros(26, **f64)
    call = L.fadb_mlp3_allreduce_async if fused else L.fadb_mlp3_async
    F.check(call(b, P(Xs), P(ys), P(parm), P(gl), F.OVERWRITE_ADJ, C.c_void_p(gl.data_ptr() + 200)))
    if not fused:
        small_all_reduce(gl)
    torch.cuda.synchronize()
    scale = float(full.abs().max())
    assert float((gl - full).abs().max()) <= 1e-11 * scale, ("network {grad[25], loss} vs unsharded", float((gl - full).abs().max()), scale)
    same_on_all_ranks("network {grad[25], loss}", gl)
    done["mlp3"] = b * world

    # ------------------------------------------------------------------ the exchange stayed healthy
    if world > 1 and (fused or hasattr(small_all_reduce, "healthy")):
        e = C.c_ulonglong(0)
        F.check(L.fadb_p2p_status(C.byref(e)))
        assert e.value == 0, f"peer-memory exchange timed out at sequence {e.value}"
    done["p2p_status"] = 0
    return done
