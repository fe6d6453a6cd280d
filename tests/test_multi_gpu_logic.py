"""World-size-2 gloo test (CPU) of the element-range sharding logic used at N > 1 GPUs:
per-shard partial sums -> one all-reduce -> value and scalar adjoints, against the oracle on the
unsharded problem.  The per-shard arithmetic here is a numpy stand-in for the CUDA kernels (same
partial-sum layout as fadb_normal_lpdf_partial / fadb_linreg_partial); the GPU kernels themselves
are covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from fastad_b200 import sharding
    import synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        n = 10007
        x, mu, sg = synth.normal_inputs(n)
        mu_s, sg_s, seed = 0.1, 1.3, 0.75
        b, e = sharding.shard_range(n, rank, world)
        assert b % 4 == 0
        ar = lambda t: dist.all_reduce(t)
        out = {}

        # ---- vvv: vector adjoints are shard-local, the value needs the exchange
        class VVV:
            def __init__(self, sigma):
                self.sg = sigma
                self.undone = None

            def sigma_check(self):
                return torch.tensor([float((self.sg[b:e] <= 0).sum())], dtype=torch.float64)

            def contrib(self, seed):
                ok = self.sg[b:e] > 0
                with np.errstate(all="ignore"):
                    g = seed * (mu[b:e] - x[b:e]) / (self.sg[b:e] * self.sg[b:e])
                return np.where(ok, g, 0.0)

            def partial(self, cnt, seed):
                s_ = self.sg[b:e]
                bad = float((s_ <= 0).sum())
                with np.errstate(all="ignore"):
                    z = (x[b:e] - mu[b:e]) / s_
                    sums = [np.sum(z * z), np.sum(np.log(s_))]
                self.xa = np.zeros(e - b)
                if cnt is None or cnt.item() == 0:        # strict form: no adjoint when the global count says invalid
                    self.xa = self.xa + self.contrib(seed)
                return torch.tensor([sums[0], sums[1], 0.0, bad], dtype=torch.float64)

            def finish(self, part, n_total, seed):
                return -np.inf if part[3].item() != 0 else -0.5 * part[0].item() - part[1].item()

            def undo(self, part, seed):
                self.undone = part[3].item() != 0
                if self.undone:
                    self.xa = self.xa - self.contrib(seed)
        ops = VVV(sg)
        out["vvv"] = sharding.sharded_normal_autodiff(ops, ar, 3, n, seed, True)
        out["vvv_xa"] = (b, e, ops.xa)
        assert ops.undone is False
        ops = VVV(sg)
        out["vvv_strict"] = sharding.sharded_normal_autodiff(ops, ar, 3, n, seed, True, strict=True)
        assert out["vvv_strict"] == out["vvv"] and ops.undone is None
        # one invalid sigma in the LAST rank's shard: every rank must end with -inf and untouched (zero) adjoints,
        # through the optimistic pass + undo and through the strict pre-pass alike
        sg_bad = sg.copy()
        sg_bad[n - 2] = -0.5
        for strict in (False, True):
            ops = VVV(sg_bad)
            v = sharding.sharded_normal_autodiff(ops, ar, 3, n, seed, True, strict=strict)
            assert v == -np.inf and np.all(ops.xa == 0.0), (rank, strict)
            assert ops.undone is (None if strict else True)

        # ---- vss: scalar mu / sigma adjoints are formed from the all-reduced sums
        class VSS:
            def partial(self, cnt, seed):
                d = x[b:e] - mu_s
                return torch.tensor([np.sum(d * d), 0.0, np.sum(d), 0.0], dtype=torch.float64)

            def finish(self, part, n_total, seed):
                inv_s = 1.0 / sg_s
                z_sq = part[0].item() / (sg_s * sg_s)
                self.sigma_adj = seed * (z_sq - n_total) * inv_s
                self.mu_adj = seed * (part[2].item() * inv_s * inv_s)
                return -0.5 * z_sq - n_total * np.log(sg_s)
        ops = VSS()
        out["vss"] = sharding.sharded_normal_autodiff(ops, ar, 0, n, seed, True)
        out["vss_adj"] = (ops.mu_adj, ops.sigma_adj)

        # ---- regression: row shards, all-reduce {sum r^2, X^T r}
        rows, cols = 1003, 7
        X, y, w_true = synth.linreg_inputs(rows, cols)
        w = w_true + 0.01
        rb, re_ = sharding.shard_range(rows, rank, world)

        class LR:
            def partial(self):
                r = y[rb:re_] - X[rb:re_] @ w
                return torch.tensor(np.concatenate([[np.sum(r * r)], X[rb:re_].T @ r]), dtype=torch.float64)

            def finish(self, part, rows_total, seed):
                p = part.numpy()
                inv_s = 1.0 / sg_s
                z_sq = p[0] / (sg_s * sg_s)
                self.w_adj = seed * inv_s * inv_s * p[1:]
                self.sigma_adj = seed * (z_sq - rows_total) * inv_s
                return -0.5 * z_sq - rows_total * np.log(sg_s)
        ops = LR()
        out["lr"] = sharding.sharded_linreg_autodiff(ops, ar, rows, 1.0)
        out["lr_adj"] = (ops.w_adj, ops.sigma_adj)
        # ---- sum(exp(x)): 1-double exchange, adjoints shard-local
        xs = synth.sum_inputs(n)

        class SUM:
            def run(self):
                self.adj = seed * np.exp(xs[b:e])
                return torch.tensor([np.sum(np.exp(xs[b:e]))], dtype=torch.float64)
        ops = SUM()
        out["sum"] = sharding.sharded_sum_autodiff(ops, ar).item()
        out["sum_adj"] = ops.adj

        # ---- prod(x) with one exact zero living in rank 1's shard: {prod of non-zeros, #zeros} all-gathered
        xp = np.exp(np.linspace(-1e-3, 1e-3, n)); xp[n - 5] = 0.0

        class PROD:
            def reduce(self):
                sh = xp[b:e]
                return torch.tensor([np.prod(sh[sh != 0]), float((sh == 0).sum())], dtype=torch.float64)

            def scatter(self, p_nz, zeros):
                sh = xp[b:e]
                with np.errstate(divide="ignore", invalid="ignore"):
                    self.adj = np.where(sh == 0, seed * (0.0 if zeros > 1 else p_nz), seed * ((0.0 if zeros > 0 else p_nz) / sh))

        def all_gather(t):
            outl = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(outl, t)
            return outl
        ops = PROD()
        out["prod"] = sharding.sharded_prod_autodiff(ops, all_gather, world)
        out["prod_adj"] = ops.adj

        # ---- network: batch shards, one 26-double all-reduce
        import oracle_py as O
        import kats
        B = 2001
        Xn, yn = synth.mlp3_inputs(B)
        bb, be = sharding.shard_range(B, rank, world)

        class NN:
            def run(self):
                g = np.zeros(25)
                loss = O.mlp3(np.asfortranarray(Xn[bb:be]), np.ascontiguousarray(yn[bb:be]), kats.NN_PARM, g) if be > bb else 0.0
                return torch.tensor(np.concatenate([g, [loss]]), dtype=torch.float64)
        out["nn"] = sharding.sharded_mlp3_autodiff(NN(), ar).numpy()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_reductions_two_ranks_gloo():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as O
    import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    n = 10007
    x, mu, sg = synth.normal_inputs(n)
    xa, ma, sa = np.zeros(n), np.zeros(n), np.zeros(n)
    ov = O.normal_lpdf(x, mu, sg, xa, ma, sa, seed=0.75)
    for r in range(2):
        assert abs(res[r]["vvv"] - ov) <= 1e-12 * abs(ov)          # every rank holds the global value
        b, e, got = res[r]["vvv_xa"]
        np.testing.assert_array_equal(got, xa[b:e])                # shard-local vector adjoints, bit-exact
    assert res[0]["vvv_xa"][1] == res[1]["vvv_xa"][0] and res[1]["vvv_xa"][1] == n   # ranges tile [0, n)

    xa2, m1, s1 = np.zeros(n), np.zeros(1), np.zeros(1)
    ov = O.normal_lpdf(x, np.array([0.1]), np.array([1.3]), xa2, m1, s1, seed=0.75)
    for r in range(2):
        assert abs(res[r]["vss"] - ov) <= 1e-12 * abs(ov)
        assert abs(res[r]["vss_adj"][0] - m1[0]) <= 1e-12 * abs(m1[0])
        assert abs(res[r]["vss_adj"][1] - s1[0]) <= 1e-12 * abs(s1[0])

    rows, cols = 1003, 7
    X, y, w_true = synth.linreg_inputs(rows, cols)
    wa, sa1 = np.zeros(cols), np.zeros(1)
    ov = O.linreg(X, y, w_true + 0.01, wa, np.array([1.3]), sa1)
    for r in range(2):
        assert abs(res[r]["lr"] - ov) <= 1e-12 * abs(ov)
        np.testing.assert_allclose(res[r]["lr_adj"][0], wa, rtol=1e-11, atol=1e-9 * np.abs(wa).max())
        assert abs(res[r]["lr_adj"][1] - sa1[0]) <= 1e-11 * abs(sa1[0])


    # sum / prod / network
    xs = synth.sum_inputs(n)
    oa = np.zeros(n)
    ov = O.sum_unary("exp", xs, oa, seed=0.75)
    xp = np.exp(np.linspace(-1e-3, 1e-3, n)); xp[n - 5] = 0.0
    # prod.hpp:172-212 through the oracle's expression interface
    eo = O.OracleExpr()
    xv = eo.var(xp)
    eo.bind(eo.prod(xv))
    pv = eo.autodiff(0.75)[0]
    pa = eo.adj(xv)
    import kats
    B = 2001
    Xn, yn = synth.mlp3_inputs(B)
    g = np.zeros(25)
    lv = O.mlp3(Xn, yn, kats.NN_PARM, g)
    for r in range(2):
        b, e, _ = res[r]["vvv_xa"]
        assert abs(res[r]["sum"] - ov) <= 1e-12 * abs(ov)
        np.testing.assert_allclose(res[r]["sum_adj"], oa[b:e], rtol=1e-15)
        assert res[r]["prod"] == pv == 0.0
        np.testing.assert_allclose(res[r]["prod_adj"], pa[b:e], rtol=1e-12, atol=0)
        np.testing.assert_allclose(res[r]["nn"][:25], g, rtol=1e-11, atol=1e-9 * np.abs(g).max())
        assert abs(res[r]["nn"][25] - lv) <= 1e-12 * abs(lv)


def test_shard_ranges_tile_and_stay_aligned():
    from fastad_b200 import sharding
    for n in (0, 1, 5, 1 << 20, (1 << 26) + 3):
        for world in (1, 2, 4, 8):
            prev = 0
            for r in range(world):
                b, e = sharding.shard_range(n, r, world)
                assert b == prev and b <= e <= n and (b % 4 == 0 or b == n)
                prev = e
            assert prev == n
