"""Element-range sharding of the reductions across ranks.

One process per GPU.  A rank owns a contiguous element (or row) range; vector adjoints need only
the seed, so every rank runs the same fused pass on its range and the ONLY data that crosses
ranks is a handful of partial sums (one all-reduce), after which each rank forms the value and
the adjoints of scalar operands.  The functions here are backend-agnostic: `ops` supplies
`partial(...)`/`finish(...)` (the C-ABI calls on a GPU; a numpy stand-in in the CPU gloo test) and
`all_reduce` is torch.distributed's.
"""


def shard_range(n, rank, world):
    """Contiguous, balanced [begin, end) of rank's elements; aligned to 4 elements (32 bytes) so
    every shard keeps the 256-bit load path."""
    per = -(-n // world)
    per = -(-per // 4) * 4
    b = min(n, rank * per)
    e = min(n, b + per)
    return b, e


def sharded_normal_autodiff(ops, all_reduce, shape_code, n_total, seed, check_sigma, strict=False):
    """normal_adj_log_pdf over element-range shards.
    ops.partial(count_or_None, seed) -> 4-slot partial buffer {.., .., .., #sigma_i <= 0 of the shard};
    ops.finish(partial, n_total, seed) -> value; buffers are whatever `all_reduce` accepts.
    A vector sigma must be valid on ALL shards for any adjoint to be written (normal.hpp:440-457):
      default: every shard runs the optimistic pass (elements with an invalid sigma_i add nothing and are counted), the
               count travels with the ONE exchange of the evaluation, and ops.undo(partial, seed) takes the shard's
               vector adjoints back when the global count is non-zero (it returns at once otherwise);
      strict:  ops.sigma_check() -> count buffer, all-reduced FIRST (a second exchange), then the pass either writes
               adjoints or does not (bit-exact "untouched")."""
    guarded = bool(shape_code & 2) and check_sigma
    cnt = None
    if guarded and strict:
        cnt = ops.sigma_check()
        all_reduce(cnt)                 # 1 double: #sigma_i <= 0 over ALL shards
    part = ops.partial(cnt, seed)
    all_reduce(part)                    # 4 doubles: the one exchange of the evaluation
    val = ops.finish(part, n_total, seed)
    if guarded and not strict:
        ops.undo(part, seed)
    return val


def sharded_linreg_autodiff(ops, all_reduce, rows_total, seed):
    """normal(y, dot(X, w), sigma) over row-range shards: all-reduce {sum r^2, X^T r}."""
    part = ops.partial()
    all_reduce(part)
    return ops.finish(part, rows_total, seed)


def sharded_sum_autodiff(ops, all_reduce):
    """sum(f(x)) over element-range shards: the root seed is known up front, so every rank's fused pass
    already wrote its adjoints; the only exchange is the 1-double value.
    ops.run() -> 1-slot value buffer (adjoints of the shard updated as a side effect)."""
    val = ops.run()
    all_reduce(val)
    return val


def sharded_prod_autodiff(ops, all_gather, world):
    """prod(x) over element-range shards (prod.hpp:172-212 semantics incl. exact zeros): every rank reduces
    {product of its non-zeros, its number of zeros}; the pairs are all-gathered and multiplied / added in
    rank order, then each rank scatters adj_i from the GLOBAL pair.
    ops.reduce() -> 2-slot buffer; ops.scatter(p_nz_global, zeros_global) -> None."""
    mine = ops.reduce()
    pairs = all_gather(mine)            # list of `world` 2-slot buffers
    p_nz, zeros = 1.0, 0.0
    for r in range(world):
        p_nz *= float(pairs[r][0])
        zeros += float(pairs[r][1])
    ops.scatter(p_nz, zeros)
    return 0.0 if zeros > 0 else p_nz


def sharded_mlp3_autodiff(ops, all_reduce):
    """The network of example/ceres_3layer_neural_net over batch shards: rows are independent, the 25
    parameter gradients and the loss are sums over rows: one 26-double all-reduce.
    ops.run() -> 26-slot buffer {grad[25], loss} of the local rows."""
    part = ops.run()
    all_reduce(part)
    return part


class PeerAllReduce:
    """The small all-reduce over NVLink peer memory (csrc/p2p.cu) for the ranks of one node: mailboxes are
    exchanged once through torch.distributed, after that `pa(t)` sums the float64 CUDA tensor `t` (<= 128
    elements) over the ranks in place with one tiny kernel per rank on the library stream."""

    def __init__(self, world, rank):
        import ctypes as C
        import torch
        import torch.distributed as dist
        import fastad_b200 as F
        self.F, self.L = F, F.lib()
        h = C.create_string_buffer(64)
        F.check(self.L.fadb_p2p_create(world, rank, h))
        mine = torch.frombuffer(bytearray(h.raw), dtype=torch.uint8).cuda()
        allh = [torch.empty(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(allh, mine)
        blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
        F.check(self.L.fadb_p2p_connect(C.c_char_p(blob)))
        dist.barrier()

    def __call__(self, t):
        import ctypes as C
        assert t.is_cuda and t.dtype.is_floating_point and t.element_size() == 8 and t.is_contiguous() and t.numel() <= 128
        self.F.check(self.L.fadb_p2p_allreduce(C.c_void_p(t.data_ptr()), t.numel()))
        return t

    def healthy(self):
        import ctypes as C
        e = C.c_ulonglong(0)
        self.F.check(self.L.fadb_p2p_status(C.byref(e)))
        return e.value == 0

    def close(self):
        self.F.check(self.L.fadb_p2p_destroy())
