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


def sharded_normal_autodiff(ops, all_reduce, shape_code, n_total, seed, check_sigma):
    """normal_adj_log_pdf over element-range shards.
    ops.sigma_check() -> count buffer; ops.partial(count_or_None) -> 4-slot partial buffer;
    ops.finish(partial, n_total) -> value.  Buffers are whatever `all_reduce` accepts."""
    cnt = None
    if (shape_code & 2) and check_sigma:
        cnt = ops.sigma_check()
        all_reduce(cnt)                 # 1 double: #sigma_i <= 0 over ALL shards
    part = ops.partial(cnt, seed)
    all_reduce(part)                    # 4 doubles: the one exchange of the evaluation
    return ops.finish(part, n_total, seed)


def sharded_linreg_autodiff(ops, all_reduce, rows_total, seed):
    """normal(y, dot(X, w), sigma) over row-range shards: all-reduce {sum r^2, X^T r}."""
    part = ops.partial()
    all_reduce(part)
    return ops.finish(part, rows_total, seed)
