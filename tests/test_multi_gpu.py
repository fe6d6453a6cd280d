"""-m gpu tests that need >= 2 GPUs on the box (skipped otherwise): the sharded evaluations across real devices.
  * torchrun, one process per GPU: tests/sharded_gate.py through the fused exchange and through partial + all-reduce +
    finish, a 10^4-exchange soak of the mailbox all-reduce, tools/p2p_probe.py (bit-identity vs the rank-ordered sum);
  * fault injection: a rank that leaves out an exchange must make the run FAIL (timeout -> NaN -> assertion), never
    produce a plausible number;
  * ONE process driving two GPUs through fadb_p2p_create_local (plain peer access, no IPC)."""
import ctypes as C
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(n, script, *args, env=None, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_port()), os.path.join(ROOT, script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, **(env or {})))


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_sharded_gate_and_soak_torchrun():
    n = min(_ngpu(), 8)
    p = _torchrun(n, "tools/sharded_gate_run.py", "--soak", "10000")
    assert p.returncode == 0 and "SHARDED GATE OK" in p.stdout, (p.stdout[-2000:], p.stderr[-2000:])


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_missing_rank_fails_loudly():
    # rank 1 leaves out one exchange: rank 0's wait times out (1.5 s), the value is NaN, the gate asserts, exit != 0
    p = _torchrun(2, "tools/sharded_gate_run.py", "--soak", "0", "--fault-rank", "1", env={"FADB_P2P_TIMEOUT_MS": "1500"}, timeout=300)
    assert p.returncode != 0 and "SHARDED GATE OK" not in p.stdout, p.stdout[-2000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_p2p_probe_torchrun():
    p = _torchrun(2, "tools/p2p_probe.py")
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout.count("bit-identical to the rank-ordered sum: True") == 2, p.stdout[-2000:]
    assert "False" not in p.stdout, p.stdout[-2000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_one_process_two_gpus_peer_access():
    """one host thread drives both devices; the exchange runs over plain peer access."""
    import torch
    import fastad_b200 as F
    L = F.load()
    devs = (C.c_int * 2)(0, 1)
    F.check(L.fadb_p2p_create_local(2, devs))
    P = lambda t: C.c_void_p(t.data_ptr())
    n = (1 << 20) + 4
    xs, adjs, vals = [], [], []
    for d in (0, 1):
        torch.cuda.set_device(d)
        F.check(L.fadb_init(d))
        F.use_torch_stream()
        x = ((torch.arange(n, device=f"cuda:{d}") * 3 + d) % 11 - 5).to(torch.float64)
        xs.append(x); adjs.append(torch.zeros(n, dtype=torch.float64, device=f"cuda:{d}"))
        vals.append(torch.zeros(1, dtype=torch.float64, device=f"cuda:{d}"))
    for rep in range(3):
        for d in (0, 1):                      # enqueue both ranks' fused launches; each waits for the other's flag on the device
            torch.cuda.set_device(d)
            F.use_torch_stream()
            F.check(L.fadb_sum_unary_allreduce_async(F._op_code(("pow", 2)), n, P(xs[d]), P(adjs[d]), 1.0, F.OVERWRITE_ADJ, P(vals[d])))
    want = float((xs[0] * xs[0]).sum()) + float((xs[1] * xs[1]).sum())
    for d in (0, 1):
        torch.cuda.set_device(d)
        torch.cuda.synchronize()
        assert float(vals[d][0]) == want
        assert torch.equal(adjs[d], 2.0 * xs[d])
        e = C.c_ulonglong(1)
        F.check(L.fadb_p2p_status(C.byref(e)))
        assert e.value == 0
    for d in (0, 1):
        torch.cuda.set_device(d)
        F.check(L.fadb_p2p_destroy())
    torch.cuda.set_device(0)
