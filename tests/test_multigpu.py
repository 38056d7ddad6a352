"""N > 1 on real GPUs: columns shard with no data-path collective, one NCCL sum at the end (needs >= 2 GPUs;
skipped on the 1-GPU box).  The CPU world_size-2 gloo test of the same logic is tests/test_host_logic.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_logpdf_nccl_sum():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "mgpu ok" in r.stdout


def test_single_process_clique_sum():
    """One process, every visible GPU: b2b_comm_init_all + b2b_allreduce_sum_f64_all (the single-Julia-session shape of
    SURVEY §8(b)): per-device logpdf sums of a column-sharded batch add up to the one-device total."""
    import numpy as np
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    import bijectors_jl_b200 as B
    from bijectors_jl_b200.distributed import Clique, shard_columns

    rng = np.random.default_rng(0)
    D, N = 64, 40000
    w, u, b = (rng.standard_normal(D) / 8).astype(np.float32), (rng.standard_normal(D) / 8).astype(np.float32), rng.standard_normal(1).astype(np.float32)
    y = rng.standard_normal((D, N)).astype(np.float32)
    clique = Clique()
    totals = []
    for d in range(n):
        torch.cuda.set_device(d)
        lo, hi = shard_columns(N, d, n)
        td = B.transformed(B.MvNormal(D, device=f"cuda:{d}"), B.PlanarLayer(w, u, b, device=f"cuda:{d}"))
        t, _ = B.logpdf_sum(td, B.from_numpy(np.ascontiguousarray(y[:, lo:hi]), device=f"cuda:{d}"))
        totals.append(t.reshape(1))
    clique.allreduce_sum_(totals)
    torch.cuda.set_device(0)
    td0 = B.transformed(B.MvNormal(D, device="cuda:0"), B.PlanarLayer(w, u, b, device="cuda:0"))
    ref, _ = B.logpdf_sum(td0, B.from_numpy(y, device="cuda:0"))
    for d in range(n):
        assert abs(float(totals[d]) - float(ref)) <= 1e-9 * abs(float(ref)), (d, float(totals[d]), float(ref))
    clique.close()
