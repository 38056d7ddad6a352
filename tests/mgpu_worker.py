"""torchrun worker: column-sharded TransformedDistribution logpdf with the single NCCL sum.
Launched by tests/test_multigpu.py:  torchrun --nproc-per-node N tests/mgpu_worker.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    import bijectors_jl_b200 as B
    from bijectors_jl_b200.distributed import Communicator, shard_columns, sharded_logpdf_sum
    from oracle import oracle_np as O

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Communicator()
    assert comm.handle is not None, "libb2b NCCL communicator was not created"
    f32 = np.float32
    rng = np.random.default_rng(11)  # identical on every rank
    D, N = 64, 40_001
    params = [((rng.standard_normal(D) / 8).astype(f32), (rng.standard_normal(D) / 8).astype(f32), rng.standard_normal(1).astype(f32))
              for _ in range(3)]
    flow = B.Composed(*[B.PlanarLayer(w, u, b) for (w, u, b) in params])
    td = B.transformed(B.MvNormal(D), flow)
    y = rng.standard_normal((D, N)).astype(f32)
    lo, hi = shard_columns(N, rank, world)
    total = sharded_logpdf_sum(td, B.from_numpy(y[:, lo:hi]), comm)  # local fused kernels + ONE 8-byte all-reduce
    ref = O.transformed_logpdf([O.Layer("planar", dict(w=w, u=u, b=b)) for (w, u, b) in params], None, None,
                               y.astype(np.float64)).sum()
    err = abs(float(total) - ref) / abs(ref)
    assert err <= 1e-5, (float(total), ref, err)
    # every rank holds the same total
    t = total.reshape(1).clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == float(total)
    # training-mode BatchNorm over the sharded batch: the second (2D+1 doubles) all-reduce of the path
    Db = 32
    xb = (rng.standard_normal((Db, N)) * 1.7 + 0.3).astype(f32)
    bn = B.InvertibleBatchNorm(Db, training=True)
    yb, ljb = bn.train_forward(B.from_numpy(xb[:, lo:hi]), comm)
    yo, ljo, (m1, v1) = O.batchnorm_forward(O.BatchNormParams.default(Db, np.float64), xb.astype(np.float64), training=True)

    def rel(a, b_):
        return np.linalg.norm(np.asarray(a, np.float64) - b_) / np.linalg.norm(b_)

    assert rel(B.to_numpy(yb), yo[:, lo:hi]) <= 1e-5 and rel(B.to_numpy(ljb), ljo[lo:hi]) <= 1e-5
    assert rel(B.to_numpy(bn.m), m1) <= 1e-5 and rel(B.to_numpy(bn.v), v1) <= 1e-5
    # reverse mode over column shards: local VJP kernels + ONE all-reduce of the 2·L·D + L parameter cotangents
    from bijectors_jl_b200.distributed import sharded_planar_chain_vjp

    x = rng.standard_normal((D, N)).astype(f32)
    ybar, ljbar = rng.standard_normal((D, N)).astype(f32), rng.standard_normal(N).astype(f32)
    xbar, grads = sharded_planar_chain_vjp(flow, B.from_numpy(x[:, lo:hi]), B.from_numpy(ybar[:, lo:hi]),
                                           torch.from_numpy(ljbar[lo:hi]).cuda(), comm)
    p64 = [tuple(a.astype(np.float64) for a in p) for p in params]
    xb_o, g_o = O.planar_chain_vjp(p64, x.astype(np.float64), ybar.astype(np.float64), ljbar.astype(np.float64))
    assert rel(B.to_numpy(xbar), xb_o[:, lo:hi]) <= 1e-5
    for l in range(len(params)):
        assert rel(grads[l]["w"].cpu().numpy(), g_o[l][0]) <= 2e-5 and rel(grads[l]["u"].cpu().numpy(), g_o[l][1]) <= 2e-5
        assert abs(float(grads[l]["b"]) - float(g_o[l][2])) <= 2e-5 * max(abs(float(g_o[l][2])), np.sqrt(N))
    comm.close()
    dist.barrier()
    if rank == 0:
        print(f"mgpu ok: world={world} total={float(total):.6f} ref={ref:.6f} rel_err={err:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
