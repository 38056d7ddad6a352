"""The C restatement (timed CPU baseline) must agree with the numpy oracle (which is pinned by the
reference's golden vectors).  CPU only."""
import numpy as np

from oracle import oracle_c as C
from oracle import oracle_np as O


def rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(b), 1e-30)


def test_c_oracle_matches_numpy_oracle():
    rng = np.random.default_rng(3)
    D, N = 32, 257
    f = np.float32
    x = rng.standard_normal((D, N)).astype(f)
    w, u, b = (rng.standard_normal(D) / np.sqrt(D)).astype(f), (rng.standard_normal(D) / np.sqrt(D)).astype(f), f(0.3)
    for nt in (1, 4):
        y, lj = C.planar_fwd(w, u, b, x, nthreads=nt)
        yo, ljo = O.planar_forward(w, u, b, x)
        assert rel(y, yo) < 2e-6 and rel(lj, ljo) < 2e-5
    z0 = rng.standard_normal(D).astype(f)
    y, lj = C.radial_fwd(0.4, -0.2, z0, x)
    yo, ljo = O.radial_forward(f(0.4), f(-0.2), z0, x)
    assert rel(y, yo) < 2e-6 and rel(lj, ljo) < 2e-5
    K = 8
    W, H, Dv = O.rqs_params(rng.standard_normal((D, K)).astype(f), rng.standard_normal((D, K)).astype(f),
                            rng.standard_normal((D, K - 1)).astype(f), 3.0)
    y, lj = C.rqs_fwd(W, H, Dv, x * f(1.5))
    yo, ljo = O.rqs_forward(W, H, Dv, x * f(1.5))
    assert rel(y, yo) < 2e-6 and rel(lj, ljo) < 2e-5
    bn = O.BatchNormParams(*(rng.standard_normal(D).astype(f) * f(0.1) for _ in range(3)),
                           rng.uniform(0.5, 1.5, D).astype(f), f(1e-5), f(0.1))
    y, lj = C.batchnorm(bn.b, bn.logs, bn.m, bn.v, bn.eps, x)
    yo, ljo = O.batchnorm_forward(bn, x)
    assert rel(y, yo) < 2e-6 and rel(lj, ljo) < 2e-5
    xi, lji = C.batchnorm(bn.b, bn.logs, bn.m, bn.v, bn.eps, y, inverse=True)
    assert rel(xi, x) < 2e-6 and rel(lji, -ljo) < 2e-5
    n1 = D // 2
    idx1, idx2 = np.arange(1, n1 + 1), np.arange(n1 + 1, D + 1)
    Wc, c = (rng.standard_normal((2 * n1, D - n1)) * 0.1).astype(f), (rng.standard_normal(2 * n1) * 0.1).astype(f)
    y, lj = C.coupling_affine(idx1, idx2, Wc, c, x)
    yo, ljo = O.coupling_affine_forward(idx1, idx2, Wc, c, x)
    assert rel(y, yo) < 2e-6 and rel(lj, ljo) < 2e-5
    xi, lji = C.coupling_affine(idx1, idx2, Wc, c, y, inverse=True)
    assert rel(xi, x) < 5e-6 and rel(lji, -ljo) < 2e-5
    lp = C.mvnormal_diag_logpdf(None, None, x)
    assert rel(lp, O.mvnormal_diag_logpdf(None, None, x)) < 2e-6
    sg = rng.uniform(0.5, 2, D).astype(f)
    lp = C.mvnormal_diag_logpdf(z0, sg, x)
    assert rel(lp, O.mvnormal_diag_logpdf(z0, sg, x)) < 2e-6
    layers = [((rng.standard_normal(D) / np.sqrt(D)).astype(f), (rng.standard_normal(D) / np.sqrt(D)).astype(f),
               f(rng.standard_normal())) for _ in range(8)]
    y, lj = C.planar_chain_fwd(layers, x, nthreads=2)
    yo, ljo = O.chain_forward([O.Layer("planar", dict(w=w_, u=u_, b=b_)) for (w_, u_, b_) in layers], x)
    assert rel(y, yo) < 5e-6 and rel(lj, ljo) < 5e-5
