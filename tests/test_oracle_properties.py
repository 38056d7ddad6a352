"""Property tests the reference uses to pin its own behaviour (test/bijectors/utils.jl:43-62,
test/normalising_flows.jl), restated for the numpy oracle.  CPU only."""
import numpy as np
import pytest

from oracle import oracle_np as O


def num_logabsdet(f, x, h=1e-6):
    """log|det J| of f at vector x by central differences (stands in for ForwardDiff.jacobian)."""
    n = x.size
    J = np.zeros((n, n))
    for j in range(n):
        e = np.zeros(n)
        e[j] = h
        J[:, j] = (f(x + e) - f(x - e)) / (2 * h)
    return np.linalg.slogdet(J)[1]


def make_layers(rng, D):
    K, B = 6, 3.0
    W, H, Dv = O.rqs_params(rng.standard_normal((D, K)), rng.standard_normal((D, K)), rng.standard_normal((D, K - 1)), B)
    n1 = D // 2
    idx1 = np.arange(1, n1 + 1)
    idx2 = np.arange(n1 + 1, D + 1)
    bn = O.BatchNormParams(
        rng.standard_normal(D) * 0.1, rng.standard_normal(D) * 0.1, rng.standard_normal(D) * 0.1,
        rng.uniform(0.5, 1.5, D), np.float64(1e-5), np.float64(0.1))
    perm = rng.permutation(D) + 1
    return {
        "planar": O.Layer("planar", dict(w=rng.standard_normal(D), u=rng.standard_normal(D), b=rng.standard_normal(1))),
        "radial": O.Layer("radial", dict(alpha_raw=rng.standard_normal(1), beta=rng.standard_normal(1), z0=rng.standard_normal(D))),
        "rqs": O.Layer("rqs", dict(widths=W, heights=H, derivs=Dv)),
        "coupling_affine": O.Layer("coupling_affine", dict(idx1=idx1, idx2=idx2, W=rng.standard_normal((2 * n1, D - n1)) * 0.3, c=rng.standard_normal(2 * n1) * 0.1)),
        "batchnorm": O.Layer("batchnorm", dict(bn=bn)),
        "permute": O.Layer("permute", dict(A=O.permute_matrix_from_indices(perm.tolist()))),
        "stacked": O.Layer("stacked", dict(ops=[(O.EW.EXP, 0.0), (O.EW.SCALE, -1.7), (O.EW.SHIFT, 0.3)], ranges=[(1, 2), (3, D - 1), (D, D)])),
        "leaky_relu": O.Layer("stacked", dict(ops=[(O.EW.LEAKY_RELU, 0.1)], ranges=[(1, D)])),
    }


KINDS = ["planar", "radial", "rqs", "coupling_affine", "batchnorm", "permute", "stacked", "leaky_relu"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_logjac_matches_numerical_jacobian(kind, seed):
    # test/normalising_flows.jl:17-20,29-32,78-81; test/bijectors/utils.jl:43-50
    rng = np.random.default_rng(seed)
    D = 6
    L = make_layers(rng, D)[kind]
    x = rng.standard_normal(D)
    y, lj = L.forward(x)
    ref = num_logabsdet(lambda v: L.forward(v)[0], x)
    assert float(lj) == pytest.approx(ref, rel=2e-6, abs=2e-7)
    # batch: per-column logjac equals the single-column result (column independence)
    X = rng.standard_normal((D, 9))
    Y, LJ = L.forward(X)
    for n in range(X.shape[1]):
        yn, ljn = L.forward(X[:, n].copy())
        np.testing.assert_allclose(Y[:, n], yn, rtol=1e-12, atol=1e-14)
        assert float(LJ[n]) == pytest.approx(float(ljn), rel=1e-11, abs=1e-13)
    # permuting columns permutes outputs
    p = rng.permutation(9)
    Yp, LJp = L.forward(X[:, p])
    np.testing.assert_allclose(Yp, Y[:, p], rtol=1e-13, atol=1e-15)  # BLAS may reassociate
    np.testing.assert_allclose(LJp, LJ[p], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_inverse_roundtrip_and_ires(kind, seed):
    # test/bijectors/utils.jl:53-62: inverse∘forward ≈ id and ires == (x, -logjac)
    rng = np.random.default_rng(100 + seed)
    D = 8
    L = make_layers(rng, D)[kind]
    X = rng.standard_normal((D, 33))
    Y, LJ = L.forward(X)
    Xi, LJi = L.inverse(Y)
    np.testing.assert_allclose(Xi, X, rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(LJi, -LJ, rtol=1e-7, atol=1e-9)


def test_chain_and_logpdf():
    rng = np.random.default_rng(5)
    D = 6
    Ls = make_layers(rng, D)
    chain = [Ls[k] for k in ["planar", "batchnorm", "coupling_affine", "permute", "radial", "rqs"]]
    X = rng.standard_normal((D, 17))
    Y, LJ = O.chain_forward(chain, X)
    ref = [num_logabsdet(lambda v: O.chain_forward(chain, v)[0], X[:, n].copy()) for n in range(3)]
    np.testing.assert_allclose(LJ[:3], ref, rtol=5e-6, atol=5e-7)
    Xi, LJi = O.chain_inverse(chain, Y)
    np.testing.assert_allclose(Xi, X, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(LJi, -LJ, rtol=1e-6, atol=1e-8)
    # logpdf(td, y) = logpdf(d, x) - logjac_fwd(x)   (test/normalising_flows.jl:97-111)
    mu, sigma = rng.standard_normal(D) * 0.1, rng.uniform(0.5, 2.0, D)
    lp = O.transformed_logpdf(chain, mu, sigma, Y)
    np.testing.assert_allclose(lp, O.mvnormal_diag_logpdf(mu, sigma, X) - LJ, rtol=1e-6, atol=1e-8)
    # MvNormal closed form vs scipy
    from scipy.stats import multivariate_normal
    ref_lp = multivariate_normal(mean=mu, cov=np.diag(sigma ** 2)).logpdf(X.T)
    np.testing.assert_allclose(O.mvnormal_diag_logpdf(mu, sigma, X), ref_lp, rtol=1e-12)


def test_float32_oracle_close_to_float64():
    rng = np.random.default_rng(9)
    D = 16
    Ls = make_layers(rng, D)
    X = rng.standard_normal((D, 64))
    for k in KINDS:
        if k == "leaky_relu":
            continue  # parameter-free in float terms; covered by the stacked case
        L64 = Ls[k]
        p32 = {}
        for key, v in L64.params.items():
            if isinstance(v, np.ndarray) and v.dtype == np.float64 and key != "A":
                p32[key] = v.astype(np.float32)
            elif isinstance(v, O.BatchNormParams):
                p32[key] = O.BatchNormParams(*(a.astype(np.float32) for a in (v.b, v.logs, v.m, v.v)), np.float32(v.eps), np.float32(v.mtm))
            else:
                p32[key] = v
        L32 = O.Layer(k, p32)
        Y64, LJ64 = L64.forward(X)
        Y32, LJ32 = L32.forward(X.astype(np.float32))
        assert Y32.dtype == np.float32 and np.asarray(LJ32).dtype == np.float32
        assert np.linalg.norm(Y32 - Y64) <= 2e-6 * np.linalg.norm(Y64), k
        assert np.linalg.norm(LJ32 - LJ64) <= 2e-5 * max(np.linalg.norm(LJ64), 1e-3), k


def test_log1pexp_tails():
    import mpmath as mp
    for x in [-100.0, -37.0, -10.0, -1.0, 0.0, 1.0, 10.0, 18.0, 33.3, 50.0, 700.0]:
        ref = float(mp.log1p(mp.e ** mp.mpf(x)))
        assert float(O.log1pexp(np.float64(x))) == pytest.approx(ref, rel=1e-14, abs=1e-300)


def test_planar_chain_vjp_matches_finite_differences():
    """The oracle VJP (reverse-mode restatement) against central finite differences of the pinned forward oracle."""
    rng = np.random.default_rng(11)
    D, N, L = 6, 5, 3
    params = [(rng.standard_normal(D) / np.sqrt(D), rng.standard_normal(D) / np.sqrt(D), rng.standard_normal(1)) for _ in range(L)]
    x = rng.standard_normal((D, N))
    ybar, ljbar = rng.standard_normal((D, N)), rng.standard_normal(N)

    def loss(ps, xx):
        z, lj = xx, np.zeros(N)
        for (w, u, b) in ps:
            z, l1 = O.planar_forward(w, u, b, z)
            lj = lj + l1
        return float(np.sum(z * ybar) + np.sum(lj * ljbar))

    xbar, grads = O.planar_chain_vjp(params, x, ybar, ljbar)
    eps = 1e-6
    fd = np.zeros_like(x)
    for i in range(D):
        for n in range(N):
            xp, xm = x.copy(), x.copy()
            xp[i, n] += eps
            xm[i, n] -= eps
            fd[i, n] = (loss(params, xp) - loss(params, xm)) / (2 * eps)
    assert np.allclose(xbar, fd, rtol=1e-6, atol=1e-8)
    for l in range(L):
        for k, name in enumerate(("w", "u", "b")):
            base = np.asarray(params[l][k], dtype=np.float64)
            fdp = np.zeros_like(base)
            for i in range(base.size):
                pp, pm = base.copy(), base.copy()
                pp[i] += eps
                pm[i] -= eps
                ps_p = [tuple(pp if (ll == l and kk == k) else params[ll][kk] for kk in range(3)) for ll in range(L)]
                ps_m = [tuple(pm if (ll == l and kk == k) else params[ll][kk] for kk in range(3)) for ll in range(L)]
                fdp[i] = (loss(ps_p, x) - loss(ps_m, x)) / (2 * eps)
            got = np.atleast_1d(np.asarray(grads[l][k], dtype=np.float64))
            assert np.allclose(got, fdp, rtol=1e-5, atol=1e-7), (l, name, got, fdp)


def test_find_alpha_partials_match_finite_differences():
    """ext/BijectorsChainRulesCoreExt.jl:42-46 (the implicit-function rule the reference registers)."""
    rng = np.random.default_rng(12)
    for _ in range(20):
        t, c, b = rng.standard_normal() * 2, rng.uniform(-0.9, 2.0), rng.standard_normal()
        a = float(O.find_alpha(np.float64(t), np.float64(c), np.float64(b)))
        px = O.find_alpha_partials(a, c, b)
        eps = 1e-6
        fd = [
            (float(O.find_alpha(np.float64(t + eps), np.float64(c), np.float64(b))) - float(O.find_alpha(np.float64(t - eps), np.float64(c), np.float64(b)))) / (2 * eps),
            (float(O.find_alpha(np.float64(t), np.float64(c + eps), np.float64(b))) - float(O.find_alpha(np.float64(t), np.float64(c - eps), np.float64(b)))) / (2 * eps),
            (float(O.find_alpha(np.float64(t), np.float64(c), np.float64(b + eps))) - float(O.find_alpha(np.float64(t), np.float64(c), np.float64(b - eps)))) / (2 * eps),
        ]
        assert np.allclose(px, fd, rtol=1e-5, atol=1e-7)


def test_planar_inverse_chain_vjp_matches_finite_differences():
    """Gradient of the inverse chain (the logpdf / NLL training path, docs/src/flows.md:66-100): the oracle VJP built
    on the reference's implicit find_alpha rule against central finite differences of the pinned inverse oracle."""
    rng = np.random.default_rng(21)
    D, N, L = 5, 4, 3
    params = [(rng.standard_normal(D) / np.sqrt(D), rng.standard_normal(D) / np.sqrt(D), rng.standard_normal(1)) for _ in range(L)]
    y = rng.standard_normal((D, N))
    xbar, ljbar = rng.standard_normal((D, N)), rng.standard_normal(N)

    def loss(ps, yy):
        z, lj = yy, np.zeros(N)
        for (w, u, b) in reversed(ps):
            z, l1 = O.planar_inverse(w, u, b, z)
            lj = lj + l1
        return float(np.sum(z * xbar) + np.sum(lj * ljbar))

    ybar, grads = O.planar_inverse_chain_vjp(params, y, xbar, ljbar)
    eps = 1e-6
    fd = np.zeros_like(y)
    for i in range(D):
        for n in range(N):
            yp, ym = y.copy(), y.copy()
            yp[i, n] += eps
            ym[i, n] -= eps
            fd[i, n] = (loss(params, yp) - loss(params, ym)) / (2 * eps)
    assert np.allclose(ybar, fd, rtol=2e-6, atol=1e-7)
    for l in range(L):
        for k, name in enumerate(("w", "u", "b")):
            base = np.asarray(params[l][k], dtype=np.float64)
            fdp = np.zeros_like(base)
            for i in range(base.size):
                pp, pm = base.copy(), base.copy()
                pp[i] += eps
                pm[i] -= eps
                ps_p = [tuple(pp if (ll == l and kk == k) else params[ll][kk] for kk in range(3)) for ll in range(L)]
                ps_m = [tuple(pm if (ll == l and kk == k) else params[ll][kk] for kk in range(3)) for ll in range(L)]
                fdp[i] = (loss(ps_p, y) - loss(ps_m, y)) / (2 * eps)
            got = np.atleast_1d(np.asarray(grads[l][k], dtype=np.float64))
            assert np.allclose(got, fdp, rtol=2e-5, atol=2e-7), (l, name, got, fdp)


def test_radial_chain_vjp_matches_finite_differences():
    rng = np.random.default_rng(31)
    D, N, L = 5, 4, 3
    params = [(rng.standard_normal(1), rng.standard_normal(1), rng.standard_normal(D)) for _ in range(L)]
    x = rng.standard_normal((D, N))
    ybar, ljbar = rng.standard_normal((D, N)), rng.standard_normal(N)

    def loss(ps, xx):
        z, lj = xx, np.zeros(N)
        for (a, b, z0) in ps:
            z, l1 = O.radial_forward(a, b, z0, z)
            lj = lj + l1
        return float(np.sum(z * ybar) + np.sum(lj * ljbar))

    xbar, grads = O.radial_chain_vjp(params, x, ybar, ljbar)
    eps = 1e-6
    fd = np.zeros_like(x)
    for i in range(D):
        for n in range(N):
            xp, xm = x.copy(), x.copy()
            xp[i, n] += eps
            xm[i, n] -= eps
            fd[i, n] = (loss(params, xp) - loss(params, xm)) / (2 * eps)
    assert np.allclose(xbar, fd, rtol=2e-6, atol=1e-7)
    for l in range(L):
        for k, name in enumerate(("alpha_raw", "beta", "z0")):
            base = np.asarray(params[l][k], dtype=np.float64)
            fdp = np.zeros_like(base)
            for i in range(base.size):
                pp, pm = base.copy(), base.copy()
                pp[i] += eps
                pm[i] -= eps
                ps_p = [tuple(pp if (ll == l and kk == k) else params[ll][kk] for kk in range(3)) for ll in range(L)]
                ps_m = [tuple(pm if (ll == l and kk == k) else params[ll][kk] for kk in range(3)) for ll in range(L)]
                fdp[i] = (loss(ps_p, x) - loss(ps_m, x)) / (2 * eps)
            got = np.atleast_1d(np.asarray(grads[l][k], dtype=np.float64))
            assert np.allclose(got, fdp, rtol=2e-5, atol=2e-7), (l, name, got, fdp)


def test_forward_and_inverse_vjp_are_mutually_inverse_maps():
    """J(inverse(f)) at y = J(f)⁻¹ at x: pulling a cotangent back through inverse(f) and then through f returns it."""
    rng = np.random.default_rng(41)
    D, N, L = 7, 6, 3
    params = [(rng.standard_normal(D) / np.sqrt(D), rng.standard_normal(D) / np.sqrt(D), rng.standard_normal(1)) for _ in range(L)]
    x = rng.standard_normal((D, N))
    y = x
    for (w, u, b) in params:
        y, _ = O.planar_forward(w, u, b, y)
    v = rng.standard_normal((D, N))
    u_, _ = O.planar_inverse_chain_vjp(params, y, v, np.zeros(N))     # J_inv(y)ᵀ v
    back, _ = O.planar_chain_vjp(params, x, u_, np.zeros(N))          # J_f(x)ᵀ J_inv(y)ᵀ v = v
    assert np.allclose(back, v, rtol=1e-9, atol=1e-10)
    # and the two logjac cotangent paths are consistent: d/dy [logjac_inv(y)] = −J_inv(y)ᵀ d/dx [logjac_f(x)]
    ones = np.ones(N)
    gx, _ = O.planar_chain_vjp(params, x, np.zeros((D, N)), ones)
    gy, _ = O.planar_inverse_chain_vjp(params, y, np.zeros((D, N)), ones)
    pulled, _ = O.planar_inverse_chain_vjp(params, y, gx, np.zeros(N))
    assert np.allclose(gy, -pulled, rtol=1e-8, atol=1e-9)


def test_logit_and_truncated_closed_forms():
    """Logit / TruncatedBijector restatements against closed forms and the reference's own identities:
    Logit(0, 1)(0.5) = 0 with logjac −log(0.25) (logit.jl:15-24); a both-sided TruncatedBijector equals Logit inside its
    support (truncated.jl:21-23 vs logit.jl:15); one-sided bounds are log links (:24-27); the inverse log-Jacobian's
    closed form (:68-76) equals minus the forward one at the recovered point; logjac equals log|dy/dx|."""
    y, l = O.logit_forward(0.0, 1.0, np.array([0.5]))
    assert y[0] == 0.0 and abs(l[0] + np.log(0.25)) < 1e-15
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.9, 2.9, 200)
    yl, ll = O.logit_forward(-1.0, 3.0, x)
    yt, lt = O.truncated_forward(-1.0, 3.0, x)
    assert np.array_equal(yl, yt) and np.array_equal(ll, lt)
    y1, l1 = O.truncated_forward(-1.0, np.inf, x)
    assert np.allclose(y1, np.log(x + 1.0)) and np.allclose(l1, -np.log(x + 1.0))
    y2, l2 = O.truncated_forward(-np.inf, 3.0, x)
    assert np.allclose(y2, np.log(3.0 - x)) and np.allclose(l2, -np.log(3.0 - x))
    for lb, ub in ((-1.0, 3.0), (-1.0, np.inf), (-np.inf, 3.0), (-np.inf, np.inf)):
        yy, lf = O.truncated_forward(lb, ub, x)
        xi, li = O.truncated_inverse(lb, ub, yy)
        assert np.allclose(xi, x, atol=1e-12) and np.allclose(li, -lf, atol=1e-12)
        h = 1e-6
        num = np.log(np.abs((O.truncated_forward(lb, ub, x + h)[0] - O.truncated_forward(lb, ub, x - h)[0]) / (2 * h)))
        assert np.allclose(num, lf, atol=1e-6)
    # out-of-support inputs are clamped to the bound first (Bijectors.jl:95-100): the link then diverges like the reference
    with np.errstate(divide="ignore"):
        yc, _ = O.truncated_forward(-1.0, np.inf, np.array([-2.0]))
    assert np.isneginf(yc[0])


def _fd_vjp_check(fwd, params, x, ybar, ljbar, analytic, h=1e-6, tol=2e-6):
    """<cotangents, outputs(θ + h e)> differenced against the analytic VJP for every input / parameter entry sampled."""
    def scalar(ps, xx):
        y, lj = fwd(ps, xx)
        return float((ybar * y).sum() + (ljbar * lj).sum())

    rng = np.random.default_rng(0)
    xbar, pbars = analytic
    for _ in range(25):
        i, n = rng.integers(x.shape[0]), rng.integers(x.shape[1])
        xp, xm = x.copy(), x.copy()
        xp[i, n] += h
        xm[i, n] -= h
        fd = (scalar(params, xp) - scalar(params, xm)) / (2 * h)
        assert abs(fd - xbar[i, n]) <= tol * max(1.0, abs(fd)), ("x", i, n, fd, xbar[i, n])
    for k, pbar in enumerate(pbars):
        flat = params[k].reshape(-1)
        for _ in range(25):
            j = rng.integers(flat.size)
            pp, pm = [q.copy() for q in params], [q.copy() for q in params]
            pp[k].reshape(-1)[j] += h
            pm[k].reshape(-1)[j] -= h
            fd = (scalar(pp, x) - scalar(pm, x)) / (2 * h)
            assert abs(fd - pbar.reshape(-1)[j]) <= tol * max(1.0, abs(fd)), ("param", k, j, fd, pbar.reshape(-1)[j])


@pytest.mark.parametrize("inverse", [False, True])
def test_coupling_and_batchnorm_vjp_oracles_match_finite_differences(inverse):
    """The reverse-mode restatements that pin the device VJP kernels (coupling incl. the combine pullback,
    ext/BijectorsChainRulesCoreExt.jl:48-62; eval-mode BatchNorm) against central finite differences of the pinned
    forward / inverse oracles."""
    rng = np.random.default_rng(11)
    D, N, n1 = 10, 7, 4
    idx1, idx2 = [2, 5, 7, 9], [1, 3, 4, 8, 10]  # row 6 passes through
    W, c = rng.standard_normal((2 * n1, len(idx2))) * 0.3, rng.standard_normal(2 * n1) * 0.2
    x, ybar, ljbar = rng.standard_normal((D, N)), rng.standard_normal((D, N)), rng.standard_normal(N)
    f = (lambda ps, xx: O.coupling_affine_inverse(idx1, idx2, ps[0], ps[1], xx)) if inverse else (lambda ps, xx: O.coupling_affine_forward(idx1, idx2, ps[0], ps[1], xx))
    xbar, Wbar, cbar = O.coupling_affine_vjp(idx1, idx2, W, c, x, ybar, ljbar, inverse=inverse)
    _fd_vjp_check(f, [W, c], x, ybar, ljbar, (xbar, [Wbar, cbar]))
    b, logs, m, v = rng.standard_normal(D) * 0.3, rng.standard_normal(D) * 0.3, rng.standard_normal(D) * 0.3, rng.uniform(0.5, 1.5, D)

    def bn_of(ps):
        return O.BatchNormParams(ps[0], ps[1], m, v, np.float64(1e-5), np.float64(0.1))

    g = (lambda ps, xx: O.batchnorm_inverse(bn_of(ps), xx)) if inverse else (lambda ps, xx: O.batchnorm_forward(bn_of(ps), xx))
    xb, bb, lb = O.batchnorm_eval_vjp(bn_of([b, logs]), x, ybar, ljbar, inverse=inverse)
    _fd_vjp_check(g, [b, logs], x, ybar, ljbar, (xb, [bb, lb]))


@pytest.mark.parametrize("flags", [(False, False, False), (True, True, True), (True, False, True)])
def test_radial_vjp_oracle_both_directions_matches_finite_differences(flags):
    """radial_chain_vjp_dir (inverse layers differentiate compute_r implicitly) against central finite differences of the
    pinned forward / inverse oracles; the all-forward case reproduces radial_chain_vjp."""
    rng = np.random.default_rng(21)
    D, N = 6, 5
    params = [[rng.standard_normal(1), rng.standard_normal(1), rng.standard_normal(D)] for _ in flags]
    x, ybar, ljbar = rng.standard_normal((D, N)), rng.standard_normal((D, N)), rng.standard_normal(N)

    def fwd(ps, xx):
        lj = np.zeros(N)
        for k, inv in enumerate(flags):
            a, b, z0 = ps[3 * k], ps[3 * k + 1], ps[3 * k + 2]
            xx, l = (O.radial_inverse(a, b, z0, xx) if inv else O.radial_forward(a, b, z0, xx))
            lj = lj + l
        return xx, lj

    xbar, grads = O.radial_chain_vjp_dir([tuple(p) for p in params], flags, x, ybar, ljbar)
    flat = [q for p in params for q in p]
    pbars = [np.asarray(g).reshape(np.asarray(q).shape) for gs in grads for g, q in zip(gs, flat[:3])]
    _fd_vjp_check(fwd, flat, x, ybar, ljbar, (xbar, pbars), h=1e-6, tol=5e-6)
    if not any(flags):
        xb0, g0 = O.radial_chain_vjp([tuple(p) for p in params], x, ybar, ljbar)
        assert np.allclose(xb0, xbar) and all(np.allclose(a, b) for ga, gb in zip(g0, grads) for a, b in zip(ga, gb))


@pytest.mark.parametrize("inverse", [False, True])
def test_rqs_vjp_oracle_matches_finite_differences(inverse):
    """rqs_vjp (input and processed-knot cotangents, both directions) against central finite differences of the pinned
    spline oracles."""
    rng = np.random.default_rng(5)
    D, K, N = 4, 6, 9
    W, H, Dv = O.rqs_params(rng.standard_normal((D, K)), rng.standard_normal((D, K)), rng.standard_normal((D, K - 1)), 3.0)
    x = rng.standard_normal((D, N)) * 1.6
    ybar, ljbar = rng.standard_normal((D, N)), rng.standard_normal(N)
    xin = O.rqs_forward(W, H, Dv, x)[0] if inverse else x
    f = (lambda ps, xx: O.rqs_inverse(ps[0], ps[1], ps[2], xx)) if inverse else (lambda ps, xx: O.rqs_forward(ps[0], ps[1], ps[2], xx))
    xb, Wb, Hb, Db = O.rqs_vjp(W, H, Dv, xin, ybar, ljbar, inverse=inverse)
    # the last knot also selects the identity branch (a jump): differentiate interior knots only
    mask = np.ones_like(W, bool)
    mask[:, -1] = False

    class Interior:
        pass

    def fd_params():
        h = 1e-6
        for T, Tb, idx in ((W, Wb, 0), (H, Hb, 1), (Dv, Db, 2)):
            for i in range(D):
                for k in range(W.shape[1] - 1):
                    ps_p, ps_m = [W.copy(), H.copy(), Dv.copy()], [W.copy(), H.copy(), Dv.copy()]
                    ps_p[idx][i, k] += h
                    ps_m[idx][i, k] -= h
                    sc = lambda ps: float((ybar * f(ps, xin)[0]).sum() + (ljbar * f(ps, xin)[1]).sum())  # noqa: E731
                    fd = (sc(ps_p) - sc(ps_m)) / (2 * h)
                    assert abs(fd - Tb[i, k]) <= 5e-6 * max(1.0, abs(fd)), (idx, i, k, fd, Tb[i, k])

    fd_params()
    _fd_vjp_check(f, [W, H, Dv], xin, ybar, ljbar, (xb, []), h=1e-6, tol=5e-6)
