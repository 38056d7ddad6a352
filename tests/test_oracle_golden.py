"""Pins the numpy oracle against every deterministic golden vector the reference's own tests hold
(tests/golden/reference_vectors.json).  CPU only."""
import math

import numpy as np
import pytest

from oracle import oracle_np as O


def test_coupling_shift_golden(golden):
    g = golden["coupling_shift"]  # test/bijectors/coupling.jl:18-42
    m = O.PartitionMask.make(g["mask"]["n"], g["mask"]["indices_1"], g["mask"]["indices_2"])
    theta = lambda x2: O.Shift(x2[0])  # noqa: E731   Coupling(x -> Shift(x[1]), m)
    x = np.array(g["x"])
    y, lj = O.coupling_forward(theta, m, x)
    assert np.array_equal(y, np.array(g["y"]))
    assert lj == g["logjac"]
    xi, lji = O.coupling_inverse(theta, m, y)
    assert np.array_equal(xi, x)  # icl1(cl1(x)) == x
    assert lji == -lj


def test_coupling_scale_golden(golden):
    g = golden["coupling_scale"]  # test/bijectors/coupling.jl:44-56
    m = O.PartitionMask.make(g["mask"]["n"], g["mask"]["indices_1"], g["mask"]["indices_2"])
    theta = lambda x2: O.Scale(x2[0])  # noqa: E731
    for case in g["cases"]:
        x = np.array(case["x"])
        y, lj = O.coupling_forward(theta, m, x)
        np.testing.assert_allclose(y, np.array(case["y"]), rtol=0, atol=0)
        assert lj == pytest.approx(math.log(2), rel=1e-15)
        xi, lji = O.coupling_inverse(theta, m, y)
        np.testing.assert_allclose(xi, x, rtol=1e-15)
        assert lji == pytest.approx(-math.log(2), rel=1e-15)


def test_partition_mask_golden(golden):
    g = golden["partition_mask"]  # test/bijectors/coupling.jl:4-16
    m1 = O.PartitionMask.make(g["n"], g["indices_1"], g["indices_2"])
    m2 = O.PartitionMask.make(g["n"], g["indices_1"], g["indices_2"], g["indices_3_inferred"])
    assert np.array_equal(m1.indices_3, m2.indices_3)
    x = np.array(g["x"])
    x1, x2, x3 = O.partition(m1, x)
    assert np.array_equal(x1, g["x1"]) and np.array_equal(x2, g["x2"]) and np.array_equal(x3, g["x3"])
    assert np.array_equal(O.combine(m1, x1, x2, x3), x)
    # default split Coupling(θ, n): first n÷2 indices transformed (src/bijectors/coupling.jl:183-186)
    m = O.PartitionMask.make(6, range(1, 6 // 2 + 1))
    assert m.indices_1.tolist() == [1, 2, 3] and m.indices_2.tolist() == [4, 5, 6] and len(m.indices_3) == 0


@pytest.mark.parametrize("key", ["permute_2", "permute_3"])
def test_permute_golden(golden, key):
    g = golden[key]  # test/bijectors/permute.jl:13-64
    n = len(g["indices"])
    A1 = np.array(g["matrix"], dtype=float)
    A2 = O.permute_matrix_from_indices(g["indices"])
    A3 = O.permute_matrix_from_pairs(n, *[tuple(p) for p in g["pairs"]])
    A4 = O.permute_matrix_from_pairs(n, *[(p[0], p[1]) for p in g["vector_pairs"]])
    assert np.array_equal(A1, A2) and np.array_equal(A2, A3) and np.array_equal(A3, A4)
    x = np.array(g["x"])
    y, lj = O.permute_forward(A2, x)
    assert np.array_equal(y, g["y"]) and lj == 0.0
    xi, _ = O.permute_inverse(A2, y)
    assert np.array_equal(xi, x)


def test_permute_invalid_golden(golden):
    for case in golden["permute_invalid"]["cases"]:  # test/bijectors/permute.jl:9-10
        with pytest.raises(ValueError):
            if "pairs" in case:
                O.permute_matrix_from_pairs(case["n"], *[tuple(p) for p in case["pairs"]])
            else:
                O.permute_matrix_from_pairs(case["n"], *[(p[0], p[1]) for p in case["vector_pairs"]])


def test_permute_bit_exact_payloads():
    # index movement must preserve NaN payloads and -0.0
    x = np.array([np.nan, -0.0, 1.5, np.inf], dtype=np.float32)
    A = O.permute_matrix_from_indices([3, 1, 4, 2])
    y, _ = O.permute_forward(A, x)
    assert y.view(np.uint32).tolist() == x.view(np.uint32)[[1, 3, 0, 2]].tolist()


def test_find_alpha_issue_204(golden):
    g = golden["find_alpha_issue_204"]  # test/normalising_flows.jl:65-70
    a = O.find_alpha(np.float64(g["wt_y"]), np.float64(g["wt_u_hat"]), np.float64(g["b"]))
    expected = g["wt_y"] + g["wt_u_hat"]
    assert a == pytest.approx(expected, rel=math.sqrt(np.finfo(np.float64).eps))


def test_find_alpha_residual_grid(golden):
    g = golden["find_alpha_grid"]  # test/normalising_flows.jl:47-63
    rtol = math.sqrt(np.finfo(np.float64).eps)  # Julia isapprox default
    for wt_y in g["wt_y"]:
        for c in g["wt_u_hat"]:
            for b in g["b"]:
                a = float(O.find_alpha(np.float64(wt_y), np.float64(c), np.float64(b)))
                rhs = a + c * math.tanh(a + b)
                atol = 1e-14 if wt_y == 0 else 0.0
                assert abs(wt_y - rhs) <= max(atol, rtol * max(abs(wt_y), abs(rhs))), (wt_y, c, b, a)


def test_planar_deterministic(golden):
    # test/normalising_flows.jl:37-42 with the analytic values of SURVEY §8(c)(6)
    w, u, b = np.ones(10), np.zeros(10), 1.0
    z = np.ones((10, 100))
    u_hat, wTu_hat = O.get_u_hat(u, w)
    np.testing.assert_allclose(u_hat, (math.log(2) - 1) / 10, rtol=1e-15)
    assert wTu_hat == pytest.approx(math.log(2) - 1, rel=1e-15)
    y, lj = O.planar_forward(w, u, b, z)
    np.testing.assert_allclose(y, 1 + (math.log(2) - 1) / 10 * math.tanh(11.0), rtol=1e-15)
    np.testing.assert_allclose(lj, math.log1p((math.log(2) - 1) / math.cosh(11.0) ** 2), rtol=1e-12, atol=1e-20)
    zi, lji = O.planar_inverse(w, u, b, y)
    np.testing.assert_allclose(zi, z, rtol=1e-12)  # inverse(flow)(flow(z)) ≈ z
    np.testing.assert_allclose(lji, -lj, rtol=1e-9, atol=1e-18)


def test_radial_deterministic(golden):
    # test/normalising_flows.jl:86-91: α_=1, β=1 ⇒ β̂ = 0 ⇒ identity, logjac 0
    z0, z = np.zeros(10), np.ones((10, 100))
    y, lj = O.radial_forward(1.0, 1.0, z0, z)
    np.testing.assert_allclose(y, z, rtol=0, atol=1e-15)
    np.testing.assert_allclose(lj, 0.0, atol=1e-14)
    zi, _ = O.radial_inverse(1.0, 1.0, z0, y)
    np.testing.assert_allclose(zi, z, rtol=1e-13)


def test_batchnorm_default(golden):
    # test/normalising_flows.jl:7-23, defaults src/bijectors/normalise.jl:26-37
    bn = O.BatchNormParams.default(2)
    x = np.random.default_rng(1).standard_normal((2, 20)).astype(np.float32)
    y, lj = O.batchnorm_forward(bn, x)
    np.testing.assert_allclose(y, x / np.sqrt(np.float32(1) + np.float32(1e-5)), rtol=1e-6)
    np.testing.assert_allclose(lj, -np.log(np.float32(1) + np.float32(1e-5)), rtol=1e-3)
    assert lj.shape == (20,)
    xi, lji = O.batchnorm_inverse(bn, y)
    np.testing.assert_allclose(xi, x, rtol=1e-6)
    np.testing.assert_allclose(lji, -lj)
    with pytest.raises(ValueError):  # channel mismatch errors (normalise.jl:43-45)
        O.batchnorm_forward(bn, np.zeros((10, 2), np.float32))


def test_rqs_outside_and_ctor(golden):
    g = golden["rqs_outside"]  # test/bijectors/rational_quadratic_spline.jl:47-61
    rng = np.random.default_rng(7)
    B, K, d = g["B"], g["K"], g["d"]
    W, H, Dv = O.rqs_params(rng.standard_normal(K), rng.standard_normal(K), rng.standard_normal(K - 1), B)
    for c in g["univariate"]:
        y, lj = O.rqs_forward_scalar(W, H, Dv, c["x"])
        assert y == c["y"] and lj == c["logjac"]
        assert O.rqs_inverse_scalar(W, H, Dv, c["y"]) == c["x"]
    Wm, Hm, Dm = O.rqs_params(rng.standard_normal((d, K)), rng.standard_normal((d, K)), rng.standard_normal((d, K - 1)), B)
    x = np.array(g["multivariate"]["x"])
    y, lj = O.rqs_forward(Wm, Hm, Dm, x)
    assert np.array_equal(y, x) and lj == 0.0
    # ctor invariants, :14-36
    for Wk, Hk, Dk in ((W[None], H[None], Dv[None]), (Wm, Hm, Dm)):
        np.testing.assert_allclose(Wk[:, 0], -B, rtol=1e-15)
        np.testing.assert_allclose(Wk[:, -1], B, rtol=1e-14)
        assert np.all(np.diff(Wk, axis=1) > 0) and np.all(np.diff(Hk, axis=1) > 0)
        assert np.all(Dk > 0) and np.all(Dk[:, 0] == 1) and np.all(Dk[:, -1] == 1)
        O.rqs_validate(Wk, Hk, Dk)
    # normalisation closed form, :78-105
    ws = rng.standard_normal((d, K))
    Wt = np.concatenate([np.zeros((d, 1)), O.softmax_rows(ws)], axis=1)
    Wp, _, _ = O.rqs_params(ws, ws, rng.standard_normal((d, K - 1)), B)
    np.testing.assert_allclose((2 * B) * (np.cumsum(Wt, axis=1) - 0.5), Wp, rtol=1e-12, atol=1e-15)
    # Float32 construction, :64-76
    W32, _, D32 = O.rqs_params(ws.astype(np.float32), ws.astype(np.float32), rng.standard_normal((d, K - 1)).astype(np.float32), B)
    assert W32.dtype == np.float32 and D32.dtype == np.float32
    with pytest.raises(AssertionError):
        O.rqs_validate(Wm, Hm, -Dm)


def test_rqs_scalar_matches_batched():
    rng = np.random.default_rng(11)
    D, K, B = 5, 8, 3.0
    W, H, Dv = O.rqs_params(rng.standard_normal((D, K)), rng.standard_normal((D, K)), rng.standard_normal((D, K - 1)), B)
    X = rng.standard_normal((D, 40)) * 2.0
    X[0, 0] = W[0, 3]  # exactly on an interior knot -> bin on its left
    Y, lj = O.rqs_forward(W, H, Dv, X)
    for n in range(X.shape[1]):
        tot = 0.0
        for i in range(D):
            y, l = O.rqs_forward_scalar(W[i], H[i], Dv[i], X[i, n])
            assert y == pytest.approx(Y[i, n], rel=1e-14, abs=1e-15)
            tot += l
            assert O.rqs_inverse_scalar(W[i], H[i], Dv[i], y) == pytest.approx(X[i, n], rel=1e-9, abs=1e-12)
        assert tot == pytest.approx(lj[n], rel=1e-12, abs=1e-14)


def test_elementwise_exp_doctest(golden):
    g = golden["elementwise_exp_doctest"]  # src/interface.jl:21-31
    y, lj = O.elementwise_exp(np.array(g["x"]))
    assert y.tolist() == g["y"] and lj == g["logjac"]
    # BASELINE config 1: Float64 vector of length 1024
    x = np.random.default_rng(1).standard_normal(1024)
    y, lj = O.elementwise_exp(x)
    assert y.dtype == np.float64 and lj == pytest.approx(x.sum(), rel=1e-12)
    xl, ljl = O.elementwise_log(y)
    np.testing.assert_allclose(xl, x, rtol=1e-12, atol=1e-14)
    assert ljl == pytest.approx(-lj, rel=1e-10)


def test_stacked_golden(golden):
    g = golden["stacked_value_test"]  # test/bijectors/stacked.jl:100-108
    ops = [(O.EW.EXP, 0.0), (O.EW.LOG, 0.0), (O.EW.SHIFT, 5.0)]
    ranges = [tuple(r) for r in g["ranges"]]
    x = np.array(g["x"])
    y, lj = O.stacked_forward(ops, ranges, x)
    assert y.tolist() == [math.exp(1.0), 0.0, 6.0] and lj == 1.0
    xi, lji = O.stacked_inverse(ops, ranges, y)
    np.testing.assert_allclose(xi, x, rtol=1e-15)
    assert lji == pytest.approx(-1.0)
    g2 = golden["stacked_single_block"]  # :232-238
    y, lj = O.stacked_forward([(O.EW.IDENTITY, 0.0)], [(1, 2)], np.array(g2["x"]))
    assert y.tolist() == g2["y"] and lj == 0
    with pytest.raises(ValueError):  # :120-121 input length mismatch
        O.stacked_forward(ops, ranges, np.ones(4))


def test_philox4x32_10_known_answer_vectors():
    """The device-side generator of rand(td, n) restated in the oracle: the three known-answer vectors of Random123
    (kat_vectors, philox4x32 10 rounds)."""
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kats:
        out = O.philox4x32_10(*[np.array([c], np.uint32) for c in ctr], key[0], key[1])
        assert tuple(int(w[0]) for w in out) == exp


def test_philox_normals_are_standard_normal_and_shard_consistently():
    z = O.philox_normals(1234, 0, 10, 20000)
    assert z.shape == (10, 20000) and abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs(np.mean(z ** 3)) < 0.03 and abs(np.mean(z ** 4) - 3) < 0.1
    # a column shard continues the same stream
    assert np.array_equal(O.philox_normals(1234, 0, 10, 500, column_offset=700), z[:, 700:1200])
    assert not np.array_equal(O.philox_normals(1235, 0, 10, 500), z[:, :500])
    zz = O.philox_normals(7, 3, 6, 100, mu=np.arange(6.0), sigma=np.full(6, 2.0))
    assert np.allclose(zz, 2 * O.philox_normals(7, 3, 6, 100) + np.arange(6.0)[:, None])
