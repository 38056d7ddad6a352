"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol include/b2b.h declares
(no compute calls without a GPU), chain flattening / inversion order, constructors and error behaviour
mirror the reference, and the world_size-2 gloo path of the sharded log-density sum."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

import bijectors_jl_b200 as B
from bijectors_jl_b200 import _lib
from oracle import oracle_np as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b2b.h")).read()
    declared = set(re.findall(r"\b(b2b_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(B.exported_symbols()), declared ^ set(B.exported_symbols())
    handle = ctypes.CDLL(B.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert B.lib().b2b_version() == 100
    assert ctypes.sizeof(_lib.LayerDesc) == 80  # layout of b2b_layer_desc
    for code in (0, -1, -2, -3, -4):
        assert B.lib().b2b_status_string(code)


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    flow = B.PlanarLayer(np.ones(4, f32), np.zeros(4, f32), f32(0.0), device="cpu")
    with pytest.raises(B.B2BError, match="no CPU fallback"):
        B.with_logabsdet_jacobian(flow, torch.zeros(4))


def kinds(descs):
    return [(d.kind, d.inverse) for d in descs]


def test_chain_flattening_and_inverse_order():
    D = 4
    p1 = B.PlanarLayer(np.ones(D, f32), np.zeros(D, f32), f32(0.0), device="cpu")
    r1 = B.RadialLayer(f32(0.1), f32(0.2), np.zeros(D, f32), device="cpu")
    bn = B.InvertibleBatchNorm(D, device="cpu")
    pm = B.Permute([2, 1, 4, 3], device="cpu")
    flow = pm @ bn @ r1 @ p1  # pm ∘ bn ∘ r1 ∘ p1: p1 applied first
    assert kinds(flow._descs(False, D)) == [(_lib.PLANAR, 0), (_lib.RADIAL, 0), (_lib.BATCHNORM, 0), (_lib.PERMUTE, 0)]
    assert [type(b) for b in B.flatten(flow)] == [B.PlanarLayer, B.RadialLayer, B.InvertibleBatchNorm, B.Permute]
    # inverse(f∘g) = inverse(g)∘inverse(f): last-applied layer inverted first (InverseFunctions)
    inv = B.inverse(flow)
    assert kinds(inv._descs(False, D)) == [(_lib.PERMUTE, 1), (_lib.BATCHNORM, 1), (_lib.RADIAL, 1), (_lib.PLANAR, 1)]
    assert kinds(B.inverse(inv)._descs(False, D)) == kinds(flow._descs(False, D))
    flat = B.Composed(p1, r1, B.Composed(bn, pm))
    assert kinds(flat._descs(False, D)) == kinds(flow._descs(False, D))
    assert kinds(B.inverse(flat)._descs(False, D)) == kinds(inv._descs(False, D))
    assert kinds(B.compose(pm, bn, r1, p1)._descs(False, D)) == kinds(flow._descs(False, D))
    assert B.inverse(B.inverse(p1)) is p1 and B.inverse(p1) == B.Inverse(p1)
    assert not B.isclosedform(inv) and B.isclosedform(B.inverse(r1))
    with pytest.raises(ValueError, match="DimensionMismatch"):
        flow._descs(False, D + 1)
    too_long = B.Composed(*([p1] * 25))
    with pytest.raises(B.B2BError):
        from bijectors_jl_b200.interface import _desc_array

        _desc_array(too_long._descs(False, D))


def test_layout_helpers():
    a = np.arange(12, dtype=f32).reshape(3, 4)
    t = B.from_numpy(a, device="cpu")
    assert t.shape == (3, 4) and t.stride() == (1, 3)
    assert np.array_equal(B.to_numpy(t), a)
    from bijectors_jl_b200.interface import _batch_view

    assert _batch_view(t) == (3, 4, 3)
    assert _batch_view(torch.zeros(5)) == (5, 1, 5)
    with pytest.raises(ValueError, match="column-major"):
        _batch_view(torch.zeros(3, 4))
    assert _batch_view(torch.zeros(3, dtype=torch.float64)) == (3, 1, 3)  # Float64 batches: b2b_chain_run_f64
    with pytest.raises(TypeError):
        _batch_view(torch.zeros(3, dtype=torch.float16))
    # parameter and batch element types must agree (Float32 hot path / Float64 reference-test precision)
    lay64 = B.PlanarLayer(np.ones(3), np.zeros(3), np.ones(1), device="cpu", dtype=torch.float64)
    assert lay64.w.dtype == torch.float64 and type(lay64._descs(False, 3, torch.float64)[0]).__name__ == "LayerDesc64"
    with pytest.raises(TypeError, match="Float"):
        lay64._descs(False, 3, torch.float32)
    e = B.colmajor_empty(7, 5, device="cpu")
    assert e.shape == (7, 5) and e.stride() == (1, 7)


def test_constructors_and_error_behaviour(golden):
    # PartitionMask (test/bijectors/coupling.jl:4-16) and default split (coupling.jl:183-186)
    m = B.PartitionMask(3, [1], [2])
    assert m == B.PartitionMask(3, [1], [2], [3])
    m6 = B.PartitionMask(6, range(1, 4))
    assert (m6.indices_1, m6.indices_2, m6.indices_3) == ([1, 2, 3], [4, 5, 6], [])
    with pytest.raises(ValueError):
        B.PartitionMask(3, [1], [1])
    # Coupling: arbitrary θ cannot run on the device path (no CPU fallback)
    with pytest.raises(B.B2BError):
        B.Coupling(lambda x2: None, m)
    cond = B.AffineConditioner(np.zeros((2, 1), f32), device="cpu")
    cl = B.Coupling(cond, m, device="cpu")
    d = cl._descs(False, 3)[0]
    assert (d.kind, d.n0, d.n1) == (_lib.COUPLING_AFFINE, 1, 1) and B.coupling(cl) is cond
    with pytest.raises(ValueError):
        B.Coupling(B.AffineConditioner(np.zeros((4, 1), f32), device="cpu"), m)
    # Permute spellings agree; invalid ones raise (test/bijectors/permute.jl:9-21)
    g = golden["permute_2"]
    bs = [B.Permute(np.array(g["matrix"]), device="cpu"), B.Permute(g["indices"], device="cpu"),
          B.Permute(2, *[tuple(p) for p in g["pairs"]], device="cpu"),
          B.Permute(2, *[(p[0], p[1]) for p in g["vector_pairs"]], device="cpu")]
    assert all(b == bs[0] for b in bs) and np.array_equal(bs[0].A, np.array(g["matrix"]))
    assert np.array_equal(B.Permute([2, 3, 1], device="cpu").A, O.permute_matrix_from_indices([2, 3, 1]))
    for case in golden["permute_invalid"]["cases"]:
        with pytest.raises(ValueError, match="ArgumentError"):
            if "pairs" in case:
                B.Permute(case["n"], *[tuple(p) for p in case["pairs"]], device="cpu")
            else:
                B.Permute(case["n"], *[(p[0], p[1]) for p in case["vector_pairs"]], device="cpu")
    # InvertibleBatchNorm channel mismatch: reference error text (normalise.jl:43-45); training mode unsupported
    bn = B.InvertibleBatchNorm(2, device="cpu")
    assert bn.eps == pytest.approx(1e-5) and bn.mtm == pytest.approx(0.1)
    assert set(bn.params()) == {"b", "logs", "m", "v"}
    with pytest.raises(RuntimeError, match="InvertibleBatchNorm expected 2 channels, got 10"):
        bn._descs(False, 10)
    with pytest.raises(B.B2BError):  # training-mode BN cannot be fused / inverted (normalise.jl:75)
        B.InvertibleBatchNorm(2, device="cpu", training=True)._descs(False, 2)
    # Stacked: ranges bookkeeping + length mismatch error text (stacked.jl:158-160)
    sb = B.Stacked([B.elementwise("exp"), B.elementwise("log"), B.Shift(5.0)], device="cpu")
    assert sb.ranges_in == [(1, 1), (2, 2), (3, 3)] and sb.length_in == 3
    with pytest.raises(RuntimeError, match=r"input length mismatch \(3 != 4\)"):
        sb._descs(False, 4)
    with pytest.raises(B.B2BError):
        B.Stacked([B.PlanarLayer(2, device="cpu")], [(1, 2)], device="cpu")
    assert B.inverse(B.elementwise("exp")) == B.elementwise("log") and B.inverse(B.Shift(2.0)) == B.Shift(-2.0)
    # RQS: the normalising constructor agrees with the oracle restatement; asserts fire
    rng = np.random.default_rng(0)
    rw, rh, rd = rng.standard_normal((5, 8)).astype(f32), rng.standard_normal((5, 8)).astype(f32), rng.standard_normal((5, 7)).astype(f32)
    lay = B.RationalQuadraticSpline(rw, rh, rd, 3.0, device="cpu")
    W, H, Dv = lay.knots()
    Wo, Ho, Do = O.rqs_params(rw.astype(np.float64), rh.astype(np.float64), rd.astype(np.float64), 3.0)
    assert np.allclose(W, Wo, atol=2e-6) and np.allclose(H, Ho, atol=2e-6) and np.allclose(Dv, Do, rtol=1e-6)
    assert lay.K1 == 9 and lay.widths.shape == (9, 5)  # device table is knot-major
    with pytest.raises(AssertionError, match="positive"):
        B.RationalQuadraticSpline(W, H, -Dv, device="cpu")
    uv = B.RationalQuadraticSpline(rw[0], rh[0], rd[0], 2, device="cpu")
    assert uv.D == 1 and uv.K1 == 9


def test_config1_elementwise_exp_plumbing(golden):
    """BASELINE configs[0]: Exp bijector with_logabsdet_jacobian on a Float64 vector of length 1024 --
    CPU plumbing, runs without a GPU (src/interface.jl:21-31, exp_log.jl:6)."""
    g = golden["elementwise_exp_doctest"]
    y, lj = B.with_logabsdet_jacobian(B.elementwise("exp"), torch.tensor(g["x"], dtype=torch.float64))
    assert y.tolist() == g["y"] and float(lj) == g["logjac"]
    x = np.random.default_rng(1).standard_normal(1024)
    y, lj = B.with_logabsdet_jacobian(B.elementwise(math.exp), torch.from_numpy(x))
    yo, ljo = O.elementwise_exp(x)
    assert y.dtype == torch.float64 and np.allclose(y.numpy(), yo, rtol=1e-15) and float(lj) == pytest.approx(ljo, rel=1e-12)
    xb, ljb = B.with_logabsdet_jacobian(B.inverse(B.elementwise("exp")), y)
    assert np.allclose(xb.numpy(), x, rtol=1e-12, atol=1e-14) and float(ljb) == pytest.approx(-ljo, rel=1e-10)
    assert float(B.logabsdetjac(B.elementwise("exp"), torch.from_numpy(x))) == pytest.approx(x.sum(), rel=1e-12)


def test_shard_columns():
    from bijectors_jl_b200.distributed import shard_columns

    for N, W in [(1 << 20, 8), (1000, 3), (7, 8), (0, 2)]:
        blocks = [shard_columns(N, r, W) for r in range(W)]
        assert blocks[0][0] == 0 and blocks[-1][1] == N
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(W - 1))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist

    from bijectors_jl_b200.distributed import Communicator, shard_columns

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        D, N = 6, 1001
        w, u, b = rng.standard_normal(D), rng.standard_normal(D), rng.standard_normal(1)
        y = rng.standard_normal((D, N))
        layers = [O.Layer("planar", dict(w=w, u=u, b=b))]
        lo, hi = shard_columns(N, rank, world)
        # each rank evaluates ITS column shard (with the oracle here: there is no GPU in this test) ...
        local = torch.tensor(O.transformed_logpdf(layers, None, None, y[:, lo:hi]).sum(), dtype=torch.float64)
        comm = Communicator()
        assert comm.handle is None and comm.world == world
        total = comm.allreduce_sum_(local.reshape(1))  # ... and ONE scalar is summed across ranks
        full = O.transformed_logpdf(layers, None, None, y).sum()
        assert abs(float(total) - full) <= 1e-9 * abs(full)
        # the training path's exchange: parameter cotangents of column shards sum to the full-batch cotangents
        from bijectors_jl_b200.distributed import pack_param_grads, unpack_param_grads

        ybar, ljbar = rng.standard_normal((D, N)), rng.standard_normal(N)
        ps = [(w, u, b), (u * 0.5, w * 0.7, b + 0.1)]
        _, g_loc = O.planar_chain_vjp(ps, y[:, lo:hi], ybar[:, lo:hi], ljbar[lo:hi])
        as_dict = [{"w": torch.tensor(gw), "u": torch.tensor(gu), "b": torch.tensor([gb])} for gw, gu, gb in g_loc]
        buf = pack_param_grads(as_dict)
        assert buf.dtype == torch.float64 and buf.numel() == 2 * (2 * D + 1)
        comm.allreduce_sum_(buf)
        g_sum = unpack_param_grads(buf, as_dict)
        _, g_full = O.planar_chain_vjp(ps, y, ybar, ljbar)
        for l in range(2):
            assert np.allclose(g_sum[l]["w"].numpy(), g_full[l][0], rtol=1e-10, atol=1e-12)
            assert np.allclose(g_sum[l]["u"].numpy(), g_full[l][1], rtol=1e-10, atol=1e-12)
            assert np.allclose(g_sum[l]["b"].numpy(), g_full[l][2], rtol=1e-10, atol=1e-12)
        if rank == 0:
            out.put(float(total))
    finally:
        dist.destroy_process_group()


def test_sharded_logdensity_sum_gloo_world2():
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert math.isfinite(out.get(timeout=5))


def test_vjp_chains_are_cut_into_runs_of_one_direction():
    """planar_chain_vjp differentiates <= 8 layers of one direction per device call: the host side cuts longer / mixed
    chains into maximal such runs, in application order."""
    from types import SimpleNamespace as NS

    from bijectors_jl_b200.interface import _direction_runs

    def lens(flags):
        runs = _direction_runs([NS(inverse=f, tag=i) for i, f in enumerate(flags)])
        assert [d.tag for r in runs for d in r] == list(range(len(flags)))          # order kept, nothing dropped
        assert all(len({int(d.inverse) for d in r}) == 1 and 1 <= len(r) <= 8 for r in runs)
        return [len(r) for r in runs]

    assert lens([0] * 8) == [8] and lens([1] * 3) == [3]
    assert lens([0] * 9 + [1, 1, 0]) == [8, 1, 2, 1]
    assert lens([1, 0, 0, 1, 1]) == [1, 2, 2]
    assert lens([0] * 20) == [8, 8, 4]


def test_autograd_module_is_importable_without_a_gpu():
    """The torch.autograd glue builds parameters with the reference's shapes (planar_layer.jl:23-28); evaluating it
    needs the device path (no CPU fallback)."""
    flow = B.autograd.PlanarFlow(5, 3, device="cpu", generator=torch.Generator().manual_seed(0))
    assert [tuple(p.shape) for p in flow.w] == [(5,)] * 3 and [tuple(p.shape) for p in flow.b] == [(1,)] * 3
    assert len(list(flow.parameters())) == 9 and all(p.requires_grad for p in flow.parameters())
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            flow(torch.zeros(5, 4))


def test_vjp_workspace_queries_are_pure_host_functions():
    """Workspace sizes of the reverse-mode entry points can be queried without a device (b2b.h)."""
    L_ = B.lib()
    a = L_.b2b_planar_chain_vjp_workspace_bytes(8, 128, 1 << 20)
    b = L_.b2b_planar_chain_vjp_workspace_bytes(8, 128, 1 << 21)
    assert a >= (1 << 20) * 3 * 8 * 4 and b > a          # per-column scalars g | t | l̄·s/(1+c·s)
    assert L_.b2b_planar_chain_vjp_workspace_bytes(3, 64, 1000) == L_.b2b_planar_chain_vjp_workspace_bytes(4, 64, 1000)
    assert L_.b2b_planar_chain_vjp_workspace_bytes(9, 128, 10) == 0   # more than 8 layers: unsupported
    r = L_.b2b_radial_chain_vjp_workspace_bytes(6, 64)
    assert r >= 592 * (6 * 64 + 12) * 4 and L_.b2b_radial_chain_vjp_workspace_bytes(6, 200) == 0


# ---- the Julia binding is checked against the header without a Julia toolchain ----------------------------------------
def _c_prototypes():
    """name -> (return class, [argument classes]) of every function include/b2b.h declares."""
    import re

    hdr = open(os.path.join(ROOT, "include", "b2b.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)

    def cls(t):
        t = t.replace("const", " ").replace("struct", " ").strip()
        stars = t.count("*")
        base = t.replace("*", " ").split()
        base = base[0] if base else ""
        if stars == 0:
            return {"int": "i32", "int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "size_t": "usize", "float": "f32", "double": "f64"}[base]
        if stars == 2:
            return "ptrptr"
        return {"float": "ptr_f32", "double": "ptr_f64", "int32_t": "ptr_i32", "int": "ptr_i32", "void": "ptr_void", "char": "ptr_u8",
                "b2b_layer_desc": "ptr_desc", "b2b_layer_desc_f64": "ptr_desc64", "b2b_comm": "ptr_void",
                "b2b_host_ctx": "ptr_void"}[base]

    protos = {}
    for m in re.finditer(r"\b(int|size_t|const char\s*\*)\s+(b2b_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        out = []
        if args and args != "void":
            for a in args.split(","):
                a = re.sub(r"\[\d*\]", "*", a.strip())          # char id[128] -> char* id
                a = re.sub(r"\b[A-Za-z_][A-Za-z0-9_]*\s*$", "", a.replace("*", " * ")).strip() if not a.strip().endswith("*") else a
                out.append(cls(a))
        protos[name] = ({"int": "i32", "size_t": "usize"}.get(ret, "cstring"), out)
    return protos


def test_python_ctypes_signatures_match_the_header():
    """bijectors.jl_b200/_lib.py declares restype / argtypes for every entry point by hand: each must agree with the
    prototype in include/b2b.h in arity, integer widths and pointer-ness (a wrong width corrupts the call silently)."""
    import ctypes

    from bijectors_jl_b200 import _lib

    protos = _c_prototypes()
    assert set(_lib._SIGS) == set(protos), (set(_lib._SIGS) ^ set(protos))

    def width(c):  # (kind, bytes) of a ctypes type
        if isinstance(c, type) and (issubclass(c, ctypes._Pointer) or c in (ctypes.c_void_p, ctypes.c_char_p)):
            return ("p", 8)
        kind = {"i": "i", "l": "i", "q": "i", "I": "u", "L": "u", "Q": "u", "f": "f", "d": "f"}[c._type_]
        return (kind, ctypes.sizeof(c))

    def hclass(c):  # (kind, bytes) of a header class
        if c.startswith("ptr") or c == "cstring":
            return ("p", 8)
        return {"i32": ("i", 4), "i64": ("i", 8), "u64": ("u", 8), "usize": ("u", 8), "f32": ("f", 4), "f64": ("f", 8)}[c]

    for name, (ret, args) in protos.items():
        pret, pargs = _lib._SIGS[name]
        assert hclass(ret) == width(pret), (name, ret, pret)
        assert [hclass(a) for a in args] == [width(a) for a in pargs], (name, args, pargs)


def _julia_ccalls():
    """[(symbol, return class, [argument classes], number of actual arguments)] of every ccall in B200Bijectors.jl."""
    import re

    src = open(os.path.join(ROOT, "bijectors.jl_b200", "julia", "B200Bijectors.jl"), encoding="utf-8").read()
    src = "\n".join(l.split("#")[0] if not l.lstrip().startswith("#") else "" for l in src.splitlines())

    def balanced(s, i):  # s[i] == '(' -> index after the matching ')'
        depth = 0
        for j in range(i, len(s)):
            depth += s[j] == "("
            depth -= s[j] == ")"
            if depth == 0:
                return j + 1
        raise AssertionError("unbalanced parentheses")

    def split_top(s):
        parts, depth, cur = [], 0, ""
        for ch in s:
            depth += ch in "({["
            depth -= ch in ")}]"
            if ch == "," and depth == 0:
                parts.append(cur.strip())
                cur = ""
            else:
                cur += ch
        if cur.strip():
            parts.append(cur.strip())
        return parts

    jl = {"Cint": "i32", "Int32": "i32", "Int64": "i64", "UInt64": "u64", "Csize_t": "usize", "Cfloat": "f32", "Float32": "f32", "Cstring": "cstring",
          "CuPtr{Float32}": "ptr_f32", "Ptr{Float32}": "ptr_f32", "CuPtr{Float64}": "ptr_f64", "CuPtr{Int32}": "ptr_i32",
          "Ptr{Int32}": "ptr_i32", "Ptr{LayerDesc}": "ptr_desc", "Ptr{Cvoid}": "ptr_void", "CuPtr{Cvoid}": "ptr_void",
          "Ptr{UInt8}": "ptr_u8", "Ptr{Ptr{Cvoid}}": "ptrptr"}
    calls = []
    for m in re.finditer(r"ccall\(", src):
        end = balanced(src, m.end() - 1)
        parts = split_top(src[m.end():end - 1])
        sym = re.match(r"\(\s*:([a-z0-9_]+)\s*,\s*libb2b\s*\)", parts[0]).group(1)
        argt = split_top(parts[2].strip()[1:-1])
        calls.append((sym, jl[parts[1]], [jl[a] for a in argt if a], len(parts) - 3))
    return calls


def test_julia_binding_ccalls_match_the_header():
    """Every `ccall` of julia/B200Bijectors.jl names an exported symbol of include/b2b.h with the same arity, the same
    return class and the same argument classes (32/64-bit integer, size_t, float, typed pointer), and passes exactly as
    many values as it declares -- the check a Julia toolchain would do at load / first call."""
    protos = _c_prototypes()
    calls = _julia_ccalls()
    assert len(calls) >= 12 and set(protos) >= {"b2b_chain_run_f32", "b2b_comm_init_rank", "b2b_numa_bind_to_device"}
    seen = set()
    for sym, ret, args, nvals in calls:
        assert sym in protos, f"{sym} is not declared in include/b2b.h"
        cret, cargs = protos[sym]
        assert ret == cret, (sym, ret, cret)
        assert len(args) == len(cargs) == nvals, (sym, len(args), len(cargs), nvals)
        for k, (a, c) in enumerate(zip(args, cargs)):
            ok = a == c or (a == "ptr_void" and c in ("ptr_void",)) or (a == "ptr_u8" and c == "ptr_u8")
            assert ok, (sym, k, a, c)
        seen.add(sym)
    assert {"b2b_chain_run_f32", "b2b_chain_workspace_bytes", "b2b_planar_chain_vjp_f32", "b2b_radial_chain_vjp_f32",
            "b2b_batchnorm_train_fwd_f32", "b2b_allreduce_sum_f64", "b2b_status_string", "b2b_coupling_affine_vjp_f32",
            "b2b_batchnorm_eval_vjp_f32", "b2b_chain_sample_f32"} <= seen
    # the LayerDesc struct mirrors b2b_layer_desc field for field
    import re

    src = open(os.path.join(ROOT, "bijectors.jl_b200", "julia", "B200Bijectors.jl"), encoding="utf-8").read()
    body = src[src.index("struct LayerDesc"):src.index("end", src.index("struct LayerDesc"))]
    fields = re.findall(r"(\w+)::(\w+(?:\{\w+\})?)", body)
    assert [f for f, _ in fields] == ["kind", "inverse", "n0", "n1", "n2", "n3", "f0", "f1", "p0", "p1", "p2", "p3", "i0", "i1"]
    assert [t for _, t in fields] == ["Int32"] * 6 + ["Float32"] * 2 + ["CuPtr{Float32}"] * 4 + ["CuPtr{Int32}"] * 2
    # imports the load needs (the round-1 file used Distributions.logpdf, findnz, mean, var without importing them)
    for needed in ("import Distributions", "using SparseArrays: findnz", "using Statistics: mean, var"):
        assert needed in src
    # exactly one method per generic function takes a bare ComposedFunction on a CuMatrix (no dispatch ambiguity)
    for fn in ("with_logabsdet_jacobian", "transform", "logabsdetjac"):
        assert len(re.findall(rf"^function {fn}\(f::ComposedFunction, x::CuMatrix\{{Float32\}}\)", src, flags=re.M)) == 1
