"""Flow layers of the hot path: parameter structs with the reference's field names whose evaluation is a
launch of the matching libb2b.so kernel (include/b2b.h).  No arithmetic on the batch happens here.

  PlanarLayer{w,u,b}                 src/bijectors/planar_layer.jl:13-18
  RadialLayer{α_,β,z_0}              src/bijectors/radial_layer.jl:11-17
  RationalQuadraticSpline            src/bijectors/rational_quadratic_spline.jl:75-123
  PartitionMask / Coupling           src/bijectors/coupling.jl:51-118,178-259
  InvertibleBatchNorm                src/bijectors/normalise.jl:9-37
  Permute                            src/bijectors/permute.jl:84-157
  Stacked / elementwise / Shift / Scale   src/bijectors/stacked.jl, exp_log.jl, shift.jl, scale.jl
  LeakyReLU                          src/bijectors/leaky_relu.jl
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import B2BError, LayerDesc, LayerDesc64
from .interface import Bijector, Inverse, Transform


def _dev_f32(v, device, dtype=torch.float32) -> torch.Tensor:
    """Parameter tensor on `device`: Float32 (the hot path) or Float64 (the reference's own test precision; evaluated by
    b2b_chain_run_f64)."""
    if dtype not in (torch.float32, torch.float64):
        raise TypeError(f"parameters are Float32 or Float64, got {dtype}")
    if isinstance(v, torch.Tensor):
        t = v.detach()
    else:
        t = torch.as_tensor(np.asarray(v, dtype=np.float64 if dtype == torch.float64 else np.float32))
    t = t.to(device=device, dtype=dtype)
    return t.reshape(-1).contiguous() if t.dim() <= 1 else t.contiguous()


def _check_dtype(param: torch.Tensor, dtype, what: str) -> None:
    if param.dtype != dtype:
        raise TypeError(f"{what} has {param.dtype} parameters but the batch is {dtype}; construct the layer with dtype={dtype} "
                        "(Float32 is the hot path, Float64 the reference's test precision)")


def _dev_i32(v, device) -> torch.Tensor:
    return torch.as_tensor(np.asarray(v, dtype=np.int32)).to(device).contiguous()


def _desc(kind, inverse=False, **kw):
    """b2b_layer_desc, or b2b_layer_desc_f64 when the layer's parameters are Float64 tensors."""
    f64 = any(isinstance(v, torch.Tensor) and v.dtype == torch.float64 for v in kw.values()) or kw.pop("_f64", False)
    d = LayerDesc64() if f64 else LayerDesc()
    d.kind = kind
    d.inverse = 1 if inverse else 0
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            v = v.data_ptr()
        setattr(d, k, v)
    return d


class _ParamLayer(Bijector):
    """Common plumbing: parameters are float32 tensors on one device (`fmap`-style movement with .to())."""

    _fields: Tuple[str, ...] = ()

    def params(self) -> Dict[str, torch.Tensor]:
        """Functors.@functor analogue: the numerical parameters by reference field name."""
        return {k: getattr(self, k) for k in self._fields}

    def to(self, device):
        new = object.__new__(type(self))
        new.__dict__.update(self.__dict__)
        for k in self._fields:
            setattr(new, k, getattr(self, k).to(device))
        new._cache = {}
        return new

    @property
    def device(self):
        return getattr(self, self._fields[0]).device

    def _keepalive(self):
        return tuple(getattr(self, k) for k in self._fields)

    def __eq__(self, other):
        return type(self) is type(other) and all(
            torch.equal(getattr(self, k).cpu(), getattr(other, k).cpu()) for k in self._fields
        )

    __hash__ = object.__hash__


# --------------------------------------------------------------------------------------------------
class PlanarLayer(_ParamLayer):
    """f(z) = z + û·tanh(wᵀz + b)  (planar_layer.jl:73-80); logjac = log1p(wᵀû·sech²(wᵀz+b)) (:102-110).
    Inverse via find_alpha (:112-127,:160-185).  `b` may be a scalar or a 1-vector (first(b), :75)."""

    _fields = ("w", "u", "b")

    def __init__(self, w, u=None, b=None, device="cuda", generator: Optional[torch.Generator] = None, dtype=torch.float32):
        if isinstance(w, int):  # PlanarLayer(dims) : randn parameters (:23-28)
            dims = w
            w = torch.randn(dims, generator=generator)
            u = torch.randn(dims, generator=generator)
            b = torch.randn(1, generator=generator)
        self.w = _dev_f32(w, device, dtype)
        self.u = _dev_f32(u, device, dtype)
        self.b = _dev_f32(b, device, dtype)
        if self.w.numel() != self.u.numel():
            raise ValueError("w and u must have the same length")

    def _descs(self, inverse, D, dtype=torch.float32):
        if D != self.w.numel():
            raise ValueError(f"DimensionMismatch: PlanarLayer has {self.w.numel()} dims, input has {D}")
        _check_dtype(self.w, dtype, "PlanarLayer")
        d = _desc(_lib.PLANAR, inverse, p0=self.w, p1=self.u, p2=self.b)
        if not self.w.is_cuda:
            # host-resident parameters (the reference's own residency: plain Arrays, planar_layer.jl:13-18):
            # run_chain routes all-planar host-parameter chains to b2b_planar_chain_hostparams_f32
            d._host_planar = (self.w, self.u, self.b)
        return [d]


class RadialLayer(_ParamLayer):
    """f(z) = z + β̂/(α+r)·(z − z₀)  (radial_layer.jl:43-53,58-72); closed-form inverse (:88-102,:124-129)."""

    _fields = ("α_", "β", "z_0")

    def __init__(self, α_, β=None, z_0=None, device="cuda", generator: Optional[torch.Generator] = None, dtype=torch.float32):
        if isinstance(α_, int) and β is None:  # RadialLayer(dims) (:22-27)
            dims = α_
            α_ = torch.randn(1, generator=generator)
            β = torch.randn(1, generator=generator)
            z_0 = torch.randn(dims, generator=generator)
        setattr(self, "α_", _dev_f32(α_, device, dtype))
        setattr(self, "β", _dev_f32(β, device, dtype))
        self.z_0 = _dev_f32(z_0, device, dtype)

    def _descs(self, inverse, D, dtype=torch.float32):
        if D != self.z_0.numel():
            raise ValueError(f"DimensionMismatch: RadialLayer has {self.z_0.numel()} dims, input has {D}")
        _check_dtype(self.z_0, dtype, "RadialLayer")
        return [_desc(_lib.RADIAL, inverse, p0=getattr(self, "α_"), p1=getattr(self, "β"), p2=self.z_0)]


class RationalQuadraticSpline(_ParamLayer):
    """Neural-spline-flow element-wise bijector on [-B, B]^D.

    RationalQuadraticSpline(widths, heights, derivatives)     processed knots, (D × K+1) each  (:75-97)
    RationalQuadraticSpline(widths, heights, derivatives, B)  raw parameters (D×K, D×K, D×(K-1)) are
        normalised on the host exactly as the reference constructor does (:109-123): softmax → cumsum →
        [-B, B] knots, softplus derivatives with unit end slopes.  (Tiny, once per construction.)
    Matrices are given with the reference's index order: row = dimension, column = knot.
    """

    _fields = ("widths", "heights", "derivatives")

    def __init__(self, widths, heights, derivatives, B=None, device="cuda", dtype=torch.float32):
        npdt = np.float64 if dtype == torch.float64 else np.float32
        w = np.asarray(widths.detach().cpu() if isinstance(widths, torch.Tensor) else widths, dtype=npdt)
        h = np.asarray(heights.detach().cpu() if isinstance(heights, torch.Tensor) else heights, dtype=npdt)
        d = np.asarray(derivatives.detach().cpu() if isinstance(derivatives, torch.Tensor) else derivatives,
                       dtype=npdt)
        if w.ndim == 1:
            w, h, d = w[None, :], h[None, :], d[None, :]
        if B is not None:
            w, h, d = _rqs_normalise(w, h, d, float(B), npdt)
        # struct asserts (:93-94)
        assert w.shape[1] == h.shape[1] == d.shape[1], "widths, heights and derivatives need the same number of knots"
        assert np.all(d > 0), "derivatives need to be positive"
        self.K1 = int(w.shape[1])
        self.D = int(w.shape[0])
        # device copies keep Julia's column-major (D × K1) memory order == knot-major [k][i]
        self.widths = _dev_f32(np.ascontiguousarray(w.T), device, dtype)
        self.heights = _dev_f32(np.ascontiguousarray(h.T), device, dtype)
        self.derivatives = _dev_f32(np.ascontiguousarray(d.T), device, dtype)

    def knots(self):
        """(widths, heights, derivatives) as (D × K+1) numpy arrays in the reference's index order."""
        return tuple(getattr(self, k).cpu().numpy().T.copy() for k in self._fields)

    def _descs(self, inverse, D, dtype=torch.float32):
        if D != self.D:
            raise ValueError(f"DimensionMismatch: RationalQuadraticSpline has {self.D} dims, input has {D}")
        _check_dtype(self.widths, dtype, "RationalQuadraticSpline")
        return [_desc(_lib.RQS, inverse, p0=self.widths, p1=self.heights, p2=self.derivatives, n0=self.K1)]


def _rqs_normalise(w, h, d, B, f=np.float32):
    """Host restatement of the normalising constructor (rational_quadratic_spline.jl:109-123) in the parameter eltype."""

    def softmax(v):
        e = np.exp(v - v.max(axis=1, keepdims=True))
        return (e / e.sum(axis=1, keepdims=True)).astype(f)

    def softplus(v):
        v = v.astype(np.float64)
        return np.where(v > 0, v + np.log1p(np.exp(-np.abs(v))), np.log1p(np.exp(-np.abs(v)))).astype(f)

    n = w.shape[0]
    ws = np.concatenate([np.zeros((n, 1), f), softmax(w)], axis=1)
    hs = np.concatenate([np.zeros((n, 1), f), softmax(h)], axis=1)
    ds = np.concatenate([np.ones((n, 1), f), softplus(d), np.ones((n, 1), f)], axis=1)
    W = (f(2 * B) * np.cumsum(ws, axis=1, dtype=f) - f(B)).astype(f)
    H = (f(2 * B) * np.cumsum(hs, axis=1, dtype=f) - f(B)).astype(f)
    return W, H, ds


# --------------------------------------------------------------------------------------------------
class PartitionMask:
    """PartitionMask(n, indices_1, indices_2[, indices_3]) / PartitionMask(n, indices)
    (coupling.jl:51-118).  Indices are 1-BASED like the reference; the sparse 0/1 selector matrices are
    kept as index lists (partition / combine are pure index movement, bit-exact)."""

    def __init__(self, n: int, indices_1, indices_2=None, indices_3=None):
        i1 = [int(i) for i in indices_1]
        if indices_2 is None and indices_3 is None:
            i2 = [i for i in range(1, n + 1) if i not in set(i1)]  # :107-115
            i3 = []
        elif indices_3 is None:
            i2 = [int(i) for i in indices_2]
            i3 = [i for i in range(1, n + 1) if i not in set(i1) | set(i2)]  # :85-92
        elif indices_2 is None:
            i3 = [int(i) for i in indices_3]
            i2 = [i for i in range(1, n + 1) if i not in set(i1) | set(i3)]  # :94-101
        else:
            i2, i3 = [int(i) for i in indices_2], [int(i) for i in indices_3]
        allidx = i1 + i2 + i3
        if any(i < 1 or i > n for i in allidx) or len(set(allidx)) != len(allidx):
            raise ValueError("PartitionMask indices must be disjoint and within 1:n")
        self.n, self.indices_1, self.indices_2, self.indices_3 = n, i1, i2, i3

    def __eq__(self, o):
        return (self.n, self.indices_1, self.indices_2, self.indices_3) == (o.n, o.indices_1, o.indices_2, o.indices_3)


class AffineConditioner:
    """The recognised coupling law θ(x₂) = Shift(t) ∘ Scale(exp.(s)) with [s; t] = W·x₂ + c
    (Scale scale.jl:13,31; Shift shift.jl:14,21).  W is (2·n1 × n2) in the reference's index order."""

    def __init__(self, W, c=None, device="cuda", dtype=torch.float32):
        npdt = np.float64 if dtype == torch.float64 else np.float32
        Wn = np.asarray(W.detach().cpu() if isinstance(W, torch.Tensor) else W, dtype=npdt)
        if Wn.ndim != 2 or Wn.shape[0] % 2:
            raise ValueError("W must be (2*n1, n2)")
        self.n1, self.n2 = Wn.shape[0] // 2, Wn.shape[1]
        self.W = _dev_f32(np.ascontiguousarray(Wn.T), device, dtype)  # column-major (2n1 × n2)
        cn = np.zeros(Wn.shape[0], npdt) if c is None else np.asarray(
            c.detach().cpu() if isinstance(c, torch.Tensor) else c, dtype=npdt)
        self.c = _dev_f32(cn, device, dtype)

    def to(self, device):
        new = object.__new__(AffineConditioner)
        new.n1, new.n2 = self.n1, self.n2
        new.W, new.c = self.W.to(device), self.c.to(device)
        return new


class Coupling(_ParamLayer):
    """Coupling(θ, mask) (coupling.jl:178-181).  θ is an arbitrary closure in the reference; the device
    path supports the recognised :class:`AffineConditioner` and raises for anything else (no CPU fallback)."""

    _fields = ()

    def __init__(self, θ, mask, device="cuda"):
        if isinstance(mask, int):  # Coupling(θ, n): first n÷2 rows transformed (:183-186)
            mask = PartitionMask(mask, range(1, mask // 2 + 1))
        if not isinstance(θ, AffineConditioner):
            raise B2BError(_lib.B2B_EUNSUPPORTED, "Coupling: only AffineConditioner laws run on the device path")
        if θ.n1 != len(mask.indices_1) or θ.n2 != len(mask.indices_2):
            raise ValueError("conditioner shape does not match the PartitionMask")
        self.θ, self.mask = θ, mask
        self._idx1 = _dev_i32(np.asarray(mask.indices_1) - 1, θ.W.device)
        self._idx2 = _dev_i32(np.asarray(mask.indices_2) - 1, θ.W.device)

        def first_row(idx):  # 0-based first row when the list is a contiguous increasing range, else -1
            a = np.asarray(idx)
            return int(a[0] - 1) if len(a) and np.array_equal(a, np.arange(a[0], a[0] + len(a))) else -1

        self._row1, self._row2 = first_row(mask.indices_1), first_row(mask.indices_2)

    @property
    def device(self):
        return self.θ.W.device

    def to(self, device):
        """fmap-style movement: the conditioner (W, c) AND the index lists follow."""
        return Coupling(self.θ.to(device), self.mask)

    def _keepalive(self):
        return (self.θ.W, self.θ.c, self._idx1, self._idx2)

    def _descs(self, inverse, D, dtype=torch.float32):
        if D != self.mask.n:
            raise ValueError(f"DimensionMismatch: Coupling mask has {self.mask.n} dims, input has {D}")
        _check_dtype(self.θ.W, dtype, "Coupling")
        return [_desc(_lib.COUPLING_AFFINE, inverse, p0=self.θ.W, p1=self.θ.c, i0=self._idx1, i1=self._idx2,
                      n0=self.θ.n1, n1=self.θ.n2, n2=self._row1, n3=self._row2)]

    def __eq__(self, o):
        return isinstance(o, Coupling) and self.mask == o.mask and torch.equal(self.θ.W, o.θ.W) and torch.equal(self.θ.c, o.θ.c)

    __hash__ = object.__hash__


def coupling(cl: Coupling):
    """coupling(cl) = cl.θ (coupling.jl:193)."""
    return cl.θ


# --------------------------------------------------------------------------------------------------
class InvertibleBatchNorm(_ParamLayer):
    """InvertibleBatchNorm(chs; eps=1f-5, mtm=1f-1) (normalise.jl:9-37); eval mode (:61-67,:74-86).
    The reference's global `istraining()` switch (:7) is the explicit `training` flag here: a training-mode layer
    computes batch statistics and updates its moving statistics in place (`train_forward`, :51-60)."""

    _fields = ("b", "logs", "m", "v")

    def __init__(self, chs=None, *, b=None, logs=None, m=None, v=None, eps=1e-5, mtm=1e-1, device="cuda",
                 training=False, dtype=torch.float32):
        if chs is not None:
            b, logs, m, v = np.zeros(chs), np.zeros(chs), np.zeros(chs), np.ones(chs)
        self.b, self.logs, self.m, self.v = (_dev_f32(t, device, dtype) for t in (b, logs, m, v))
        self.eps, self.mtm, self.training = float(np.float32(eps)), float(np.float32(mtm)), training

    def _descs(self, inverse, D, dtype=torch.float32):
        if D != self.b.numel():
            # error text of normalise.jl:43-45
            raise RuntimeError(f"InvertibleBatchNorm expected {self.b.numel()} channels, got {D}")
        if self.training:
            raise B2BError(_lib.B2B_EUNSUPPORTED, "InvertibleBatchNorm in training mode cannot be fused into a chain "
                                                  "or inverted (normalise.jl:75); call it on its own")
        _check_dtype(self.b, dtype, "InvertibleBatchNorm")
        return [_desc(_lib.BATCHNORM, inverse, p0=self.b, p1=self.logs, p2=self.m, p3=self.v, f0=self.eps)]

    def _expanded(self, S: int):
        """The same eval-mode layer with every channel's parameters repeated S times (rows s + S·c of a (d₁⋯d_k·C)×B
        view of an array with k leading spatial axes; interface._batchnorm_nd).  Rebuilt on every call: the fields are
        trainable and may have changed."""
        out = InvertibleBatchNorm.__new__(InvertibleBatchNorm)
        out.b, out.logs, out.m, out.v = (torch.repeat_interleave(p, S) for p in (self.b, self.logs, self.m, self.v))
        out.eps, out.mtm, out.training = self.eps, self.mtm, False
        return out

    def train_forward(self, x, comm=None):
        """with_logabsdet_jacobian(bn, x) with istraining() == true (normalise.jl:51-69): batch statistics (over all
        ranks of `comm`, a distributed.Communicator, when given), in-place moving-average update of self.m / self.v,
        output and logjac from the batch statistics."""
        from .interface import _batch_view, _stream, colmajor_empty
        from ._lib import check, lib

        D, N, ldx = _batch_view(x)
        if D != self.b.numel():
            raise RuntimeError(f"InvertibleBatchNorm expected {self.b.numel()} channels, got {D}")
        y = colmajor_empty(D, N, x.device)
        lj = torch.empty((N,), dtype=torch.float32, device=x.device)
        nbytes = lib().b2b_batchnorm_train_workspace_bytes(D)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
        handle = comm.handle if (comm is not None and getattr(comm, "handle", None) is not None) else None
        check(lib().b2b_batchnorm_train_fwd_f32(x.data_ptr(), y.data_ptr(), lj.data_ptr(), self.b.data_ptr(),
                                                self.logs.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.eps,
                                                self.mtm, D, N, ldx, D, 0, handle, ws.data_ptr(), nbytes, _stream()),
              "b2b_batchnorm_train_fwd_f32")
        return y, lj


# --------------------------------------------------------------------------------------------------
class Permute(_ParamLayer):
    """Permute(indices) / Permute(n, src=>dst...) / Permute(n, [srcs]=>[dsts]...) / Permute(A)
    (permute.jl:84-150).  1-based like the reference.  transform = A*x as index movement (:152),
    batched logjac = zeros(N) (:155)."""

    _fields = ()

    def __init__(self, *args, device="cuda"):
        if len(args) == 1 and np.ndim(args[0]) == 1:
            dst = self._from_indices(list(args[0]))
        elif len(args) == 1 and np.ndim(args[0]) == 2:
            dst = self._from_matrix(np.asarray(args[0]))
        else:
            dst = self._from_pairs(int(args[0]), args[1:])
        self.dst_of_src = dst  # 0-based numpy: y[dst[i]] = x[i]
        self._dst = _dev_i32(dst, device)

    @staticmethod
    def _from_indices(indices):
        n = len(indices)
        if sorted(indices) != list(range(1, n + 1)):
            raise ValueError("ArgumentError: indices is not a permutation of 1:n")
        return np.asarray(indices, dtype=np.int64) - 1  # A[idx, i] = 1  (:95-97)

    @staticmethod
    def _from_matrix(A):
        if A.shape[0] != A.shape[1] or not (np.all((A == 0) | (A == 1)) and np.all(A.sum(0) == 1) and np.all(A.sum(1) == 1)):
            raise ValueError("ArgumentError: not a permutation matrix")
        return np.argmax(A, axis=0).astype(np.int64)

    @staticmethod
    def _from_pairs(n, pairs):
        dst = np.arange(n, dtype=np.int64)
        dests, sources = set(), set()
        for src, d in pairs:
            srcs = list(src) if np.ndim(src) else [src]
            dsts = list(d) if np.ndim(d) else [d]
            if len(srcs) != len(dsts):
                raise ValueError(f"ArgumentError: {srcs} => {dsts} is not bijective")  # :132
            for s_, d_ in zip(srcs, dsts):
                if d_ in dests:
                    raise ValueError(f"ArgumentError: {d_} used more than once")
                if s_ in sources:
                    raise ValueError(f"ArgumentError: {s_} used more than once")
                dests.add(d_)
                sources.add(s_)
                dst[s_ - 1] = d_ - 1
        if (sources & dests) != (sources | dests):  # :119,:145
            raise ValueError(f"ArgumentError: {sources} ∩ {dests} ≠ {sources} ∪ {dests}")
        return dst

    @property
    def A(self):
        n = len(self.dst_of_src)
        A = np.zeros((n, n))
        A[self.dst_of_src, np.arange(n)] = 1.0
        return A

    @property
    def device(self):
        return self._dst.device

    def to(self, device):
        new = object.__new__(Permute)
        new.dst_of_src = self.dst_of_src
        new._dst = self._dst.to(device)
        return new

    def _keepalive(self):
        return (self._dst,)

    def _descs(self, inverse, D, dtype=torch.float32):
        if D != len(self.dst_of_src):
            raise ValueError(f"DimensionMismatch: Permute has {len(self.dst_of_src)} dims, input has {D}")
        return [_desc(_lib.PERMUTE, inverse, i0=self._dst, _f64=dtype == torch.float64)]

    def __eq__(self, o):
        return isinstance(o, Permute) and np.array_equal(self.dst_of_src, o.dst_of_src)

    __hash__ = object.__hash__


# --------------------------------------------------------------------------------------------------
class Elementwise(Bijector):
    """elementwise(f) = Base.Fix1(broadcast, f) (interface.jl:33) for f ∈ {exp, log, identity}.

    On a HOST vector (BASELINE config 1: Float64, length 1024 -- API plumbing, not the hot path) it returns
    (f.(x), Σ logjac) with the reference's scalar-sum semantics (exp_log.jl:6,9).  Inside Stacked, and on
    device batches, rows are evaluated by the stacked_elementwise kernel with a per-COLUMN logjac."""

    def __init__(self, f: str):
        if f not in ("exp", "log", "identity"):
            raise B2BError(_lib.B2B_EUNSUPPORTED, f"elementwise({f})")
        self.f = f

    code = property(lambda s: {"exp": _lib.EW_EXP, "log": _lib.EW_LOG, "identity": _lib.EW_IDENTITY}[s.f])
    a = 0.0

    def _inverse(self):
        return Elementwise({"exp": "log", "log": "exp", "identity": "identity"}[self.f])

    def _host_wladj(self, x):
        xt = torch.as_tensor(x)
        if self.f == "exp":
            return torch.exp(xt), xt.sum()
        if self.f == "log":
            return torch.log(xt), -torch.log(xt).sum()
        return xt, xt.new_zeros(())

    def _descs(self, inverse, D, dtype=torch.float32):
        return _as_stacked(self, D, dtype)._descs(inverse, D, dtype)

    def __eq__(self, o):
        return isinstance(o, Elementwise) and o.f == self.f

    __hash__ = object.__hash__


def elementwise(f):
    name = f if isinstance(f, str) else {math.exp: "exp", math.log: "log", np.exp: "exp", np.log: "log",
                                         torch.exp: "exp", torch.log: "log"}.get(f, getattr(f, "__name__", str(f)))
    return Elementwise(name)


def _as_stacked(b, D, dtype=torch.float32):
    """A whole-column elementwise law is a one-block Stacked; cached on the object so the device tables
    outlive the asynchronous launch."""
    cache = b.__dict__.setdefault("_stacked_cache", {})
    key = (D, dtype)
    if key not in cache:
        cache[key] = Stacked([b], [(1, D)], dtype=dtype)
    return cache[key]


class Shift(Bijector):
    """Shift(a): y = a .+ x, logjac 0 (shift.jl:4-24); scalar `a` inside Stacked blocks."""

    def __init__(self, a):
        self.a = float(a)

    code = _lib.EW_SHIFT

    def _inverse(self):
        return Shift(-self.a)  # shift.jl:12

    def _descs(self, inverse, D, dtype=torch.float32):
        return _as_stacked(self, D, dtype)._descs(inverse, D, dtype)

    def __eq__(self, o):
        return isinstance(o, Shift) and o.a == self.a

    __hash__ = object.__hash__


class Scale(Bijector):
    """Scale(a): y = a .* x, logjac = log|a| per element (scale.jl:1-39); scalar `a`."""

    def __init__(self, a):
        self.a = float(a)

    code = _lib.EW_SCALE

    def _descs(self, inverse, D, dtype=torch.float32):
        return _as_stacked(self, D, dtype)._descs(inverse, D, dtype)

    def __eq__(self, o):
        return isinstance(o, Scale) and o.a == self.a

    __hash__ = object.__hash__


class LeakyReLU(Bijector):
    """LeakyReLU(α): x ↦ x if x ≥ 0 else αx, α > 0 (leaky_relu.jl:9-29); inverse = LeakyReLU(1/α) (:16).
    Batched logjac is per column (the reference sums over the whole array)."""

    def __init__(self, α):
        self.a = float(α)
        if not self.a > 0:
            raise ValueError("LeakyReLU needs α > 0")

    code = _lib.EW_LEAKY_RELU
    α = property(lambda s: s.a)

    def _inverse(self):
        return LeakyReLU(1.0 / self.a)

    def _descs(self, inverse, D, dtype=torch.float32):
        return _as_stacked(self, D, dtype)._descs(inverse, D, dtype)

    def __eq__(self, o):
        return isinstance(o, LeakyReLU) and o.a == self.a

    __hash__ = object.__hash__


class Logit(Bijector):
    """Logit(a, b): y = logit((x − a)/(b − a)) element-wise, logjac = −Σ log((x − a)(b − x)/(b − a))
    (logit.jl:4-29); scalar bounds; the reference's `bounded flow` building block (docs/src/flows.md:25-36)."""

    def __init__(self, a, b):
        self.a, self.b = float(a), float(b)
        if not self.b > self.a:
            raise ValueError("Logit needs a < b")

    code = _lib.EW_LOGIT

    def _descs(self, inverse, D, dtype=torch.float32):
        return _as_stacked(self, D, dtype)._descs(inverse, D, dtype)

    def __eq__(self, o):
        return isinstance(o, Logit) and (o.a, o.b) == (self.a, self.b)  # logit.jl:12

    __hash__ = object.__hash__


class TruncatedBijector(Bijector):
    """TruncatedBijector(lb, ub) (truncated.jl:4-91): clamp to [lb, ub], then logit((x−lb)/(ub−lb)) / log(x−lb) /
    log(ub−x) / identity depending on which bounds are finite (±inf allowed); scalar bounds."""

    def __init__(self, lb, ub):
        self.a, self.b = float(lb), float(ub)
        if not self.b > self.a:
            raise ValueError("TruncatedBijector needs lb < ub")

    lb = property(lambda s: s.a)
    ub = property(lambda s: s.b)
    code = _lib.EW_TRUNCATED

    def _descs(self, inverse, D, dtype=torch.float32):
        return _as_stacked(self, D, dtype)._descs(inverse, D, dtype)

    def __eq__(self, o):
        return isinstance(o, TruncatedBijector) and (o.a, o.b) == (self.a, self.b)

    __hash__ = object.__hash__


class Stacked(Transform):
    """Stacked(bs, ranges): bs[i] applied to rows ranges[i] (1-based inclusive (lo, hi) like Julia
    UnitRanges; stacked.jl:25-59).  Device scope: elementwise blocks (exp, log, identity, Shift, Scale, LeakyReLU,
    Logit, TruncatedBijector)."""

    def __init__(self, bs, ranges=None, device="cuda", dtype=torch.float32):
        bs = list(bs)
        if ranges is None:
            ranges = [(i + 1, i + 1) for i in range(len(bs))]  # Stacked(bs...) = ranges i:i (:49)
        ranges = [(int(lo), int(hi)) for lo, hi in ranges]
        if len(bs) != len(ranges):
            raise ValueError("length(bs) == length(ranges) needs to be true")
        for b in bs:
            if not isinstance(b, (Elementwise, Shift, Scale, LeakyReLU, Logit, TruncatedBijector)) and b is not None:
                raise B2BError(_lib.B2B_EUNSUPPORTED, f"Stacked block {type(b).__name__}")
        self.bs, self.ranges_in = bs, ranges
        self.length_in = sum(hi - lo + 1 for lo, hi in ranges)
        self.length_out = self.length_in
        code = np.zeros(self.length_in, np.int32)
        a = np.zeros(self.length_in, np.float64)
        b2 = np.zeros(self.length_in, np.float64)
        for b, (lo, hi) in zip(bs, ranges):
            code[lo - 1:hi] = _lib.EW_IDENTITY if b is None else b.code
            a[lo - 1:hi] = 0.0 if b is None else b.a
            b2[lo - 1:hi] = getattr(b, "b", 0.0) if b is not None else 0.0
        self._code = _dev_i32(code, device)
        self._a = _dev_f32(a.astype(np.float64), device, dtype) if dtype == torch.float64 else _dev_f32(a, device)
        self._b = _dev_f32(b2.astype(np.float64), device, dtype) if dtype == torch.float64 else _dev_f32(b2, device)
        self._dtype = dtype

    def _keepalive(self):
        return (self._code, self._a, self._b)

    def to(self, device):
        new = object.__new__(Stacked)
        new.__dict__.update(self.__dict__)
        new._code, new._a, new._b = self._code.to(device), self._a.to(device), self._b.to(device)
        return new

    @property
    def device(self):
        return self._code.device

    def _descs(self, inverse, D, dtype=torch.float32):
        if self.length_in != D:
            raise RuntimeError(f"input length mismatch ({self.length_in} != {D})")  # stacked.jl:158-160,243-245
        _check_dtype(self._a, dtype, "Stacked")
        return [_desc(_lib.STACKED_EW, inverse, i0=self._code, p0=self._a, p1=self._b)]
