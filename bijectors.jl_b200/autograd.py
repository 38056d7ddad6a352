"""torch.autograd glue for planar flows: makes ``with_logabsdet_jacobian`` of a PlanarLayer chain (either direction)
differentiable by routing the backward pass to ``b2b_planar_chain_vjp_f32`` -- the role the AD extensions play for the
reference (ext/BijectorsChainRulesCoreExt.jl, docs/src/flows.md:93-100).  No arithmetic on batches happens here."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from .interface import Composed, batchnorm_vjp, colmajor_empty, coupling_vjp, inverse, planar_chain_vjp, radial_chain_vjp, rqs_vjp, run_chain
from .layers import (AffineConditioner, Coupling, InvertibleBatchNorm, PartitionMask, PlanarLayer, RadialLayer,
                     RationalQuadraticSpline)


_VJP_DIMS = (32, 64, 128)


def _colmajor(t: torch.Tensor) -> torch.Tensor:
    """A (D, N) tensor with strides (1, D) (no copy when it already has them)."""
    if t.dim() == 2 and t.stride(0) == 1 and (t.shape[1] == 1 or t.stride(1) == t.shape[0]):
        return t
    out = colmajor_empty(t.shape[0], t.shape[1], t.device)
    out.copy_(t)
    return out


class _PlanarChainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inv: bool, *wub):
        L = len(wub) // 3
        layers = [PlanarLayer(wub[3 * l].detach(), wub[3 * l + 1].detach(), wub[3 * l + 2].detach()) for l in range(L)]
        flow = Composed(*layers)
        t = inverse(flow) if inv else flow
        xc = _colmajor(x.detach())
        y, lj = run_chain(t, xc)
        ctx.t, ctx.inv, ctx.L, ctx.layers = t, inv, L, layers
        ctx.save_for_backward(xc)
        return y, lj

    @staticmethod
    def backward(ctx, ybar, ljbar):
        (xc,) = ctx.saved_tensors
        D, N = xc.shape
        yb = _colmajor(ybar) if ybar is not None else _colmajor(torch.zeros((D, N), device=xc.device))
        lb = ljbar.contiguous() if ljbar is not None else None
        t = ctx.t
        Dp = next((d for d in _VJP_DIMS if d >= D), None)
        if Dp is None:
            raise NotImplementedError(f"planar VJP kernels cover D <= {_VJP_DIMS[-1]} (got {D})")
        if Dp != D:
            # The reverse-mode kernels are built for D in {32, 64, 128}.  A planar layer on zero-padded rows is the same
            # map (w, u padded with zeros: wᵀz, wᵀu, ‖w‖² and the first D rows of û are unchanged), so smaller flows
            # (the reference's own example is D = 2, docs/src/flows.md:40-60) run embedded in the next supported D.
            pad = lambda v: torch.cat([v, v.new_zeros(Dp - D)])
            layers = [PlanarLayer(pad(l.w), pad(l.u), l.b) for l in ctx.layers]
            t = inverse(Composed(*layers)) if ctx.inv else Composed(*layers)
            xp, ybp = colmajor_empty(Dp, N, xc.device), colmajor_empty(Dp, N, xc.device)
            xp.zero_()
            ybp.zero_()
            xp[:D].copy_(xc)
            ybp[:D].copy_(yb)
            xc, yb = xp, ybp
        xbar, grads = planar_chain_vjp(t, xc, yb, lb)
        if Dp != D:
            xbar = xbar[:D]
            grads = [{"w": g["w"][:D], "u": g["u"][:D], "b": g["b"]} for g in grads]
        if ctx.inv:  # application order of inverse(flow) is the flow's layers reversed
            grads = grads[::-1]
        flat: List[torch.Tensor] = []
        for g in grads:
            flat += [g["w"], g["u"], g["b"]]
        return (xbar, None, *flat)


class PlanarFlow(torch.nn.Module):
    """A trainable ∘-chain of L PlanarLayers (planar_layer.jl:13-28: randn-initialised w, u, b) on the device."""

    def __init__(self, dims: int, n_layers: int, device="cuda", generator=None, scale: float = 1.0):
        super().__init__()
        mk = lambda n: torch.nn.Parameter((torch.randn(n, generator=generator) * scale).to(device))
        self.w = torch.nn.ParameterList([mk(dims) for _ in range(n_layers)])
        self.u = torch.nn.ParameterList([mk(dims) for _ in range(n_layers)])
        self.b = torch.nn.ParameterList([mk(1) for _ in range(n_layers)])

    def _wub(self) -> Sequence[torch.Tensor]:
        out = []
        for w, u, b in zip(self.w, self.u, self.b):
            out += [w, u, b]
        return out

    def layers(self) -> List[PlanarLayer]:
        return [PlanarLayer(w.detach(), u.detach(), b.detach()) for w, u, b in zip(self.w, self.u, self.b)]

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """with_logabsdet_jacobian(flow, x), differentiable w.r.t. x and the parameters."""
        return _PlanarChainFn.apply(x, False, *self._wub())

    def inverse(self, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """with_logabsdet_jacobian(inverse(flow), y), differentiable (find_alpha through its implicit rule)."""
        return _PlanarChainFn.apply(y, True, *self._wub())


class _RadialChainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inv: bool, *abz):
        L = len(abz) // 3
        flow = Composed(*[RadialLayer(abz[3 * l].detach(), abz[3 * l + 1].detach(), abz[3 * l + 2].detach()) for l in range(L)])
        t = inverse(flow) if inv else flow
        xc = _colmajor(x.detach())
        y, lj = run_chain(t, xc)
        ctx.t, ctx.inv = t, inv
        ctx.save_for_backward(xc)
        return y, lj

    @staticmethod
    def backward(ctx, ybar, ljbar):
        (xc,) = ctx.saved_tensors
        D, N = xc.shape
        yb = _colmajor(ybar) if ybar is not None else _colmajor(torch.zeros((D, N), device=xc.device))
        xbar, grads = radial_chain_vjp(ctx.t, xc, yb, ljbar.contiguous() if ljbar is not None else None)
        if ctx.inv:  # application order of inverse(flow) is the flow's layers reversed
            grads = grads[::-1]
        flat: List[torch.Tensor] = []
        for g in grads:
            flat += [g["α_"], g["β"], g["z_0"]]
        return (xbar, None, *flat)


class RadialFlow(torch.nn.Module):
    """A trainable ∘-chain of L RadialLayers (radial_layer.jl:11-27: randn-initialised α_, β, z_0), forward direction
    (the sampling / variational path): ``with_logabsdet_jacobian`` differentiable w.r.t. x and the raw parameters."""

    def __init__(self, dims: int, n_layers: int, device="cuda", generator=None):
        super().__init__()
        mk = lambda n: torch.nn.Parameter(torch.randn(n, generator=generator).to(device))
        self.alpha_ = torch.nn.ParameterList([mk(1) for _ in range(n_layers)])
        self.beta = torch.nn.ParameterList([mk(1) for _ in range(n_layers)])
        self.z_0 = torch.nn.ParameterList([mk(dims) for _ in range(n_layers)])

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return _RadialChainFn.apply(x, False, *self._abz())

    def _abz(self):
        abz = []
        for a, b, z in zip(self.alpha_, self.beta, self.z_0):
            abz += [a, b, z]
        return abz

    def inverse(self, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """with_logabsdet_jacobian(inverse(flow), y), differentiable (compute_r through its implicit rule): the
        logpdf / NLL path of a radial flow."""
        return _RadialChainFn.apply(y, True, *self._abz())


# ---- RealNVP: affine Coupling + eval-mode InvertibleBatchNorm blocks (BASELINE config 5) ---------------------------------
class _CouplingFn(torch.autograd.Function):
    """with_logabsdet_jacobian of ONE affine coupling layer (either direction); backward = b2b_coupling_affine_vjp_f32."""

    @staticmethod
    def forward(ctx, x, W, c, mask, inv: bool):
        cond = AffineConditioner.__new__(AffineConditioner)
        cond.n1, cond.n2 = W.shape[0] // 2, W.shape[1]
        cond.W, cond.c = W.detach().t().contiguous(), c.detach().contiguous()  # device layout: column-major (2n1 × n2)
        lay = Coupling(cond, mask)
        t = inverse(lay) if inv else lay
        xc = _colmajor(x.detach())
        y, lj = run_chain(t, xc)
        ctx.t = t
        ctx.save_for_backward(xc)
        return y, lj

    @staticmethod
    def backward(ctx, ybar, ljbar):
        (xc,) = ctx.saved_tensors
        D, N = xc.shape
        yb = _colmajor(ybar) if ybar is not None else _colmajor(torch.zeros((D, N), device=xc.device))
        xbar, g = coupling_vjp(ctx.t, xc, yb, ljbar.contiguous() if ljbar is not None else None)
        return xbar, g["W"], g["c"], None, None


class _BatchNormFn(torch.autograd.Function):
    """with_logabsdet_jacobian of ONE eval-mode InvertibleBatchNorm (either direction); backward = b2b_batchnorm_eval_vjp_f32."""

    @staticmethod
    def forward(ctx, x, b, logs, m, v, eps: float, inv: bool):
        lay = InvertibleBatchNorm(b=b.detach(), logs=logs.detach(), m=m, v=v, eps=eps, device=x.device)
        t = inverse(lay) if inv else lay
        xc = _colmajor(x.detach())
        y, lj = run_chain(t, xc)
        ctx.t = t
        ctx.save_for_backward(xc)
        return y, lj

    @staticmethod
    def backward(ctx, ybar, ljbar):
        (xc,) = ctx.saved_tensors
        D, N = xc.shape
        yb = _colmajor(ybar) if ybar is not None else _colmajor(torch.zeros((D, N), device=xc.device))
        xbar, g = batchnorm_vjp(ctx.t, xc, yb, ljbar.contiguous() if ljbar is not None else None)
        return xbar, g["b"], g["logs"], None, None, None, None


class RealNVP(torch.nn.Module):
    """A trainable RealNVP flow: `n_blocks` x (affine Coupling with alternating half masks + eval-mode InvertibleBatchNorm),
    the structure of BASELINE config 5.  ``forward(x)`` / ``inverse(y)`` return (result, logjac), differentiable w.r.t.
    the input and the parameters W, c (conditioners) and b, logs (BatchNorm; m, v are statistics); ``nll(y)`` is the
    training objective of docs/src/flows.md:74-77 with a standard-normal base."""

    def __init__(self, dims: int, n_blocks: int, device="cuda", generator=None, scale: float = 0.05):
        super().__init__()
        h = dims // 2
        self.dims, self.masks = dims, []
        Ws, cs, bs, ls = [], [], [], []
        for l in range(n_blocks):
            first = l % 2 == 0
            idx1 = list(range(1, h + 1)) if first else list(range(h + 1, dims + 1))
            idx2 = list(range(h + 1, dims + 1)) if first else list(range(1, h + 1))
            self.masks.append(PartitionMask(dims, idx1, idx2))
            n1, n2 = len(idx1), len(idx2)
            Ws.append(torch.nn.Parameter((torch.randn((2 * n1, n2), generator=generator) * scale / n2 ** 0.5).to(device)))
            cs.append(torch.nn.Parameter(torch.zeros(2 * n1, device=device)))
            bs.append(torch.nn.Parameter(torch.zeros(dims, device=device)))
            ls.append(torch.nn.Parameter(torch.zeros(dims, device=device)))
        self.W, self.c = torch.nn.ParameterList(Ws), torch.nn.ParameterList(cs)
        self.b, self.logs = torch.nn.ParameterList(bs), torch.nn.ParameterList(ls)
        self.register_buffer("m", torch.zeros(n_blocks, dims, device=device))
        self.register_buffer("v", torch.ones(n_blocks, dims, device=device))
        self.eps = 1e-5

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        lj = None
        for l in range(len(self.W)):
            x, l1 = _CouplingFn.apply(x, self.W[l], self.c[l], self.masks[l], False)
            x, l2 = _BatchNormFn.apply(x, self.b[l], self.logs[l], self.m[l], self.v[l], self.eps, False)
            lj = l1 + l2 if lj is None else lj + l1 + l2
        return x, lj

    def inverse(self, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        lj = None
        for l in reversed(range(len(self.W))):
            y, l2 = _BatchNormFn.apply(y, self.b[l], self.logs[l], self.m[l], self.v[l], self.eps, True)
            y, l1 = _CouplingFn.apply(y, self.W[l], self.c[l], self.masks[l], True)
            lj = l1 + l2 if lj is None else lj + l1 + l2
        return y, lj

    def nll(self, y: torch.Tensor) -> torch.Tensor:
        """−Σ_n logpdf(transformed(MvNormal(0, I), flow), y_n) (transformed_distribution.jl:165-169)."""
        x, lj = self.inverse(y)
        base = -0.5 * (x * x).sum(0) - 0.5 * self.dims * 1.8378770664093453
        return -(base + lj).sum()


# ---- RationalQuadraticSpline -------------------------------------------------------------------------------------------
class _SplineFn(torch.autograd.Function):
    """with_logabsdet_jacobian of ONE RationalQuadraticSpline given its PROCESSED knots (D × K+1 each), either direction;
    backward = b2b_rqs_vjp_f32 (cotangents of the input and of the three knot arrays)."""

    @staticmethod
    def forward(ctx, x, widths, heights, derivs, inv: bool):
        lay = RationalQuadraticSpline(widths.detach(), heights.detach(), derivs.detach(), device=x.device)
        t = inverse(lay) if inv else lay
        xc = _colmajor(x.detach())
        y, lj = run_chain(t, xc)
        ctx.t = t
        ctx.save_for_backward(xc)
        return y, lj

    @staticmethod
    def backward(ctx, ybar, ljbar):
        (xc,) = ctx.saved_tensors
        D, N = xc.shape
        yb = _colmajor(ybar) if ybar is not None else _colmajor(torch.zeros((D, N), device=xc.device))
        xbar, g = rqs_vjp(ctx.t, xc, yb, ljbar.contiguous() if ljbar is not None else None)
        return xbar, g["widths"], g["heights"], g["derivatives"], None


class SplineLayer(torch.nn.Module):
    """A trainable RationalQuadraticSpline(raw widths, raw heights, raw derivatives, B) (rational_quadratic_spline.jl:
    99-123).  The constructor's normalisation (softmax -> cumsum -> [-B, B] knots, softplus derivatives with unit end
    slopes) is parameter-sized torch code, differentiated by torch as the reference's AD differentiates the constructor;
    the spline over the batch and its reverse mode run in the device kernels."""

    def __init__(self, dims: int, K: int, B: float = 3.0, device="cuda", generator=None):
        super().__init__()
        self.B = float(B)
        self.w = torch.nn.Parameter(torch.randn((dims, K), generator=generator).to(device))
        self.h = torch.nn.Parameter(torch.randn((dims, K), generator=generator).to(device))
        self.d = torch.nn.Parameter(torch.randn((dims, K - 1), generator=generator).to(device))

    def knots(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        n, dev = self.w.shape[0], self.w.device
        zero, one = torch.zeros((n, 1), device=dev), torch.ones((n, 1), device=dev)
        W = 2 * self.B * torch.cumsum(torch.cat([zero, torch.softmax(self.w, dim=1)], dim=1), dim=1) - self.B
        H = 2 * self.B * torch.cumsum(torch.cat([zero, torch.softmax(self.h, dim=1)], dim=1), dim=1) - self.B
        Dv = torch.cat([one, torch.nn.functional.softplus(self.d), one], dim=1)
        return W, H, Dv

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return _SplineFn.apply(x, *self.knots(), False)

    def inverse(self, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return _SplineFn.apply(y, *self.knots(), True)
