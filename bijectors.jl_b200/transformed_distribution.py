"""TransformedDistribution on the device path (src/transformed_distribution.jl).

  TransformedDistribution(dist, transform) / transformed(d, b)   :9-17, :37-38
  logpdf(td, y::Matrix) = logpdf(td.dist, x) + logjac with
      (x, logjac) = with_logabsdet_jacobian(inverse(td.transform), y)          :165-169
  rand(td, n): base samples pushed through the forward chain                   :212-224

The inverse chain, the base MvNormal log-density and (optionally) the batch sum run as ONE fused chain
launch per fusable segment: the terminal B2B_MVNORMAL_DIAG op consumes the recovered x in registers, so no
D×N intermediate is ever written.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib
from .interface import Transform, _batch_view, inverse, run_chain
from .layers import _desc, _dev_f32


class MvNormal:
    """MvNormal(μ, Diagonal(σ²)) / MvNormal(zeros(D), I) (Distributions + PDMats; third-party arithmetic
    restated in the kernel: −(D·log2π + Σ log σ²)/2 − Σ((x−μ)/σ)²/2).  `sigma` is the std-dev vector."""

    def __init__(self, D: int, mu=None, sigma=None, device="cuda"):
        self.D = int(D)
        self.mu = None if mu is None else _dev_f32(mu, device)
        self.sigma = None if sigma is None else _dev_f32(sigma, device)
        self.device = device
        for t in (self.mu, self.sigma):
            if t is not None and t.numel() != self.D:
                raise ValueError("DimensionMismatch: MvNormal parameter length")

    def __len__(self):
        return self.D

    def _terminal_desc(self):
        return _desc(_lib.MVNORMAL_DIAG, False, p0=self.mu if self.mu is not None else None,
                     p1=self.sigma if self.sigma is not None else None)

    def rand(self, n: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """D×n base samples (column-major); device Philox stream via torch (plumbing)."""
        z = torch.randn((n, self.D), device=self.device, dtype=torch.float32, generator=generator)
        if self.sigma is not None:
            z = z * self.sigma
        if self.mu is not None:
            z = z + self.mu
        return z.t()


class _Identity(Transform):
    def _descs(self, inverse_, D):
        return []


class TransformedDistribution:
    """struct TransformedDistribution{D,B} (transformed_distribution.jl:9-17)."""

    def __init__(self, dist: MvNormal, transform: Transform):
        self.dist, self.transform = dist, transform

    def __len__(self):
        return len(self.dist)


def transformed(d: MvNormal, b: Transform) -> TransformedDistribution:
    """transformed(d, b) (transformed_distribution.jl:37-38)."""
    return TransformedDistribution(d, b)


def logpdf(td, y: torch.Tensor) -> torch.Tensor:
    """logpdf(td::MvTransformed, y::Matrix) -> N-vector (transformed_distribution.jl:165-169);
    logpdf(d::MvNormal, x) for a bare base distribution."""
    if isinstance(td, MvNormal):
        return run_chain(_Identity(), y, want_y=False, extra_descs=[td._terminal_desc()])[1]
    D, _, _ = _batch_view(y)
    if D != len(td.dist):
        raise ValueError(f"DimensionMismatch: distribution has {len(td.dist)} dims, input has {D}")
    inv = inverse(td.transform)
    return run_chain(inv, y, want_y=False, extra_descs=[td.dist._terminal_desc()])[1]


def logpdf_sum(td, y: torch.Tensor, out: Optional[torch.Tensor] = None):
    """(Σ_n logpdf(td, y_n) as a device float64 scalar, logpdf vector): the training objective of
    docs/src/flows.md:74-77, reduced on the device in a fixed order."""
    if out is None:
        out = torch.zeros((), dtype=torch.float64, device=y.device if y.is_cuda else "cpu")
    if isinstance(td, MvNormal):
        t, term = _Identity(), td._terminal_desc()
    else:
        t, term = inverse(td.transform), td.dist._terminal_desc()
    _, lp = run_chain(t, y, want_y=False, extra_descs=[term], sum_out=out)
    return out, lp


def rand(td: TransformedDistribution, n: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """rand(rng, td, n) (transformed_distribution.jl:212-224): the reference maps the transform over
    columns one by one; here the whole D×n batch goes through the fused forward chain."""
    z = td.dist.rand(n, generator)
    return run_chain(td.transform, z, want_logjac=False)[0]
