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

    def __init__(self, D: int, mu=None, sigma=None, device="cuda", dtype=torch.float32):
        self.D = int(D)
        self.dtype = dtype
        self.mu = None if mu is None else _dev_f32(mu, device, dtype)
        self.sigma = None if sigma is None else _dev_f32(sigma, device, dtype)
        self.device = device
        for t in (self.mu, self.sigma):
            if t is not None and t.numel() != self.D:
                raise ValueError("DimensionMismatch: MvNormal parameter length")

    def __len__(self):
        return self.D

    def _terminal_desc(self):
        return _desc(_lib.MVNORMAL_DIAG, False, p0=self.mu if self.mu is not None else None,
                     p1=self.sigma if self.sigma is not None else None, _f64=self.dtype == torch.float64)

    def rand(self, n: int, seed: Optional[int] = None, offset: int = 0, column_offset: int = 0) -> torch.Tensor:
        """D×n base samples mu + sigma .* z (column-major) from the library's Philox4x32-10 + Box-Muller stream
        (b2b_randn_f32): a pure function of (seed, offset, global column, row)."""
        return _sample(self, (), n, seed, offset, column_offset, want_logjac=False)[0]


class _Identity(Transform):
    def _descs(self, inverse_, D, dtype=torch.float32):
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


def _seed(seed: Optional[int]) -> int:
    """Explicit seed, or one drawn from torch's default CPU generator (so torch.manual_seed makes sampling reproducible,
    the role `rng` plays in rand(rng, td, n))."""
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return int(seed) & 0xFFFFFFFFFFFFFFFF


def _sample(dist: MvNormal, transform, n: int, seed, offset, column_offset, want_logjac):
    from ._lib import check, lib
    from .interface import _desc_array, _stream, colmajor_empty

    import ctypes

    D = dist.D
    if isinstance(transform, tuple):
        descs = list(transform)
    else:
        descs = list(transform._descs(False, D, torch.float32))
    L = len(descs)
    arr = _desc_array(descs) if L else None
    dev = dist.device if not isinstance(dist.device, str) or dist.device != "cuda" else torch.device("cuda", torch.cuda.current_device())
    y = colmajor_empty(D, n, dev)
    lj = torch.empty((n,), dtype=torch.float32, device=y.device) if want_logjac else None
    ws_bytes = lib().b2b_chain_workspace_bytes(arr, L, D, n, 1, 0) if L else 0
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=y.device) if ws_bytes else None
    rc = lib().b2b_chain_sample_f32(
        arr, L, dist.mu.data_ptr() if dist.mu is not None else None, dist.sigma.data_ptr() if dist.sigma is not None else None,
        ctypes.c_uint64(_seed(seed)), ctypes.c_uint64(int(offset)), int(column_offset), y.data_ptr(),
        lj.data_ptr() if lj is not None else None, D, n, D, ws.data_ptr() if ws is not None else None, ws_bytes, _stream())
    check(rc, "b2b_chain_sample_f32")
    return y, lj


def rand(td, n: int, seed: Optional[int] = None, offset: int = 0, column_offset: int = 0, with_logjac: bool = False):
    """rand(rng, td, n) (transformed_distribution.jl:212-224): base samples pushed through the forward chain.  The
    reference draws z on the host and maps the transform over the columns one by one; here the normals are generated
    INSIDE the chain kernel (b2b_chain_sample_f32: Philox4x32-10 + Box-Muller), so the D×n base samples never exist in
    device memory.  `seed` plays the role of `rng` (default: drawn from torch's generator); `column_offset` lets a
    column shard continue ONE global stream (rank r passes the index of its first column).  `with_logjac=True` also
    returns log|det J| of the transform at the samples."""
    if isinstance(td, MvNormal):
        y, lj = _sample(td, (), n, seed, offset, column_offset, with_logjac)
    else:
        y, lj = _sample(td.dist, td.transform, n, seed, offset, column_offset, with_logjac)
    return (y, lj) if with_logjac else y
