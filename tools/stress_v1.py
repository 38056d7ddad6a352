#!/usr/bin/env python
"""Repeat small ragged-N evaluations of the v1 kernels and compare with the v0 kernel (flakiness hunt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bijectors_jl_b200 as B
f32 = np.float32
rng = np.random.default_rng(0)
bad = 0
for D in (32, 64, 128, 256):
    lay = [B.RadialLayer(rng.standard_normal(1).astype(f32), rng.standard_normal(1).astype(f32), rng.standard_normal(D).astype(f32)),
           B.InvertibleBatchNorm(b=(rng.standard_normal(D) * .1).astype(f32), logs=(rng.standard_normal(D) * .1).astype(f32),
                                 m=(rng.standard_normal(D) * .1).astype(f32), v=rng.uniform(.5, 1.5, D).astype(f32))]
    flow = B.Composed(*lay)
    for N in (3001, 5000, 33, 100000, 517):
        x = B.from_numpy(rng.standard_normal((D, N)).astype(f32))
        B.lib().b2b_set_kernel_variant(1)
        y0, l0 = B.with_logabsdet_jacobian(flow, x)
        t0 = torch.zeros((), dtype=torch.float64, device="cuda"); B.run_chain(flow, x, want_y=False, sum_out=t0)
        B.lib().b2b_set_kernel_variant(0)
        nb = 0
        for rep in range(40):
            y1, l1 = B.with_logabsdet_jacobian(flow, x)
            t1 = torch.zeros((), dtype=torch.float64, device="cuda"); B.run_chain(flow, x, want_y=False, sum_out=t1)
            ey = float((y1 - y0).norm() / y0.norm()); el = float((l1 - l0).norm() / l0.norm()); et = abs(float(t1 - t0)) / abs(float(t0))
            if ey > 1e-5 or el > 1e-5 or et > 1e-6:
                nb += 1
                if nb <= 2: print(f"  MISMATCH D={D} N={N} rep={rep}: y {ey:.2e} lj {el:.2e} total {et:.2e}")
        print(f"D={D} N={N}: {nb}/40 mismatches")
        bad += nb
print("TOTAL MISMATCHES", bad)
