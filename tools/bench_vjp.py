"""Times the reverse-mode (VJP) path of the headline chain: 8 x PlanarLayer, D=128, N=2^20."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bijectors_jl_b200 as B

D, N, L = int(sys.argv[1]) if len(sys.argv) > 1 else 128, 1 << 20, int(sys.argv[2]) if len(sys.argv) > 2 else 8
rng = np.random.default_rng(0)
flow = B.Composed(*[B.PlanarLayer((rng.standard_normal(D) / np.sqrt(D)).astype(np.float32),
                                  (rng.standard_normal(D) / np.sqrt(D)).astype(np.float32),
                                  rng.standard_normal(1).astype(np.float32)) for _ in range(L)])
x = B.from_numpy(rng.standard_normal((D, N)).astype(np.float32))
yb = B.from_numpy(rng.standard_normal((D, N)).astype(np.float32))
ljb = torch.randn(N, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, pg in (("xbar only", False), ("xbar + parameter cotangents", True)):
    ts = []
    for it in range(11):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        B.planar_chain_vjp(flow, x, yb, ljb, want_param_grads=pg)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts[3:]))
    # algorithmic bytes: main kernel reads x, ybar, ljbar and writes xbar + 3L scalars; the parameter pass re-reads
    # x, ybar and the scalars (twice: skinny reductions + S statistics)
    gb = 4.0 * N * (3 * D + 1 + 3 * L) / 1e9 + (4.0 * N * (2 * D + 6 * L) / 1e9 if pg else 0.0)
    print(f"{name:30s} D={D} L={L}  {t:.4f} ms  {N / t / 1e6:.2f} G samples/s  {gb / t * 1e3:.0f} GB/s "
          f"({gb / t * 1e3 / 6570.9 * 100:.1f} % of 6570.9)")

# radial chain (BASELINE C3 shape: 6 layers, D = 64)
Dr, Lr = 64, 6
rflow = B.Composed(*[B.RadialLayer(rng.standard_normal(1).astype(np.float32), rng.standard_normal(1).astype(np.float32),
                                   rng.standard_normal(Dr).astype(np.float32)) for _ in range(Lr)])
xr = B.from_numpy(rng.standard_normal((Dr, N)).astype(np.float32))
ybr = B.from_numpy(rng.standard_normal((Dr, N)).astype(np.float32))
ts = []
for it in range(11):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    B.radial_chain_vjp(rflow, xr, ybr, ljb)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = float(np.median(ts[3:]))
gb = 4.0 * N * (3 * Dr + 1) / 1e9
print(f"radial VJP (xbar + parameter cotangents) D={Dr} L={Lr}  {t:.4f} ms  {N / t / 1e6:.2f} G samples/s  {gb / t * 1e3:.0f} GB/s "
      f"({gb / t * 1e3 / 6570.9 * 100:.1f} % of 6570.9)")
