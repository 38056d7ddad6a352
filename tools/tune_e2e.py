"""e2e (host buffers) throughput of the headline chain for several chunk sizes / stream counts of the host pipeline."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bijectors_jl_b200 as B
from bijectors_jl_b200 import interface as I

D, N, L = 128, 1 << 20, 8
rng = np.random.default_rng(0)
flow = B.Composed(*[B.PlanarLayer((rng.standard_normal(D) / np.sqrt(D)).astype(np.float32),
                                  (rng.standard_normal(D) / np.sqrt(D)).astype(np.float32),
                                  rng.standard_normal(1).astype(np.float32)) for _ in range(L)])
xh = torch.empty((N, D), dtype=torch.float32, pin_memory=True).t()
xh.copy_(torch.randn(N, D).t())
for streams in (2, 3, 4):
    for lg in (13, 14, 15, 16, 17):
        I.HOST_CHUNK_COLS, I.HOST_STREAMS = 1 << lg, streams
        for _ in range(2):
            B.with_logabsdet_jacobian(flow, xh)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            B.with_logabsdet_jacobian(flow, xh)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print(f"streams={streams} chunk=2^{lg}: {dt * 1e3:.2f} ms/step  {N / dt / 1e6:.1f} M samples/s  "
              f"({(2 * N * D * 4 + N * 4) / dt / 1e9:.1f} GB/s both directions)", flush=True)
