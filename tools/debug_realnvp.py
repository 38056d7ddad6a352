import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bijectors_jl_b200 as B
for nb in (2, 3):
    for pert in (False, True):
        torch.manual_seed(0)
        D, N = 64, 2048
        flow = B.autograd.RealNVP(D, nb, scale=0.5)
        if pert:
            with torch.no_grad():
                for p_ in list(flow.c) + list(flow.b) + list(flow.logs):
                    p_.add_(0.1 * torch.randn_like(p_))
        y = (torch.randn((N, D), device="cuda") * 1.3 + 0.2).t()
        x, lj = flow.inverse(y)
        loss = flow.nll(y)
        loss.backward()
        torch.cuda.synchronize()
        print(nb, pert, "fwd finite", torch.isfinite(x).all().item(), float(x.abs().max()), float(loss),
              {n: (torch.isfinite(p.grad).all().item(), float(p.grad.abs().max())) for n, p in flow.named_parameters()})
