#!/usr/bin/env python3
"""Developer tool: the device No-U-Turn tree kernels on a standard-normal target whose leapfrog is done in torch
(isolates the tree logic from the model kernels): tree depth histogram, acceptance, moments."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from abdpymc_b200.cohort import CohortArrays  # noqa: E402
from abdpymc_b200.engine import AbdEngine  # noqa: E402

dev = torch.device("cuda:0")
f64 = dict(dtype=torch.float64, device=dev)
C, D, iters = 64, 8, 400
sd = torch.linspace(0.5, 3.0, 17, **f64)
eng = AbdEngine(CohortArrays.load("test_cohort"), splits=(14, 20))


def logp_grad(q):
    z = q / sd
    return -0.5 * (z * z).sum(dim=1), -z / sd


q = torch.randn(C, 17, **f64) * sd
lp, g = logp_grad(q)
inv_mass = torch.diag(sd * sd).contiguous()
linv_t = torch.diag(1.0 / sd).contiguous()     # inv_mass = L L^T with L = diag(sd); (L^T)^-1 = diag(1 / sd)
eps = torch.full((C,), float(sys.argv[1]) if len(sys.argv) > 1 else 0.5, **f64)
state = torch.zeros(C, eng.nuts_state_doubles(D), **f64)
eps_signed = torch.zeros(C, **f64)
any_active = torch.zeros(D + 1, dtype=torch.int32, device=dev)
qw, pw, gw = torch.zeros(C, 17, **f64), torch.zeros(C, 17, **f64), torch.zeros(C, 17, **f64)
lpw = torch.zeros(C, **f64)
acc, depth, div = torch.zeros(C, **f64), torch.zeros(C, **f64), torch.zeros(C, **f64)
da = torch.zeros(C, 4, **f64)
draws, depths, accs, divs = [], [], [], []
for it in range(iters):
    eng.nuts_begin_dev(C, D, q.data_ptr(), g.data_ptr(), lp.data_ptr(), linv_t.data_ptr(), eps.data_ptr(), 5, it, state.data_ptr(),
                       qw.data_ptr(), pw.data_ptr(), gw.data_ptr(), eps_signed.data_ptr(), any_active.data_ptr())
    for j in range(D):
        for n in range(1 << j):
            e = eps_signed[:, None]
            ph = pw + 0.5 * e * gw
            qn = qw + e * (ph @ inv_mass)
            lpn, gn = logp_grad(qn)
            qw.copy_(qn), gw.copy_(gn), pw.copy_(ph + 0.5 * e * gn), lpw.copy_(lpn)
            eng.nuts_leaf_dev(C, D, j, n, qw.data_ptr(), pw.data_ptr(), gw.data_ptr(), lpw.data_ptr(), inv_mass.data_ptr(), eps.data_ptr(),
                              5, it, state.data_ptr(), eps_signed.data_ptr(), any_active.data_ptr())
        if j + 1 < D and int(any_active[j + 1].item()) == 0:
            break
    eng.nuts_end_dev(C, D, q.data_ptr(), g.data_ptr(), lp.data_ptr(), state.data_ptr(), acc.data_ptr(), depth.data_ptr(),
                     div.data_ptr(), da.data_ptr(), eps.data_ptr(), 0, 0.8)
    torch.cuda.synchronize()
    draws.append(q.clone()), depths.append(depth.clone()), accs.append(acc.clone()), divs.append(div.clone())
x = torch.stack(draws[50:]).reshape(-1, 17)
dd = torch.stack(depths).cpu().numpy().astype(int)
print("depth histogram", np.bincount(dd.ravel(), minlength=D + 1))
print("accept", float(torch.stack(accs).mean()), "diverging", float(torch.stack(divs).mean()))
print("sd ratio", (x.std(dim=0) / sd).cpu().numpy().round(2))
print("mean / sd", (x.mean(dim=0) / sd).cpu().numpy().round(2))
lp2, g2 = logp_grad(q)
print("logp consistent", float((lp2 - lp).abs().max()), float((g2 - g).abs().max()))
