#!/usr/bin/env python3
"""Developer tool: where one sampler iteration's time goes (device time of each launch on its own,
back to back, against the full five-launch iteration)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from abdpymc_b200.engine import AbdEngine  # noqa: E402
from abdpymc_b200.sampler import AbdTarget  # noqa: E402

C, L, n = 4, int(sys.argv[1]) if len(sys.argv) > 1 else 5, 300
co, q, vals, i_raw, w = bench.workload()
eng = AbdEngine(co, splits=bench.SPLITS)
tgt = AbdTarget(eng, C, i_raw, w, seed=1)
dev = tgt.device
f64 = dict(dtype=torch.float64, device=dev)
tq = torch.from_numpy(q).to(dev)
logp, grad = torch.empty(C, **f64), torch.empty(C, 17, **f64)
tgt.logp_dlogp_into(tq, logp, grad)
eye = torch.eye(17, **f64)
eps = torch.full((C,), 1e-3, **f64)
da = torch.zeros(C, 4, **f64)
qw, pw, gw = torch.empty(C, 17, **f64), torch.empty(C, 17, **f64), torch.empty(C, 17, **f64)
lpw, h0, acc = torch.empty(C, **f64), torch.empty(C, **f64), torch.empty(C, **f64)
for _ in range(20):  # a few sweeps so that the Gibbs cost is the stationary one
    tgt.gibbs(tq, _)


def timed(fn, reps=n):
    for _ in range(10):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(reps):
        fn(it)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


parts = {
    "hmc_begin": lambda it: tgt.hmc_begin(tq, grad, logp, eye, it, qw, pw, gw, h0),
    f"leapfrog x{L} (persistent)": lambda it: tgt.leapfrog_inplace(qw, pw, gw, lpw, eps, eye, L),
    "hmc_end": lambda it: tgt.hmc_end(tq, grad, logp, qw, pw, gw, lpw, eye, h0, it, acc, da, eps, 0, 0.8),
    "gibbs": lambda it: tgt.gibbs(tq, 100 + it),
    "logp_dlogp": lambda it: tgt.logp_dlogp_into(tq, logp, grad),
}
def single_steps(it):
    for _ in range(L):
        tgt.leapfrog_inplace(qw, pw, gw, lpw, eps, eye, 1)


print(f"{'leapfrog as ' + str(L) + ' x 1-step launches':28s} {timed(single_steps):8.1f} us")
tot = 0.0
for name, fn in parts.items():
    t = timed(fn)
    tot += t
    print(f"{name:28s} {t:8.1f} us")


def iteration(it):
    for fn in parts.values():
        fn(it)


t_all = timed(iteration)
print(f"{'sum of the parts':28s} {tot:8.1f} us\n{'one iteration (5 launches)':28s} {t_all:8.1f} us")
