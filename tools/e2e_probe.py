#!/usr/bin/env python3
"""Developer tool: where the time of one host-pointer call goes (resident chain state, 4 chains):
raw ctypes call of abd_logp_dlogp, the engine method around it, and the device-pointer launch +
synchronise."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from abdpymc_b200.engine import AbdEngine  # noqa: E402

C = 4
co, q, vals, i_raw, w = bench.workload()
eng = AbdEngine(co, splits=bench.SPLITS)
eng.upload_state(i_raw, w)
lib, h = eng._lib, eng._h
lp, g = np.empty(C), np.empty((C, 17))
n = 3000


def timeit(f):
    for _ in range(200):
        f()
    t = time.perf_counter()
    for _ in range(n):
        f()
    return (time.perf_counter() - t) / n * 1e6


qp, lpp, gp = q.ctypes.data, lp.ctypes.data, g.ctypes.data
print(f"raw ctypes abd_logp_dlogp(q, NULL, NULL): {timeit(lambda: lib.abd_logp_dlogp(h, C, qp, None, None, lpp, gp)):6.1f} us")
print(f"engine.logp_dlogp(q):                     {timeit(lambda: eng.logp_dlogp(q)):6.1f} us")
di, dw = eng.state_dev(C)
tq = torch.from_numpy(q).cuda()
o1 = torch.zeros(C, dtype=torch.float64, device="cuda")
o2 = torch.zeros(C, 17, dtype=torch.float64, device="cuda")
a = (C, tq.data_ptr(), di, dw, o1.data_ptr(), o2.data_ptr(), 0)


def dev():
    eng.logp_dlogp_dev(*a)
    torch.cuda.synchronize()


print(f"logp_dlogp_dev + torch synchronize:        {timeit(dev):6.1f} us")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    eng.logp_dlogp_dev(*a)
e1.record()
torch.cuda.synchronize()
print(f"logp_dlogp_dev back to back (no sync):     {e0.elapsed_time(e1) / n * 1e3:6.1f} us per launch")
hi = torch.from_numpy(i_raw).pin_memory().numpy()
hw = torch.from_numpy(w).pin_memory().numpy()
print(f"engine.logp_dlogp(q, i_raw, w) pinned:     {timeit(lambda: eng.logp_dlogp(q, hi, hw)):6.1f} us")
print(f"engine.upload_state(i_raw, w) pinned:      {timeit(lambda: eng.upload_state(hi, hw)):6.1f} us")
