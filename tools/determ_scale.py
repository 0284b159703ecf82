#!/usr/bin/env python3
"""Developer tool: abd_deterministics_dev (k_determ: i, ab_n_mu, ab_s_mu for every (gap, individual) of
every chain -- a write-bound kernel, 17 G N bytes per chain) against the number of chains."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from abdpymc_b200.engine import AbdEngine  # noqa: E402

peak = bench.measured_peak_gbs()[0]
for C in [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["4", "32", "128"])]:
    co, q, vals, i_raw, w = bench.workload(n_chains=C)
    G, N = co.n_gaps, co.n_inds
    eng = AbdEngine(co, splits=bench.SPLITS)
    eng.upload_state(i_raw, w)
    di, dw = eng.state_dev(C)
    th = torch.from_numpy(vals[:, [1, 2, 3, 4, 5, 6, 10, 11, 12, 13, 14, 15, 16]].copy()).cuda()
    oi = torch.empty(C, G, N, dtype=torch.int8, device="cuda")
    mn = torch.empty(C, G, N, dtype=torch.float64, device="cuda")
    ms = torch.empty(C, G, N, dtype=torch.float64, device="cuda")
    a = (C, th.data_ptr(), di, dw, oi.data_ptr(), mn.data_ptr(), ms.data_ptr(), 0)
    for _ in range(5):
        eng.deterministics_dev(*a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        eng.deterministics_dev(*a)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    nbytes = C * (18 * G * N + N)  # 17 G N written, G N + N read
    print(f"C={C:4d}: {us:8.1f} us per launch, {nbytes / us / 1e3:8.1f} GB/s = {100 * nbytes / us / 1e3 / peak:5.1f} % of the HBM copy peak", flush=True)
    eng.close()
