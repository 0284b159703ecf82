#!/usr/bin/env python3
"""Developer tool (run on the GPU box): the CUDA Gibbs sweep against the compiled CPU restatement
(oracle/abd_oracle_c.c abd_c_gibbs_sweep) at BASELINE's full size -- 10 000 individuals x 4 chains, both
single-site rules -- bit for bit, and the CPU / GPU sweep times side by side.

    python tests/gibbs_fullsize_check.py [n_inds] [n_chains]

Kept under tests/ (it calls the oracle) but not collected by pytest yet: written after the round's GPU budget was spent, so it has not run on a GPU."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from abdpymc_b200.engine import AbdEngine  # noqa: E402
from oracle import c_oracle  # noqa: E402

n_inds = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 4
co, q, vals, i_raw, w = bench.workload(n_inds=n_inds, n_chains=C)
th = np.ascontiguousarray(vals[:, [1, 2, 3, 4, 5, 6, 10, 11, 12, 13, 14, 15, 16]])
p, pw = np.ascontiguousarray(vals[:, 0]), np.ascontiguousarray(vals[:, 7])
o = c_oracle.COracle(co, splits=bench.SPLITS)
ok = True
with AbdEngine(co, splits=bench.SPLITS) as eng:
    for mode in (0, 1):
        t0 = time.perf_counter()
        gi, gw, st = eng.gibbs_sweep(th, p, pw, i_raw, w, seed=5, sweep=2, mode=mode)
        t_gpu = time.perf_counter() - t0
        t0 = time.perf_counter()
        same = True
        for c in range(C):
            ci, cw, cst = o.gibbs_sweep(th[c], p[c], pw[c], i_raw[c], w[c], 5, 2, c, mode=mode)
            same &= bool(np.array_equal(ci, gi[c]) and np.array_equal(cw, gw[c]) and list(st[c]) == cst)
        t_cpu = time.perf_counter() - t0
        print(f"mode {mode}: identical = {same}; host-call GPU sweep of {C} chains {t_gpu * 1e3:.1f} ms, "
              f"C port ({o.threads or 'all'} threads) {t_cpu * 1e3:.0f} ms")
        ok &= same
sys.exit(0 if ok else 1)
