#!/usr/bin/env python3
"""Developer tool: Gibbs sweep time against the number of chains batched in one launch."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from abdpymc_b200.engine import AbdEngine  # noqa: E402

for C in [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["4", "16", "64", "128"])]:
    co, q, vals, i_raw, w = bench.workload(n_chains=C)
    eng = AbdEngine(co, splits=bench.SPLITS)
    eng.upload_state(i_raw, w)
    di, dw = eng.state_dev(C)
    tq = torch.from_numpy(q).cuda()
    for k in range(3):
        eng.gibbs_sweep_dev(C, tq.data_ptr(), 1, None, None, di, dw, 1, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for k in range(n):
        eng.gibbs_sweep_dev(C, tq.data_ptr(), 1, None, None, di, dw, 1, 3 + k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"C={C:4d}: {ms * 1e3:9.1f} us per sweep launch, {C / ms * 1e3:9.0f} chain sweeps/s", flush=True)
    eng.close()
