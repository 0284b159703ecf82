#!/usr/bin/env python3
"""Developer tool: the built-in sampler on the 10k benchmark cohort under each Gibbs update rule
(0 = BinaryGibbsMetropolis semantics, 1 = single-site heat bath, 2 = per-chunk block draw):
iteration rate, per-parameter bulk ESS and R-hat, min / median ESS per second.
usage: tools/gibbs_modes.py [n_inds] [tune] [draws] [modes, e.g. 0,2]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from abdpymc_b200 import diagnostics as dg  # noqa: E402
from abdpymc_b200.engine import Q17_RV, AbdEngine, backward, forward  # noqa: E402
from abdpymc_b200.sampler import AbdTarget, SamplerConfig, sample  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
tune = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
draws = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
modes = [int(v) for v in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["0", "1", "2"])]
C = 4
co, q, vals, i_raw, w = bench.workload(n_inds=n)
G = co.n_gaps
dev = torch.device("cuda:0")
x0 = np.array([1.0 / G, 2, 1, 10 / 11, -2, 2, 10 / 11, 0.5, 1, 1, -2, -1, 2, 1, -1, 2, 1], dtype=np.float64)
q0 = forward(x0)[None, :] + np.random.default_rng(1).uniform(-1, 1, size=(C, 17))
for mode in modes:
    eng = AbdEngine(co, splits=bench.SPLITS)
    tgt = AbdTarget(eng, C, np.zeros_like(i_raw), np.zeros_like(w), seed=1, gibbs_mode=mode)
    res = sample(tgt, torch.from_numpy(q0).to(dev), SamplerConfig(tune=tune, draws=draws, seed=1))
    xq = backward(res.q)
    summ = dg.summary({name: xq[:, :, k] for k, (name, _) in enumerate(Q17_RV)})
    ess = sorted(v["ess_bulk"] for v in summ.values())
    print(f"mode {mode}: {(tune + draws) / res.wall_s:7.0f} it/s  min ESS/s {ess[0] / res.wall_s:8.1f}  median ESS/s "
          f"{ess[len(ess) // 2] / res.wall_s:8.1f}  max R-hat {max(v['rhat'] for v in summ.values()):.3f}  accept {res.accept.mean():.2f}")
    for name, v in summ.items():
        print(f"    {name:14s} mean {v['mean']:9.4f} sd {v['sd']:7.4f} ess {v['ess_bulk']:8.1f} rhat {v['rhat']:5.3f}")
    eng.close()
