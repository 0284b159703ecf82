#!/usr/bin/env python3
"""Developer tool: ESS/s of the built-in sampler against the mean trajectory length.
usage: python tools/sampler_sweep.py [n_inds] [tune] [draws] [L,L,...]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from abdpymc_b200 import diagnostics as dg  # noqa: E402
from abdpymc_b200.cohort import synthetic_cohort  # noqa: E402
from abdpymc_b200.engine import AbdEngine, forward  # noqa: E402
from abdpymc_b200.sampler import AbdTarget, SamplerConfig, sample  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
tune = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
draws = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
Ls = [int(v) for v in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["3", "6", "12"])]
jit = tuple(float(v) for v in sys.argv[5].split(",")) if len(sys.argv) > 5 else (0.6, 1.4)
seeds = [int(v) for v in sys.argv[6].split(",")] if len(sys.argv) > 6 else [1]
C = 4
co = synthetic_cohort(n)
G, N = co.n_gaps, co.n_inds
x0 = np.array([1.0 / G, 2, 1, 10 / 11, -2, 2, 10 / 11, 0.5, 1, 1, -2, -1, 2, 1, -1, 2, 1], dtype=np.float64)
for L, seed in [(L, sd) for L in Ls for sd in seeds]:
    eng = AbdEngine(co, splits=(14, 20))
    q0 = forward(x0)[None, :] + np.random.default_rng(seed).uniform(-1, 1, size=(C, 17))
    tgt = AbdTarget(eng, C, np.zeros((C, G, N), np.int8), np.zeros((C, N), np.int8), seed=seed)
    res = sample(tgt, torch.from_numpy(q0).cuda(), SamplerConfig(tune=tune, draws=draws, seed=seed, n_leapfrog=L, jitter=jit))
    summ = dg.summary(res.posterior())
    ess = sorted((v["ess_bulk"], k) for k, v in summ.items())
    print(f"L={L:2d} jitter={jit} seed={seed}: wall {res.wall_s:.2f} s, {(tune + draws) / res.wall_s:.0f} it/s, accept {res.accept.mean():.2f}, "
          f"min ESS/s {ess[0][0] / res.wall_s:.1f} ({ess[0][1]}), 2nd {ess[1][0] / res.wall_s:.1f}, 3rd {ess[2][0] / res.wall_s:.1f}, "
          f"median {ess[len(ess) // 2][0] / res.wall_s:.0f}, max rhat {max(v['rhat'] for v in summ.values()):.2f}", flush=True)
    eng.close()
