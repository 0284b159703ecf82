#!/usr/bin/env python3
"""Developer tool: run the built-in sampler on a simulated cohort and print diagnostics."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from abdpymc_b200 import diagnostics as dg  # noqa: E402
from abdpymc_b200.abd import infer_builtin  # noqa: E402
from abdpymc_b200.cohort import synthetic_cohort  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
tune = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
draws = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
chains = int(sys.argv[4]) if len(sys.argv) > 4 else 4
co = synthetic_cohort(n)
t = time.time()
res, post, last = infer_builtin(co, (14, 20), False, tune=tune, draws=draws, chains=chains, seed=1)
print(f"N={n} tune={tune} draws={draws} chains={chains}: wall {res.wall_s:.1f}s ({time.time() - t:.1f}s total), "
      f"{res.n_grad_evals} batched grad evals, accept {res.accept.mean():.2f}")
print("step sizes", np.round(res.step_size, 4), " sd(metric)", np.round(np.sqrt(np.diag(res.inv_mass)), 4))
print("logp per chain: first", np.round(res.logp[:, 0], 1), "last", np.round(res.logp[:, -1], 1))
for k, v in dg.summary(post).items():
    per_chain = np.round(post[k].mean(axis=1), 3)
    print(f"  {k:14s} mean {v['mean']:8.4f} sd {v['sd']:7.4f} ess {v['ess_bulk']:7.1f} rhat {v['rhat']:5.2f}  chains {per_chain}")
pi = res.means["i"]
truth = co.truth["infections"].T
print("P(i)|truth=1", pi[truth == 1].mean(), " P(i)|truth=0", pi[truth == 0].mean(), " total inferred/true", pi.sum(), truth.sum())
print("min ESS/s", min(v["ess_bulk"] for v in dg.summary(post).values()) / res.wall_s)
