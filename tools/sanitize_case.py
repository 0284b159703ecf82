#!/usr/bin/env python3
"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel once on the
test cohort and on a ragged wide-mask cohort."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from abdpymc_b200.cohort import CohortArrays  # noqa: E402
from abdpymc_b200.engine import AbdEngine, backward, Q_OF_THETA  # noqa: E402

rng = np.random.default_rng(0)


def run(co, splits, C):
    G, N = co.n_gaps, co.n_inds
    q = rng.normal(size=(C, 17)) * 0.3 + np.array([-3, .7, 0, 2.3, -2, .7, 2.3, 0, 0, 0, -2, -1, 2, -1, -1, 2, -1.0])
    x = backward(q)
    i_raw = (rng.random((C, G, N)) < 0.05).astype(np.int8)
    w = (rng.random((C, N)) < 0.5).astype(np.int8)
    with AbdEngine(co, splits=splits) as eng:
        lp, g = eng.logp_dlogp(q, i_raw, w)
        ll, g13, cnt = eng.loglik_grad(x[:, Q_OF_THETA], i_raw, w)
        eng.cond_logodds(x[:, Q_OF_THETA], x[:, 0], x[:, 7], i_raw, w)
        for mode in (0, 1):
            eng.gibbs_sweep(x[:, Q_OF_THETA], x[:, 0], x[:, 7], i_raw, w, seed=1, sweep=mode, mode=mode)
        eng.deterministics(x[:, Q_OF_THETA], i_raw, w)
        dev = torch.device("cuda:0")
        tq, tp, tg = (torch.from_numpy(v.copy()).to(dev) for v in (q, np.zeros_like(q), g))
        te = torch.full((C,), 1e-3, dtype=torch.float64, device=dev)
        tm = torch.eye(17, dtype=torch.float64, device=dev) * 1e-4
        tl = torch.zeros(C, dtype=torch.float64, device=dev)
        di, dw = eng.state_dev(C)
        eng.leapfrog_dev(C, 3, tq.data_ptr(), tp.data_ptr(), tg.data_ptr(), tl.data_ptr(), te.data_ptr(), tm.data_ptr(), di, dw, 0)
        torch.cuda.synchronize()
        eng.leapfrog_status(C)
        # round 2: block draw, pointwise log-likelihood, the cache file, the device-resident samplers (packed resident
        # state, single-step leapfrog launches with the tree bookkeeping fused, streaming Deterministic sums)
        eng.gibbs_sweep(x[:, Q_OF_THETA], x[:, 0], x[:, 7], i_raw, w, seed=1, sweep=2, mode=2)
        eng.loglik_rows(x[:, Q_OF_THETA], i_raw != 0, w != 0)
        import tempfile

        with tempfile.TemporaryDirectory() as tmp:
            eng.save_cache(tmp + "/c.bin")
            with AbdEngine.from_cache(tmp + "/c.bin", splits=splits or ()) as e2:
                lp2, _ = e2.logp_dlogp(q, i_raw, w)
                assert np.array_equal(lp, lp2)
    from abdpymc_b200.sampler import AbdTarget, SamplerConfig, sample

    for kernel in ("hmc", "nuts"):
        with AbdEngine(co, splits=splits) as eng:
            tgt = AbdTarget(eng, C, i_raw, w, seed=3)
            sample(tgt, torch.from_numpy(q).to(dev), SamplerConfig(tune=4, draws=4, seed=1, kernel=kernel, max_treedepth=3,
                                                                   thinned_deterministics=2, record_deterministics_every=1))
    print("ok", co.n_inds, co.n_gaps, lp[:2])


run(CohortArrays.load("test_cohort"), (14, 20), 2)
run(CohortArrays.load("cohort").bootstrap(700, seed=3), (14, 20), 3)
G, N = 45, 37
counts = rng.poisson(7, size=N)
ind = np.repeat(np.arange(N), counts)
r = len(ind)
run(CohortArrays(vacs=rng.random((N, G)) < 0.05, pcrpos=rng.random((N, G)) < 0.03, ind=ind, gap=rng.integers(0, G, r),
                 antigen=rng.integers(0, 2, r), x=rng.integers(0, 8, r) + 0.5 * (rng.random(r) < 0.3), od=rng.normal(0.8, 0.4, r)),
    (14, 20), 2)
