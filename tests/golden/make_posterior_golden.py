#!/usr/bin/env python3
"""Posterior summaries of the model under the REFERENCE'S SAMPLING ALGORITHM, restated on the CPU:
NUTS over the 17 continuous value variables + PyMC's BinaryGibbsMetropolis over i_raw / ab_s_waner
(what pm.sample assigns at abd.py:922), both driven by the oracle's NumPy restatement of the model
(oracle/abd_oracle.py: Oracle.logp_dlogp, binary_gibbs_metropolis_sweep[_local]).  No GPU code is involved.

    python tests/golden/make_posterior_golden.py [n_workers]        (about 40 minutes on 8 cores)

Writes tests/golden/posterior_goldens.npz: per cohort, the mean / sd / bulk ESS / split R-hat of the 17
scalars (constrained scale), the posterior mean of the Deterministic "i" (gap, ind) and of ab_s_waner, from
4 chains.  tests/test_gpu_sampler.py compares the GPU sampler's summaries with these within Monte-Carlo error.
Chains run one per process; the NUTS driver is abdpymc_b200.sampler.sample(kernel="nuts") on CPU tensors (host
logic only: every density evaluation and every Gibbs decision comes from the oracle).
"""
import json
import sys
import time
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

CASES = {
    # name: (cohort, splits, tune, draws)
    "test_cohort": ("test_cohort", (14, 20), 1000, 4000),
    "sim100": ("sim100", (14, 20), 1000, 2000),
}
N_CHAINS = 4


def load_cohort(name):
    from abdpymc_b200.cohort import CohortArrays, synthetic_cohort

    return synthetic_cohort(100) if name == "sim100" else CohortArrays.load(name)


class OracleTarget:
    """One chain: joint logp + gradient from Oracle.logp_dlogp, binaries by the restated BinaryGibbsMetropolis."""

    def __init__(self, cohort, splits, seed, tune):
        from oracle import abd_oracle as ora

        self.ora, self.G, self.N = ora, cohort.n_gaps, cohort.n_inds
        self.o = ora.Oracle(cohort, splits=splits, dense=False)
        self.subs = [ora.Oracle(cohort.take(np.array([n])), splits=splits, dense=False) for n in range(self.N)]
        self.i_raw = np.zeros((self.G, self.N), np.int8)
        self.w = np.zeros(self.N, np.int8)
        self.rng = np.random.default_rng(seed)
        self.sum_i = np.zeros((self.G, self.N))
        self.sum_w = np.zeros(self.N)
        self.n_rec, self.tune = 0, tune

    def logp_dlogp(self, q):
        import torch

        lp, g = self.o.logp_dlogp(q[0].numpy(), self.i_raw, self.w)
        return torch.tensor([lp], dtype=torch.float64), torch.from_numpy(np.asarray(g))[None, :]

    def gibbs(self, q, sweep):
        vals = self.ora.backward(q[0].numpy())[0]
        th = np.array([vals[n] for n in self.ora.THETA13])
        self.i_raw, self.w = self.ora.binary_gibbs_metropolis_sweep_local(
            self.subs, self.G, self.N, th, vals["p"], vals["ab_s_p_waner"], self.i_raw, self.w, self.rng)
        if sweep >= self.tune:   # post-warm-up states only
            self.sum_i += self.o.constrain(self.i_raw)
            self.sum_w += self.w
            self.n_rec += 1


def run_chain(args):
    import torch

    from abdpymc_b200.engine import forward
    from abdpymc_b200.sampler import SamplerConfig, sample

    case, chain = args
    cname, splits, tune, draws = CASES[case]
    co = load_cohort(cname)
    torch.set_num_threads(1)
    tgt = OracleTarget(co, splits, seed=1000 + chain, tune=tune)
    G = co.n_gaps
    x0 = np.array([1.0 / G, 2, 1, 10 / 11, -2, 2, 10 / 11, 0.5, 1, 1, -2, -1, 2, 1, -1, 2, 1], dtype=np.float64)
    q0 = forward(x0)[None, :] + np.random.default_rng(50 + chain).uniform(-1, 1, size=(1, 17))
    t0 = time.time()
    cfg = SamplerConfig(tune=tune, draws=draws, seed=chain, kernel="nuts", max_treedepth=8)
    res = sample(tgt, torch.from_numpy(q0), cfg)
    return case, chain, res.q[0], tgt.sum_i / tgt.n_rec, tgt.sum_w / tgt.n_rec, float(res.accept.mean()), time.time() - t0


def main():
    from abdpymc_b200 import diagnostics as dg
    from abdpymc_b200.engine import Q17_RV, backward

    workers = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    jobs = [(case, c) for case in CASES for c in range(N_CHAINS)]
    out = {}
    with ProcessPoolExecutor(workers) as ex:
        results = list(ex.map(run_chain, jobs))
    meta = {}
    for case in CASES:
        rs = sorted([r for r in results if r[0] == case], key=lambda r: r[1])
        q = np.stack([r[2] for r in rs])                      # (chains, draws, 17)
        x = backward(q)
        summ = dg.summary({name: x[:, :, k] for k, (name, _) in enumerate(Q17_RV)})
        names = [n for n, _ in Q17_RV]
        out[f"{case}/mean"] = np.array([x[:, :, k].mean() for k in range(17)])
        out[f"{case}/sd"] = np.array([x[:, :, k].std() for k in range(17)])
        out[f"{case}/ess_bulk"] = np.array([summ[n]["ess_bulk"] for n in names])
        out[f"{case}/rhat"] = np.array([summ[n]["rhat"] for n in names])
        out[f"{case}/mean_i"] = np.mean([r[3] for r in rs], axis=0)
        out[f"{case}/mean_i_by_chain"] = np.stack([r[3] for r in rs])
        out[f"{case}/mean_w"] = np.mean([r[4] for r in rs], axis=0)
        meta[case] = dict(cohort=CASES[case][0], splits=list(CASES[case][1]), tune=CASES[case][2], draws=CASES[case][3],
                          chains=N_CHAINS, accept=[r[5] for r in rs], seconds=[r[6] for r in rs], names=names,
                          algorithm="NUTS (multinomial, max_treedepth 8) + restated PyMC BinaryGibbsMetropolis (transit_p 0.8), "
                                    "densities from oracle.Oracle(dense=False)")
        print(case, "max rhat", out[f"{case}/rhat"].max(), "min ess", out[f"{case}/ess_bulk"].min(), flush=True)
    np.savez_compressed(ROOT / "tests" / "golden" / "posterior_goldens.npz", meta=np.array(json.dumps(meta)), **out)


if __name__ == "__main__":
    main()
