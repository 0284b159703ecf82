#!/usr/bin/env python3
"""Developer tool: randomised differential test of the CUDA path against the CPU oracle on many
small random cohorts (random G in 2..63, N, splits, ragged / empty row sets, integer and continuous
dilutions, PCR+ on or off): logp + gradient, conditional log-odds, Deterministics, one Gibbs sweep.
usage: python tools/fuzz_parity.py [n_cases] [seed]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from abdpymc_b200.cohort import CohortArrays  # noqa: E402
from abdpymc_b200.engine import AbdEngine  # noqa: E402
from oracle import abd_oracle as ora  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
RTOL = 1e-10
bad = 0
for case in range(n_cases):
    G = int(rng.choice([2, 3, 5, 12, 26, 31, 32, 33, 47, 63]))
    N = int(rng.choice([1, 2, 7, 33, 129, 300]))
    k = int(rng.integers(0, 3))
    splits = tuple(sorted(rng.choice(np.arange(0, G + 1), size=k, replace=False).tolist())) if k else ()
    ignore = bool(rng.random() < 0.3)
    counts = rng.poisson(rng.choice([0.5, 6, 20]), size=N) * (rng.random(N) > 0.2)
    ind = np.repeat(np.arange(N), counts)
    r = len(ind)
    perm = rng.permutation(r)
    if rng.random() < 0.5:
        x = rng.integers(0, 8, size=r).astype(float)
    else:
        x = rng.integers(0, 8, size=r) + rng.random(r) * (rng.random(r) < 0.5)
    co = CohortArrays(vacs=rng.random((N, G)) < 0.06, pcrpos=rng.random((N, G)) < 0.04, ind=ind[perm],
                      gap=rng.integers(0, G, size=r), antigen=rng.integers(0, 2, size=r), x=x, od=rng.normal(0.8, 0.5, size=r))
    C = int(rng.integers(1, 4))
    vals = [ora.sample_prior(rng, G) for _ in range(C)]
    q = np.stack([ora.forward(v) for v in vals])
    th = np.array([[v[n] for n in ora.THETA13] for v in vals])
    dens = rng.choice([0.0, 0.05, 0.3, 1.0])
    i_raw = (rng.random((C, G, N)) < dens).astype(np.int8)
    w = (rng.random((C, N)) < 0.5).astype(np.int8)
    o = ora.Oracle(co, splits=splits, ignore_pcrpos=ignore, dense=True)
    try:
        with AbdEngine(co, splits=splits, ignore_pcrpos=ignore) as eng:
            lp, g = eng.logp_dlogp(q, i_raw, w)
            eng.upload_state(i_raw, w)                    # the resident (packed) route: bitwise the same numbers
            lp_r, g_r = eng.logp_dlogp(q)
            assert np.array_equal(lp, lp_r) and np.array_equal(g, g_r), "packed resident state vs int8 route"
            ls, ln = eng.loglik_rows(th, i_raw, w)         # pointwise log-likelihood, the cohort's row order
            for c in range(C):
                ref = o.loglik_rows(th[c], i_raw[c], w[c])
                assert np.allclose(ls[c], ref["s"], rtol=1e-10, atol=1e-10) and np.allclose(ln[c], ref["n"], rtol=1e-10, atol=1e-10), "loglik rows"
            for c in range(C):
                rl, rg = o.logp_dlogp(q[c], i_raw[c], w[c])
                scale = np.maximum(np.abs(rg), 1e-3 * np.abs(rg).max())
                assert abs(lp[c] - rl) <= RTOL * abs(rl), ("logp", lp[c], rl)
                assert np.all(np.abs(g[c] - rg) <= RTOL * scale + 1e-300), ("grad", np.abs(g[c] - rg).max())
            di, mn, ms = eng.deterministics(th[0], i_raw[0], w[0])
            ri, rn, rs = o.deterministics(th[0], i_raw[0], w[0])
            assert np.array_equal(di, ri), "deterministic i"
            assert np.allclose(mn, rn, rtol=1e-12, atol=1e-12) and np.allclose(ms, rs, rtol=1e-12, atol=1e-12), "mu"
            if N * G <= 2500:
                p0, pw0 = vals[0]["p"], vals[0]["ab_s_p_waner"]
                lo, low = eng.cond_logodds(th[0], p0, pw0, i_raw[0], w[0])
                rlo, rlow = o.cond_logodds(th[0], p0, pw0, i_raw[0], w[0])
                assert np.all(np.abs(lo - rlo) <= 1e-9 * np.maximum(1, np.abs(rlo))), "cond logodds i"
                assert np.all(np.abs(low - rlow) <= 1e-9 * np.maximum(1, np.abs(rlow))), "cond logodds w"
                mode = int(rng.integers(0, 3))
                pa, pwa = np.array([v["p"] for v in vals]), np.array([v["ab_s_p_waner"] for v in vals])
                off = int(rng.integers(0, 5))
                eng.set_chain_offset(off)
                gi, gw, st = eng.gibbs_sweep(th, pa, pwa, i_raw, w, seed=case, sweep=3, mode=mode)   # all chains, one launch
                # the block draw needs time chunks and 32-bit masks; the library runs it as heat bath otherwise
                omode = 1 if (mode == 2 and (G > 31 or not splits)) else mode
                for c in range(C):
                    ri2, rw2, rst = ora.device_gibbs_sweep(co, splits, ignore, th[c], pa[c], pwa[c], i_raw[c], w[c], case, 3, off + c,
                                                           mode=omode)
                    assert np.array_equal(gi[c], ri2) and np.array_equal(gw[c], rw2) and list(st[c]) == rst, "gibbs sweep"
    except AssertionError as ex:
        bad += 1
        print(f"case {case}: G={G} N={N} splits={splits} ignore_pcrpos={ignore} rows={r} C={C} dens={dens}: MISMATCH {ex}", flush=True)
print(f"{n_cases - bad} of {n_cases} random cases agree with the oracle")
sys.exit(1 if bad else 0)
