#!/usr/bin/env python3
"""Generate the committed golden fixtures by EXECUTING THE REFERENCE's own code.

Run in the build container only (needs /root/reference; the GPU box has neither it nor PyMC):

    python tests/golden/make_golden.py

Step 0  runs the reference's own unit tests (abdpymc/test_abd.py) on the NumPy stand-in
        (refshim.py) -- 41 of 45 pass; the 4 that do not are TestModel.test_indexes_*, which
        need pm.sample_prior_predictive.  This validates the stand-in's tensor primitives
        against every golden vector the reference holds for the hot path.
Step 1  reference_kats.json : known-answer tests for the integer prologue (K1-K3) and the
        temp responses (K5, K6): seeded / hand-made inputs, outputs computed by the reference
        functions abd.mask_multiple_infections*, abd.incorporate_pcrpos, abd.mask_three_gaps,
        abd.{One,Two,Three}TimeChunks.constrain_infections, abd._temp_response_*_rho.
Step 2  model_goldens.npz   : abd.model(data, splits, ignore_pcrpos) evaluated at seeded
        points on the bundled test cohort (10 x 26) and the bundled full cohort (1520 x 31):
        joint logp in PyMC's unconstrained space, per-RV log-densities, d logp / d q17 by
        complex-step differentiation of the reference code (h = 1e-30, exact to rounding),
        the three Deterministics (test cohort), and brute-force conditional log-odds of every
        binary variable (test cohort).
"""
import importlib.util
import io
import json
import sys
import unittest
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import refshim  # noqa: E402

abd = refshim.install()
at = sys.modules["pytensor.tensor"]
REF_DATA = Path("/root/reference/data")


# ----------------------------------------------------------------------------- step 0
def run_reference_tests():
    spec = importlib.util.spec_from_file_location("ref_test_abd", "/root/reference/abdpymc/test_abd.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
    res = unittest.TextTestRunner(stream=io.StringIO(), verbosity=0).run(suite)
    bad = [str(t) for t, _ in res.errors + res.failures]
    unexpected = [b for b in bad if "TestModel" not in b]
    print(f"reference test_abd.py under the stand-in: ran {res.testsRun}, not passing {len(bad)} "
          f"(expected: the 4 TestModel prior-predictive tests), unexpected {len(unexpected)}")
    assert res.testsRun == 45 and len(bad) == 4 and not unexpected, bad
    return dict(ran=res.testsRun, passed=res.testsRun - len(bad), skipped_need_pymc=sorted(bad))


# ----------------------------------------------------------------------------- step 1
def make_kats():
    rng = np.random.default_rng(7)
    kats = dict(mask_multiple=[], incorporate_pcrpos=[], mask_three_gaps=[], constrain=[], temp_response=[])

    def L(a):
        return np.asarray(a).tolist()

    for g, n, dens in [(3, 6, 0.5), (29, 50, 0.5), (8, 15, 0.5), (20, 15, 0.4), (31, 40, 0.1), (1, 4, 0.5)]:
        arr = (rng.random((g, n)) < dens).astype(int)
        for splits in [(), (min(4, g),), (min(4, g), min(12, g))]:
            if len(set(splits)) != len(splits):
                continue
            if len(splits) == 0:
                out = abd.mask_multiple_infections(at.as_tensor(arr)).eval()
            elif len(splits) == 1:
                out = abd.mask_multiple_infections_2_chunks(at.as_tensor(arr), split=splits[0]).eval()
            else:
                out = abd.mask_multiple_infections_3_chunks(at.as_tensor(arr), *splits).eval()
            kats["mask_multiple"].append(dict(arr=L(arr), splits=list(splits), out=L(out)))

    for g, n in [(7, 6), (12, 9), (31, 20)]:
        i_raw = (rng.random((g, n)) < 0.3).astype(int)
        pcr = (rng.random((g, n)) < 0.08).astype(int)
        out = abd.incorporate_pcrpos(at.as_tensor(i_raw), at.as_tensor(pcr)).eval()
        kats["incorporate_pcrpos"].append(dict(i_raw=L(i_raw), pcrpos=L(pcr), out=L(out)))

    for g, n, dens in [(5, 3, 0.9), (31, 64, 0.3), (26, 10, 0.1), (31, 32, 1.0), (2, 5, 0.5)]:
        arr = (rng.random((g, n)) < dens).astype(int)
        out = abd.mask_three_gaps(at.as_tensor(arr)).eval()
        kats["mask_three_gaps"].append(dict(arr=L(arr), out=L(out)))

    for g, n, d_raw, d_pcr in [(10, 3, 0.0, 0.2), (26, 10, 0.1, 0.03), (31, 100, 0.05, 0.02), (31, 60, 0.5, 0.1),
                               (31, 50, 1.0, 0.0), (31, 50, 0.0, 0.0), (40, 30, 0.1, 0.05), (64, 20, 0.1, 0.03)]:
        i_raw = (rng.random((g, n)) < d_raw).astype(int)
        pcr = (rng.random((g, n)) < d_pcr).astype(float)
        for splits in [(), (4,), (14,), (4, 8), (14, 20), (0, g), (g,)]:
            if splits and (splits[-1] > g):
                continue
            if len(splits) == 0:
                tc = abd.OneTimeChunk(pcrpos=at.as_tensor(pcr))
            elif len(splits) == 1:
                tc = abd.TwoTimeChunks(split=splits[0], pcrpos=at.as_tensor(pcr))
            else:
                tc = abd.ThreeTimeChunks(splits=splits, pcrpos=at.as_tensor(pcr))
            out = tc.constrain_infections(at.as_tensor(i_raw)).eval()
            kats["constrain"].append(dict(i_raw=L(i_raw), pcrpos=L(pcr.astype(int)), splits=list(splits), out=L(out)))

    for g, n in [(5, 3), (15, 11), (31, 25)]:
        expo = (rng.random((g, n)) < 0.2).astype(float)
        for rho in [0.5, 0.93, 1.0]:
            out = abd._temp_response_scalar_rho(at.as_tensor(expo), n, temp=1.7, rho=rho).eval()
            kats["temp_response"].append(dict(kind="scalar", exposure=L(expo), rho=rho, temp=1.7, out=L(out)))
        rho_v = rng.uniform(0.4, 1.0, size=n)
        out = abd._temp_response_vector_rho(at.as_tensor(expo), n, temp=1.7, rho=at.as_tensor(rho_v)).eval()
        kats["temp_response"].append(dict(kind="vector", exposure=L(expo), rho=L(rho_v), temp=1.7, out=L(out)))
    return kats


# ----------------------------------------------------------------------------- step 2
def draw_point(rng, n_gaps, mode):
    def gam(mu, sigma):
        return rng.gamma(mu * mu / sigma**2, sigma**2 / mu)

    v = {
        "p": rng.beta(1, n_gaps - 1), "ab_n_perm": gam(2, 0.5), "ab_n_temp": gam(1, 0.5),
        "ab_n_rho": rng.beta(10, 1), "ab_n_init": rng.normal(-2, 1), "ab_s_perm": gam(2, 0.5),
        "ab_s_rho": rng.beta(10, 1), "ab_s_p_waner": rng.beta(1, 1), "ab_s_tempinf": gam(1, 0.5),
        "ab_s_tempvac": gam(1, 0.5), "ab_s_init": rng.normal(-2, 1), "it_n_b": rng.normal(-1, 0.5),
        "it_n_d": rng.normal(2, 0.5), "it_n_sigma": rng.exponential(1), "it_s_b": rng.normal(-1, 0.5),
        "it_s_d": rng.normal(2, 0.5), "it_s_sigma": rng.exponential(1),
    }  # fmt: skip
    if mode == "rho_near_1":
        v["ab_n_rho"], v["ab_s_rho"] = 1 - 1e-9, 1 - 1e-12
    if mode == "small_sigma":
        v["it_n_sigma"], v["it_s_sigma"] = 0.05, 0.02
    if mode == "flat_b":
        v["it_n_b"], v["it_s_b"] = 1e-8, -1e-8
    v["p"] = min(max(v["p"], 1e-6), 1 - 1e-6)
    v["ab_s_p_waner"] = min(max(v["ab_s_p_waner"], 1e-6), 1 - 1e-6)
    q = []
    for name, tr in refshim.VALUE_VARS:
        x = v[name]
        q.append(np.log(x) if tr == "log" else np.log(x) - np.log1p(-x) if tr == "logodds" else x)
    return np.array(q)


def draw_binary(rng, g, n, mode):
    if mode == "all_zero":
        return np.zeros((g, n), np.int64), np.zeros(n, np.int64)
    if mode == "all_one":
        return np.ones((g, n), np.int64), np.ones(n, np.int64)
    dens = {"dense": 0.3, "sparse": 1.0 / g}.get(mode, 0.05)
    return (rng.random((g, n)) < dens).astype(np.int64), (rng.random(n) < 0.5).astype(np.int64)


def ref_eval(data, splits, ignore_pcrpos, q, i_raw, w):
    """The reference model at one point: (joint logp, per-RV logps, Deterministics)."""
    vals, logj = refshim.backward(q)
    point = dict(vals, i_raw=i_raw, ab_s_waner=w)
    with refshim.evaluate(point) as ev:
        abd.model(data, splits=splits, ignore_pcrpos=ignore_pcrpos)
    return ev.total + logj, ev.logp, ev.det


def ref_grad(data, splits, ignore_pcrpos, q, i_raw, w, h=1e-30):
    g = np.empty(len(q))
    for k in range(len(q)):
        qc = q.astype(complex)
        qc[k] += 1j * h
        total, _, _ = ref_eval(data, splits, ignore_pcrpos, qc, i_raw, w)
        g[k] = total.imag / h
    return g


def ref_cond_logodds(data, splits, ignore_pcrpos, q, i_raw, w):
    g, n = i_raw.shape
    out = np.empty((g, n))
    for t in range(g):
        for k in range(n):
            hi, lo = i_raw.copy(), i_raw.copy()
            hi[t, k], lo[t, k] = 1, 0
            out[t, k] = (ref_eval(data, splits, ignore_pcrpos, q, hi, w)[0]
                         - ref_eval(data, splits, ignore_pcrpos, q, lo, w)[0]).real
    out_w = np.empty(n)
    for k in range(n):
        hi, lo = w.copy(), w.copy()
        hi[k], lo[k] = 1, 0
        out_w[k] = (ref_eval(data, splits, ignore_pcrpos, q, i_raw, hi)[0]
                    - ref_eval(data, splits, ignore_pcrpos, q, i_raw, lo)[0]).real
    return out, out_w


def make_model_goldens():
    rng = np.random.default_rng(20240518)
    out = {}
    cases = []
    plans = [
        # cohort, (splits, ignore_pcrpos) configurations, point modes, extras
        ("test_cohort", REF_DATA / "test_data" / "cohort_data",
         [(None, False), ((14,), False), ((20,), False), ((14, 20), False), ((14, 20), True), (None, True)],
         ["prior", "prior", "dense", "sparse", "all_zero", "all_one", "rho_near_1", "small_sigma", "flat_b"], True),
        ("cohort", REF_DATA / "cohort_data",
         [(None, False), ((14, 20), False), ((14,), False)],
         ["prior", "prior", "sparse", "dense", "all_zero", "rho_near_1"], False),
    ]
    for cname, path, configs, modes, full in plans:
        data = abd.TiterData.from_disk(str(path))
        n, g = data.vacs.shape
        for ci, (splits, ign) in enumerate(configs):
            for pi, mode in enumerate(modes):
                q = draw_point(rng, g, mode)
                i_raw, w = draw_binary(rng, g, n, mode)
                total, terms, det = ref_eval(data, splits, ign, q, i_raw, w)
                grad = ref_grad(data, splits, ign, q, i_raw, w)
                key = f"{cname}/{ci}/{pi}"
                cases.append(dict(key=key, cohort=cname, splits=list(splits or ()), ignore_pcrpos=bool(ign), mode=mode))
                out[f"{key}/q"] = q
                out[f"{key}/i_raw"] = i_raw.astype(np.int8)
                out[f"{key}/w"] = w.astype(np.int8)
                out[f"{key}/logp"] = np.float64(total.real)
                out[f"{key}/grad"] = grad
                out[f"{key}/terms"] = np.array([np.real(terms[k]) for k in sorted(terms)])
                out[f"{key}/i_sum"] = np.int64(det["i"].sum())
                out[f"{key}/mu_n_sum"] = np.float64(det["ab_n_mu"].sum())
                out[f"{key}/mu_s_sum"] = np.float64(det["ab_s_mu"].sum())
                if full:
                    out[f"{key}/i"] = det["i"].astype(np.int8)
                    out[f"{key}/mu_n"] = det["ab_n_mu"].astype(float)
                    out[f"{key}/mu_s"] = det["ab_s_mu"].astype(float)
                    if pi in (0, 2):
                        lo, lo_w = ref_cond_logodds(data, splits, ign, q, i_raw, w)
                        out[f"{key}/cond"] = lo
                        out[f"{key}/cond_w"] = lo_w
                print(key, mode, f"logp={total.real:.6f}")
        out[f"{cname}/term_names"] = np.array(sorted(terms))
    out["cases"] = np.array(json.dumps(cases))
    return out


def make_sim_goldens():
    """Deterministic part of the reference simulator (simulation.py:222-279): with lam0 = 0 the
    only infections are the PCR+ ones, so titers depend on pcrpos / vacs alone."""
    sim = sys.modules["abdpymc.simulation"]
    out = {}
    for name, path in (("test_cohort", REF_DATA / "test_data" / "cohort_data"), ("cohort", REF_DATA / "cohort_data")):
        data = abd.TiterData.from_disk(str(path))
        n, g = data.vacs.shape
        take = np.arange(n) if n <= 10 else np.arange(0, n, 40)
        s_t, n_t, inf = [], [], []
        for k in take:
            r = sim.Individual(pcrpos=data.pcrpos[k], vacs=data.vacs[k]).infection_responses(lam0=np.zeros(g))
            s_t.append(r.s_response)
            n_t.append(r.n_response)
            inf.append(r.infections)
        out[f"sim/{name}/individuals"] = take
        out[f"sim/{name}/s_titer"] = np.array(s_t)
        out[f"sim/{name}/n_titer"] = np.array(n_t)
        out[f"sim/{name}/infections"] = np.array(inf)
    return out


if __name__ == "__main__":
    info = run_reference_tests()
    kats = make_kats()
    kats["_reference_tests_under_standin"] = info
    (HERE / "reference_kats.json").write_text(json.dumps(kats, separators=(",", ":")))
    print("wrote reference_kats.json", {k: len(v) for k, v in kats.items() if isinstance(v, list)})
    gold = make_model_goldens()
    gold.update(make_sim_goldens())
    np.savez_compressed(HERE / "model_goldens.npz", **gold)
    print("wrote model_goldens.npz", len(gold), "arrays")
