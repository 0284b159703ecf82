"""CPU tests: the oracle (oracle/abd_oracle.py) against the fixtures produced by executing the
reference's own code (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import abd_oracle as ora


def test_reference_tests_ran_under_standin(kats):
    info = kats["_reference_tests_under_standin"]
    assert info["ran"] == 45 and info["passed"] == 41
    assert all("TestModel" in s for s in info["skipped_need_pymc"])


def test_mask_multiple_infections(kats):
    for k in kats["mask_multiple"]:
        out = ora.mask_multiple_infections_chunks(np.array(k["arr"]), k["splits"])
        assert np.array_equal(out, np.array(k["out"]))


def test_incorporate_pcrpos(kats):
    for k in kats["incorporate_pcrpos"]:
        assert np.array_equal(ora.incorporate_pcrpos(k["i_raw"], k["pcrpos"]), np.array(k["out"]))


def test_mask_three_gaps(kats):
    for k in kats["mask_three_gaps"]:
        assert np.array_equal(ora.mask_three_gaps(k["arr"]), np.array(k["out"]))


def test_mask_future_infection_truth_table():
    # reference test_abd.py:14-63 (i0 = 10 passes through, :30)
    assert ora.mask_future_infection(10, 0, 0, 0) == 10
    for i0 in (0, 1):
        for taps in range(8):
            im3, im2, im1 = taps >> 2 & 1, taps >> 1 & 1, taps & 1
            assert ora.mask_future_infection(i0, im3, im2, im1) == (0 if taps else i0)


def test_constrain_infections(kats):
    for k in kats["constrain"]:
        out = ora.constrain_infections(np.array(k["i_raw"]), np.array(k["pcrpos"]), tuple(k["splits"]))
        assert np.array_equal(out, np.array(k["out"])), k["splits"]


def test_temp_response(kats):
    for k in kats["temp_response"]:
        expo = np.array(k["exposure"])
        rho = np.array(k["rho"])
        dense = ora.temp_response_dense(expo, rho)
        scan = ora.temp_response_scan(expo, rho)
        scale = k["temp"] if k["kind"] == "scalar" else 1.0  # the vector version ignores temp (abd.py:263-274)
        np.testing.assert_allclose(dense * scale, np.array(k["out"]), rtol=1e-14, atol=1e-15)
        np.testing.assert_allclose(scan * scale, np.array(k["out"]), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("dense", [True, False])
def test_joint_logp_and_grad_match_reference(goldens, cohorts, dense):
    z, cases = goldens
    for c in cases:
        if not dense and c["cohort"] == "cohort" and not c["key"].endswith("/0"):
            continue
        o = ora.Oracle(cohorts[c["cohort"]], splits=c["splits"], ignore_pcrpos=c["ignore_pcrpos"], dense=dense)
        k = c["key"]
        logp, grad = o.logp_dlogp(z[f"{k}/q"], z[f"{k}/i_raw"], z[f"{k}/w"])
        ref = float(z[f"{k}/logp"])
        assert abs(logp - ref) <= 1e-11 * abs(ref), (k, logp, ref)
        gref = z[f"{k}/grad"]
        tol = 1e-10 * np.maximum(np.abs(gref), 1e-3 * np.abs(gref).max())
        assert np.all(np.abs(grad - gref) <= tol), (k, grad - gref)


def test_c_oracle_matches_reference_goldens_and_the_numpy_oracle(goldens, cohorts):
    """oracle/abd_oracle_c.c (plain C, OpenMP; the compiled CPU port bench.py times beside the GPU) against the
    goldens produced by executing the reference's own code -- every case: all split configurations, PCR+ ignored
    or not, both bundled cohorts -- and against the NumPy oracle on a simulated 1k-individual cohort, with 1 and
    with several threads."""
    from abdpymc_b200.cohort import synthetic_cohort
    from oracle import c_oracle

    z, cases = goldens
    for c in cases:
        o = c_oracle.COracle(cohorts[c["cohort"]], splits=c["splits"], ignore_pcrpos=c["ignore_pcrpos"])
        k = c["key"]
        logp, grad = o.logp_dlogp(z[f"{k}/q"], z[f"{k}/i_raw"], z[f"{k}/w"])
        ref, gref = float(z[f"{k}/logp"]), z[f"{k}/grad"]
        assert abs(logp - ref) <= 1e-11 * abs(ref), (k, logp, ref)
        tol = 1e-10 * np.maximum(np.abs(gref), 1e-3 * np.abs(gref).max())
        assert np.all(np.abs(grad - gref) <= tol), (k, grad - gref)
    co = synthetic_cohort(1000)
    rng = np.random.default_rng(5)
    for splits in ((), (14,), (14, 20)):
        o_np = ora.Oracle(co, splits=splits, dense=False)
        for threads in (1, 3):
            o_c = c_oracle.COracle(co, splits=splits, threads=threads)
            for _ in range(3):
                q = ora.forward(ora.sample_prior(rng, co.n_gaps))
                i_raw = (rng.random((co.n_gaps, co.n_inds)) < 0.06).astype(np.int8)
                w = (rng.random(co.n_inds) < 0.5).astype(np.int8)
                a, ga = o_np.logp_dlogp(q, i_raw, w)
                b, gb = o_c.logp_dlogp(q, i_raw, w)
                assert abs(a - b) <= 1e-12 * abs(a)
                assert np.all(np.abs(ga - gb) <= 1e-10 * np.maximum(np.abs(ga), 1e-3 * np.abs(ga).max()))


def test_c_gibbs_sweep_makes_the_numpy_restatements_decisions(cohorts):
    """oracle/abd_oracle_c.c abd_c_gibbs_sweep (the compiled CPU sweep bench.py times) against
    abd_oracle.device_gibbs_sweep: same Philox streams, visiting order and accept decisions, bit for bit, for the
    Metropolis and heat-bath rules, with and without splits / PCR+, a chain index, a 64-bit seed and sweep counter
    and an individual offset; and the result does not depend on the number of threads."""
    from oracle import c_oracle

    co = cohorts["test_cohort"]
    rng = np.random.default_rng(11)
    seed, sweep = 7 + (1 << 35), 3 + (1 << 33)
    for splits, ignore in (((), False), ((14,), True), ((14, 20), False)):
        for mode in (0, 1):
            v = ora.sample_prior(rng, co.n_gaps)
            th = np.array([v[n] for n in ora.THETA13])
            i_raw = (rng.random((co.n_gaps, co.n_inds)) < 0.1).astype(np.int8)
            w = (rng.random(co.n_inds) < 0.5).astype(np.int8)
            want = ora.device_gibbs_sweep(co, splits, ignore, th, v["p"], v["ab_s_p_waner"], i_raw, w, seed, sweep, 2,
                                          mode=mode, ind_offset=1000)
            for threads in (1, 3):
                o = c_oracle.COracle(co, splits=splits, ignore_pcrpos=ignore, threads=threads)
                got = o.gibbs_sweep(th, v["p"], v["ab_s_p_waner"], i_raw, w, seed, sweep, 2, mode=mode, ind_offset=1000)
                assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and got[2] == want[2]
            assert want[2][1] > 0


def test_deterministics_match_reference(goldens, cohorts):
    z, cases = goldens
    for c in cases:
        k = c["key"]
        o = ora.Oracle(cohorts[c["cohort"]], splits=c["splits"], ignore_pcrpos=c["ignore_pcrpos"])
        vals = ora.backward(z[f"{k}/q"])[0]
        th = np.array([vals[n] for n in ora.THETA13])
        i, mu_n, mu_s = o.deterministics(th, z[f"{k}/i_raw"], z[f"{k}/w"])
        assert int(i.sum()) == int(z[f"{k}/i_sum"])
        np.testing.assert_allclose(mu_n.sum(), float(z[f"{k}/mu_n_sum"]), rtol=1e-12)
        np.testing.assert_allclose(mu_s.sum(), float(z[f"{k}/mu_s_sum"]), rtol=1e-12)
        if f"{k}/i" in z:
            assert np.array_equal(i, z[f"{k}/i"])
            np.testing.assert_allclose(mu_n, z[f"{k}/mu_n"], rtol=1e-13, atol=1e-13)
            np.testing.assert_allclose(mu_s, z[f"{k}/mu_s"], rtol=1e-13, atol=1e-13)


def test_cond_logodds_match_reference(goldens, cohorts):
    z, cases = goldens
    n = 0
    for c in cases:
        k = c["key"]
        if f"{k}/cond" not in z:
            continue
        o = ora.Oracle(cohorts[c["cohort"]], splits=c["splits"], ignore_pcrpos=c["ignore_pcrpos"])
        vals = ora.backward(z[f"{k}/q"])[0]
        th = np.array([vals[m] for m in ora.THETA13])
        lo, lo_w = o.cond_logodds(th, vals["p"], vals["ab_s_p_waner"], z[f"{k}/i_raw"], z[f"{k}/w"])
        # the reference differences two ~1e3-magnitude joint logps: absolute tolerance
        np.testing.assert_allclose(lo, z[f"{k}/cond"], rtol=1e-9, atol=2e-9)
        np.testing.assert_allclose(lo_w, z[f"{k}/cond_w"], rtol=1e-9, atol=2e-9)
        n += 1
    assert n >= 8


def test_gradient_finite_difference(cohorts):
    rng = np.random.default_rng(3)
    co = cohorts["test_cohort"]
    o = ora.Oracle(co, splits=(14, 20))
    q = ora.forward(ora.sample_prior(rng, o.G))
    i_raw = (rng.random((o.G, o.N)) < 0.06).astype(np.int8)
    w = (rng.random(o.N) < 0.5).astype(np.int8)
    _, g = o.logp_dlogp(q, i_raw, w)
    for k in range(17):
        h = 1e-6
        qp, qm = q.copy(), q.copy()
        qp[k] += h
        qm[k] -= h
        fd = (o.logp(qp, i_raw, w) - o.logp(qm, i_raw, w)) / (2 * h)
        assert abs(fd - g[k]) <= 1e-5 * max(1.0, abs(g[k])), (k, fd, g[k])


def test_gibbs_sweep_restatement_keeps_shapes(cohorts):
    rng = np.random.default_rng(0)
    co = cohorts["test_cohort"]
    o = ora.Oracle(co, splits=(14, 20), dense=False)
    vals = ora.sample_prior(rng, o.G)
    th = np.array([vals[m] for m in ora.THETA13])
    i_raw = np.zeros((o.G, o.N), np.int8)
    w = np.zeros(o.N, np.int8)
    i2, w2 = ora.binary_gibbs_metropolis_sweep(o, th, 0.04, 0.5, i_raw, w, rng)
    assert i2.shape == (o.G, o.N) and w2.shape == (o.N,)
    assert set(np.unique(i2)) <= {0, 1} and set(np.unique(w2)) <= {0, 1}


def test_local_gibbs_sweep_restatement_equals_the_global_one(cohorts):
    """binary_gibbs_metropolis_sweep_local (one individual's likelihood per proposal; what generates
    tests/golden/posterior_goldens.npz) makes the same decisions as the whole-cohort restatement for the same rng."""
    co = cohorts["test_cohort"]
    o = ora.Oracle(co, splits=(14, 20), dense=False)
    subs = [ora.Oracle(co.take(np.array([n])), splits=(14, 20), dense=False) for n in range(co.n_inds)]
    rng = np.random.default_rng(3)
    vals = ora.sample_prior(rng, o.G)
    th = np.array([vals[m] for m in ora.THETA13])
    i_raw = (rng.random((o.G, o.N)) < 0.1).astype(np.int8)
    w = (rng.random(o.N) < 0.5).astype(np.int8)
    a_i, a_w = ora.binary_gibbs_metropolis_sweep(o, th, 0.06, 0.5, i_raw, w, np.random.default_rng(9))
    b_i, b_w = ora.binary_gibbs_metropolis_sweep_local(subs, o.G, o.N, th, 0.06, 0.5, i_raw, w, np.random.default_rng(9))
    assert np.array_equal(a_i, b_i) and np.array_equal(a_w, b_w) and not np.array_equal(a_i, i_raw)


def test_restated_blocked_sweep_leaves_exact_conditional_invariant():
    """ABD_GIBBS_BLOCKED as the oracle restates it (the checker of the CUDA kernel) is a valid Gibbs
    kernel: on a G = 5 individual with two time chunks (one of them PCR+ for the second individual)
    the long-run state frequencies match the enumerated conditional posterior."""
    from abdpymc_b200.cohort import CohortArrays

    rng = np.random.default_rng(7)
    G, N, sweeps = 5, 2, 4000
    r = 8
    pcrpos = np.zeros((N, G))
    pcrpos[1, 3] = 1
    co = CohortArrays(vacs=rng.random((N, G)) < 0.2, pcrpos=pcrpos, ind=np.repeat(np.arange(N), r // N),
                      gap=rng.integers(0, G, size=r), antigen=rng.integers(0, 2, size=r),
                      x=rng.integers(0, 8, size=r).astype(float), od=rng.normal(0.9, 0.4, size=r))
    v = ora.sample_prior(rng, G)
    v.update(it_n_sigma=0.8, it_s_sigma=0.8, p=0.3, ab_s_p_waner=0.4)
    th = np.array([v[n] for n in ora.THETA13])
    splits = (2,)
    states = [(np.array([(s >> t) & 1 for t in range(G)]), (s >> G) & 1) for s in range(2 ** (G + 1))]
    exact = np.empty((N, len(states)))
    for n in range(N):
        sub = ora.Oracle(co.take(np.array([n])), splits=splits, dense=False)
        for s, (col, wn) in enumerate(states):
            k = col.sum()
            exact[n, s] = (sub.loglik(th, col.reshape(G, 1), np.array([wn])) + k * np.log(v["p"])
                           + (G - k) * np.log1p(-v["p"]) + (np.log(v["ab_s_p_waner"]) if wn else np.log1p(-v["ab_s_p_waner"])))
    exact = np.exp(exact - exact.max(axis=1, keepdims=True))
    exact /= exact.sum(axis=1, keepdims=True)
    counts = np.zeros_like(exact)
    i_raw, w = np.zeros((G, N), np.int8), np.zeros(N, np.int8)
    weights = 1 << np.arange(G)
    for sweep in range(sweeps + 20):
        i_raw, w, st = ora.device_gibbs_sweep(co, splits, False, th, v["p"], v["ab_s_p_waner"], i_raw, w, 5, sweep, 0, mode=2)
        assert st[0] == N * 3
        if sweep >= 20:
            code = (i_raw.astype(np.int64) * weights[:, None]).sum(axis=0) + (w.astype(np.int64) << G)
            for n in range(N):
                counts[n, code[n]] += 1
    tv = 0.5 * np.abs(counts / sweeps - exact).sum(axis=1)
    assert np.all(tv < 0.06), tv


def _scipy_priors(n_gaps):
    """The 17 priors as scipy.stats frozen distributions, from the reference's declarations
    (abd.py:424, 329-340, 367-388, 464-467) and PyMC's (mu, sigma) Gamma parametrisation
    alpha = mu^2 / sigma^2, beta = mu / sigma^2."""
    from scipy import stats

    def gam(mu, sigma):
        return stats.gamma(a=mu * mu / sigma**2, scale=sigma**2 / mu)

    return {
        "p": stats.beta(1, n_gaps - 1), "ab_n_perm": gam(2, 0.5), "ab_n_temp": gam(1, 0.5), "ab_n_rho": stats.beta(10, 1),
        "ab_n_init": stats.norm(-2, 1), "ab_s_perm": gam(2, 0.5), "ab_s_rho": stats.beta(10, 1),
        "ab_s_p_waner": stats.beta(1, 1), "ab_s_tempinf": gam(1, 0.5), "ab_s_tempvac": gam(1, 0.5),
        "ab_s_init": stats.norm(-2, 1), "it_n_b": stats.norm(-1, 0.5), "it_n_d": stats.norm(2, 0.5),
        "it_n_sigma": stats.expon(scale=1.0), "it_s_b": stats.norm(-1, 0.5), "it_s_d": stats.norm(2, 0.5),
        "it_s_sigma": stats.expon(scale=1.0),
    }


def test_prior_log_densities_against_scipy():
    """The restated PyMC log-densities (the part no live PyMC pins here) against an independent
    implementation: scipy.stats logpdf of the same distributions, and the gamma means / sds the
    reference declares (mu = 2 or 1, sigma = 0.5)."""
    rng = np.random.default_rng(3)
    for G in (5, 26, 31):
        pri = _scipy_priors(G)
        assert np.isclose(pri["ab_n_perm"].mean(), 2.0) and np.isclose(pri["ab_n_perm"].std(), 0.5)
        assert np.isclose(pri["ab_n_temp"].mean(), 1.0) and np.isclose(pri["ab_s_tempvac"].std(), 0.5)
        for _ in range(20):
            vals = ora.sample_prior(rng, G)
            lp, dlp = ora.prior_logp(vals, G)
            for name, dist in pri.items():
                assert np.isclose(lp[name], dist.logpdf(vals[name]), rtol=1e-12, atol=1e-12), name
                h = 1e-6 * max(1.0, abs(vals[name]))
                fd = (dist.logpdf(vals[name] + h) - dist.logpdf(vals[name] - h)) / (2 * h)
                assert np.isclose(dlp[name], fd, rtol=1e-5, atol=1e-6), name


def test_joint_logp_without_data_is_priors_plus_jacobians_plus_bernoulli():
    """On a cohort without OD rows the joint logp is the sum of the scipy prior log-densities, the
    log / logodds Jacobians and the Bernoulli terms of i_raw and ab_s_waner -- assembled here from
    scratch, independently of the oracle's own bookkeeping."""
    from abdpymc_b200.cohort import CohortArrays

    rng = np.random.default_rng(5)
    G, N = 7, 9
    empty = np.zeros(0)
    co = CohortArrays(vacs=rng.random((N, G)) < 0.1, pcrpos=np.zeros((N, G)), ind=empty.astype(int), gap=empty.astype(int),
                      antigen=empty.astype(int), x=empty, od=empty)
    o = ora.Oracle(co, splits=(3,), dense=False)
    pri = _scipy_priors(G)
    for _ in range(10):
        vals = ora.sample_prior(rng, G)
        q = ora.forward(vals)
        i_raw = (rng.random((G, N)) < 0.2).astype(np.int8)
        w = (rng.random(N) < 0.5).astype(np.int8)
        want = sum(d.logpdf(vals[k]) for k, d in pri.items())
        for (name, tr), y in zip(ora.VALUE_VARS, q):
            if tr == "log":
                want += y
            elif tr == "logodds":
                want += -np.logaddexp(0, -y) - np.logaddexp(0, y)  # log sigmoid(y) + log(1 - sigmoid(y))
        k = i_raw.sum()
        want += k * np.log(vals["p"]) + (G * N - k) * np.log1p(-vals["p"])
        kw = w.sum()
        want += kw * np.log(vals["ab_s_p_waner"]) + (N - kw) * np.log1p(-vals["ab_s_p_waner"])
        assert np.isclose(o.logp(q, i_raw, w), want, rtol=1e-12)
