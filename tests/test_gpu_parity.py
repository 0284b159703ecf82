"""GPU parity tests (run with ``-m gpu`` on the B200 box): the CUDA path, called through the
C ABI (abdpymc_b200.engine -> ctypes -> libabd_b200.so), against
  * the committed goldens produced by executing the reference's own code
    (tests/golden/model_goldens.npz), and
  * the CPU oracle (oracle/abd_oracle.py) on the same seeded inputs.

Tolerances (north star: <= 1e-10 relative in fp64):
  logp / loglik      |gpu - ref| <= 1e-10 * |ref|
  gradient entries   |gpu - ref| <= 1e-10 * max(|ref_k|, 1e-3 * max_j |ref_j|)
                     (an entry that is ~0 by cancellation is judged against the gradient's scale)
  conditional log-odds   |gpu - ref| <= 1e-9 * max(1, |ref|)  (both sides difference two sums)
  integers (i, counts, Gibbs states)  bit-exact
"""
import numpy as np
import pytest

from oracle import abd_oracle as ora

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def grad_ok(g, ref, rtol=RTOL):
    scale = np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max(axis=-1, keepdims=True))
    return np.all(np.abs(g - ref) <= rtol * scale + 1e-300)


@pytest.fixture(scope="module")
def Engine():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from abdpymc_b200 import build

    build.build()
    from abdpymc_b200.engine import AbdEngine

    return AbdEngine


def case_engine(Engine, cohorts, c):
    return Engine(cohorts[c["cohort"]], splits=c["splits"], ignore_pcrpos=c["ignore_pcrpos"])


def group_cases(cases):
    groups = {}
    for c in cases:
        groups.setdefault((c["cohort"], tuple(c["splits"]), c["ignore_pcrpos"]), []).append(c)
    return groups


# ------------------------------------------------------------------------------ vs reference goldens
def test_joint_logp_dlogp_vs_reference_goldens(Engine, goldens, cohorts):
    z, cases = goldens
    for (_, _, _), cs in group_cases(cases).items():
        with case_engine(Engine, cohorts, cs[0]) as eng:
            q = np.stack([z[f"{c['key']}/q"] for c in cs])
            i_raw = np.stack([z[f"{c['key']}/i_raw"] for c in cs])
            w = np.stack([z[f"{c['key']}/w"] for c in cs])
            ref = np.array([float(z[f"{c['key']}/logp"]) for c in cs])
            gref = np.stack([z[f"{c['key']}/grad"] for c in cs])
            # batched over chains ...
            lp, g = eng.logp_dlogp(q, i_raw, w)
            assert np.all(np.abs(lp - ref) <= RTOL * np.abs(ref)), (lp - ref) / ref
            assert grad_ok(g, gref), np.abs(g - gref).max()
            # ... and one chain at a time (another grid plan, so another summation order: equal
            # to rounding), and the same call twice: bitwise the same numbers
            for k in range(0, len(cs), 3):
                lp1, g1 = eng.logp_dlogp(q[k], i_raw[k], w[k])
                assert abs(lp1 - lp[k]) <= 1e-13 * abs(lp[k]) and grad_ok(g1, g[k], 1e-12)
            lp_again, g_again = eng.logp_dlogp(q, i_raw, w)
            assert np.array_equal(lp_again, lp) and np.array_equal(g_again, g)


def test_deterministics_vs_reference_goldens(Engine, goldens, cohorts):
    z, cases = goldens
    for c in cases:
        k = c["key"]
        with case_engine(Engine, cohorts, c) as eng:
            vals = ora.backward(z[f"{k}/q"])[0]
            th = np.array([vals[n] for n in ora.THETA13])
            i, mu_n, mu_s = eng.deterministics(th, z[f"{k}/i_raw"], z[f"{k}/w"])
            assert int(i.sum()) == int(z[f"{k}/i_sum"])
            np.testing.assert_allclose(mu_n.sum(), float(z[f"{k}/mu_n_sum"]), rtol=1e-12)
            np.testing.assert_allclose(mu_s.sum(), float(z[f"{k}/mu_s_sum"]), rtol=1e-12)
            if f"{k}/i" in z:
                assert np.array_equal(i, z[f"{k}/i"])
                np.testing.assert_allclose(mu_n, z[f"{k}/mu_n"], rtol=1e-12, atol=1e-12)
                np.testing.assert_allclose(mu_s, z[f"{k}/mu_s"], rtol=1e-12, atol=1e-12)


def test_cond_logodds_vs_reference_goldens(Engine, goldens, cohorts):
    z, cases = goldens
    n = 0
    for c in cases:
        k = c["key"]
        if f"{k}/cond" not in z:
            continue
        with case_engine(Engine, cohorts, c) as eng:
            vals = ora.backward(z[f"{k}/q"])[0]
            th = np.array([vals[m] for m in ora.THETA13])
            lo, lo_w = eng.cond_logodds(th, vals["p"], vals["ab_s_p_waner"], z[f"{k}/i_raw"], z[f"{k}/w"])
            ref, ref_w = z[f"{k}/cond"], z[f"{k}/cond_w"]
            assert np.all(np.abs(lo - ref) <= 3e-9 * np.maximum(1, np.abs(ref)))
            assert np.all(np.abs(lo_w - ref_w) <= 3e-9 * np.maximum(1, np.abs(ref_w)))
            n += 1
    assert n >= 8


@pytest.mark.parametrize("name,splits", [("test_cohort", (14, 20)), ("cohort", ())])
def test_pointwise_log_likelihood_vs_oracle(Engine, cohorts, name, splits):
    """abd_loglik_rows: the log-density of every OD row (the InferenceData's log_likelihood group for the reference's
    observed nodes it_s_lik / it_n_lik, abd.py:459-469), in the cohort's own row order, and its sum = the loglik."""
    from abdpymc_b200 import abd

    co = cohorts[name]
    rng = np.random.default_rng(6)
    C = 3
    q, i_raw, w = draw_points(rng, co.n_gaps, co.n_inds, C)
    vals = [ora.backward(q[k])[0] for k in range(C)]
    th = np.array([[v[n] for n in ora.THETA13] for v in vals])
    o = ora.Oracle(co, splits=splits, dense=False)
    with Engine(co, splits=splits) as eng:
        ls, ln = eng.loglik_rows(th, i_raw, w)
        ll, _, _ = eng.loglik_grad(th, i_raw, w)
        # the post-hoc helper for an InferenceData-like posterior (chain, draw, ...)
        post = {n: th[None, :, k] for k, n in enumerate(ora.THETA13)}
        post.update(i_raw=i_raw[None].astype(np.int64), ab_s_waner=w[None].astype(np.int64))
        pw = abd.pointwise_log_likelihood(eng, post)
    for c in range(C):
        ref = o.loglik_rows(th[c], i_raw[c], w[c])
        np.testing.assert_allclose(ls[c], ref["s"], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(ln[c], ref["n"], rtol=1e-10, atol=1e-10)
        assert abs(ls[c].sum() + ln[c].sum() - ll[c]) <= 1e-10 * abs(ll[c])
    assert pw["it_s_lik"].shape == (1, C, o.rows["s"]["x"].size) and np.array_equal(pw["it_s_lik"][0], ls)
    assert np.array_equal(pw["it_n_lik"][0], ln)


def test_constrained_infections_vs_reference_kats(Engine, kats):
    """K1 -> K2 -> K3 through the deterministics kernel on the reference-generated cases."""
    from abdpymc_b200.cohort import CohortArrays

    th = np.array([2.0, 1.0, 0.9, -2.0, 2.0, 0.9, -2.0, -1.0, 2.0, 1.0, -1.0, 2.0, 1.0])
    done = 0
    for k in kats["constrain"]:
        i_raw, pcr, want = np.array(k["i_raw"]), np.array(k["pcrpos"]), np.array(k["out"])
        g, n = i_raw.shape
        if g > 63:
            continue
        co = CohortArrays(vacs=np.zeros((n, g)), pcrpos=pcr.T, ind=[0], gap=[0], antigen=[0], x=[0.0], od=[0.0])
        with Engine(co, splits=tuple(k["splits"])) as eng:
            i, _, _ = eng.deterministics(th, i_raw, np.zeros(n))
        assert np.array_equal(i, want), k["splits"]
        done += 1
    assert done >= 40


# ------------------------------------------------------------------------------ vs oracle, seeded sweeps
def draw_points(rng, G, N, n_points):
    q = np.stack([ora.forward(ora.sample_prior(rng, G)) for _ in range(n_points)])
    dens = np.where(rng.random(n_points) < 0.5, 0.05, 1.0 / G)
    i_raw = (rng.random((n_points, G, N)) < dens[:, None, None]).astype(np.int8)
    w = (rng.random((n_points, N)) < 0.5).astype(np.int8)
    # adversarial points
    if n_points >= 6:
        i_raw[0] = 0
        i_raw[1] = 1
        w[2] = 0
        q[3, 3] = 18.0   # ab_n_rho -> 1
        q[3, 6] = 25.0   # ab_s_rho -> 1
        q[4, 11] = 1e-9  # it_n_b ~ 0
        q[5, 13] = np.log(0.02)  # small sigma
    return q, i_raw, w


@pytest.mark.parametrize("splits", [(), (14,), (20,), (14, 20)])
def test_parity_sweep_1k_cohort(Engine, splits):
    """configs[1]: simulated 1k-individual cohort, 256 parameter points per split config."""
    from abdpymc_b200.cohort import synthetic_cohort

    co = synthetic_cohort(1000)
    rng = np.random.default_rng(100 + len(splits) + sum(splits))
    q, i_raw, w = draw_points(rng, co.n_gaps, co.n_inds, 256)
    o = ora.Oracle(co, splits=splits, dense=False)
    ref = [o.logp_dlogp(q[k], i_raw[k], w[k]) for k in range(len(q))]
    ref_lp = np.array([r[0] for r in ref])
    ref_g = np.stack([r[1] for r in ref])
    with Engine(co, splits=splits) as eng:
        lp, g = eng.logp_dlogp(q, i_raw, w)
        assert np.all(np.abs(lp - ref_lp) <= RTOL * np.abs(ref_lp)), np.abs((lp - ref_lp) / ref_lp).max()
        assert grad_ok(g, ref_g), np.abs(g - ref_g).max()
        # the constrained-space entry point agrees with its own oracle too
        vals = np.array([[ora.backward(q[k])[0][n] for n in ora.THETA13] for k in range(8)])
        ll, g13, cnt = eng.loglik_grad(vals, i_raw[:8], w[:8])
        for k in range(8):
            rl, rg = o.loglik_grad(vals[k], i_raw[k], w[k])
            assert abs(ll[k] - rl) <= RTOL * abs(rl)
            assert grad_ok(g13[k], rg)
            assert cnt[k, 0] == i_raw[k].sum() and cnt[k, 1] == w[k].sum()
        # tile size must not change the result beyond rounding
        eng.set_tuning(300, 3)
        lp2, g2 = eng.logp_dlogp(q[:16], i_raw[:16], w[:16])
        assert np.all(np.abs(lp2 - ref_lp[:16]) <= RTOL * np.abs(ref_lp[:16]))
        assert grad_ok(g2, ref_g[:16])


def test_factored_exponential_vs_fallback_and_direct_rows(Engine, monkeypatch):
    """k_sums has three ways to obtain exp(-b (x - m)): factored (a cohort with <= 32 distinct
    dilutions: exp(-b x_j) table x one exp(b m) per cell), the general fallback (dilutions staged as
    doubles, one exp per row) and, inside the factored kernel, direct rows when |b| x is too large
    to factor.  All three against the oracle, including extreme slopes and titers."""
    from abdpymc_b200.cohort import synthetic_cohort

    co = synthetic_cohort(1000)
    rng = np.random.default_rng(77)
    q, i_raw, w = draw_points(rng, co.n_gaps, co.n_inds, 24)
    q[6, 11], q[6, 14] = -80.0, 55.0      # |b| x_max > 300: direct rows
    q[7, 11], q[7, 14] = -42.0, -42.8     # just below the switch (|b| x = 294 .. 299.6)
    q[8, 4], q[8, 10] = 90.0, -120.0      # ab_n_init / ab_s_init: b m far outside exp's range
    q[9, 11], q[9, 4] = -40.0, 30.0       # b m beyond the cap
    q[10, 14], q[10, 10] = 35.0, 25.0
    o = ora.Oracle(co, splits=(14, 20), dense=False)
    ref = [o.logp_dlogp(q[k], i_raw[k], w[k]) for k in range(len(q))]
    ref_lp, ref_g = np.array([r[0] for r in ref]), np.stack([r[1] for r in ref])
    assert np.all(np.isfinite(ref_lp)) and np.all(np.isfinite(ref_g))
    with Engine(co, splits=(14, 20)) as eng:
        lp, g = eng.logp_dlogp(q, i_raw, w)
    monkeypatch.setenv("ABD_B200_NO_FACTORED", "1")
    with Engine(co, splits=(14, 20)) as eng:
        lp_fb, g_fb = eng.logp_dlogp(q, i_raw, w)
    for a, b in ((lp, g), (lp_fb, g_fb)):
        assert np.all(np.abs(a - ref_lp) <= RTOL * np.abs(ref_lp)), np.abs((a - ref_lp) / ref_lp).max()
        assert grad_ok(b, ref_g), np.abs(b - ref_g).max()


def test_compact_cell_layout_against_the_oracle(Engine, monkeypatch):
    """Tiles too large for the 32-byte (m, T, dT, exp(b m)) cells at full occupancy take the compact layout
    (16-byte cells + an exp(b m) array; the N antigen's  sum q (x - m)  is rebuilt from sums the rows accumulate
    anyway).  (1) Forced on the 1k cohort (ABD_B200_COMPACT_CELLS=1): logp + gradient against the oracle at ordinary
    and extreme points (direct rows, capped b m), host state and resident packed state, one single-step leapfrog
    launch (the trajectory variant of the kernel), beside the ordinary layout on the same points.  (2) Chosen by the
    planner itself: 12 500 individuals x 4 chains (an eighth of the 100k cohort) run as ONE wave of larger tiles."""
    import torch

    from abdpymc_b200.cohort import synthetic_cohort

    co = synthetic_cohort(1000)
    rng = np.random.default_rng(78)
    q, i_raw, w = draw_points(rng, co.n_gaps, co.n_inds, 16)
    q[6, 11], q[6, 14] = -80.0, 55.0      # |b| x_max > 300: direct rows
    q[7, 11], q[7, 14] = -42.0, -42.8     # just below the switch
    q[8, 4], q[8, 10] = 90.0, -120.0      # b m far outside exp's range
    q[9, 11], q[9, 4] = -40.0, 30.0       # b m beyond the cap
    o = ora.Oracle(co, splits=(14, 20), dense=False)
    ref = [o.logp_dlogp(q[k], i_raw[k], w[k]) for k in range(len(q))]
    ref_lp, ref_g = np.array([r[0] for r in ref]), np.stack([r[1] for r in ref])
    with Engine(co, splits=(14, 20)) as eng:
        lp0, g0 = eng.logp_dlogp(q, i_raw, w)
        assert eng.last_plan()["compact_cells"] == 0
    monkeypatch.setenv("ABD_B200_COMPACT_CELLS", "1")
    with Engine(co, splits=(14, 20)) as eng:
        lp, g = eng.logp_dlogp(q, i_raw, w)
        plan = eng.last_plan()
        assert plan["compact_cells"] == 1, plan
        eng.upload_state(i_raw, w)
        lp_r, g_r = eng.logp_dlogp(q)
        assert np.array_equal(lp_r, lp) and np.array_equal(g_r, g)
        for a, b in ((lp0, g0), (lp, g)):
            assert np.all(np.abs(a - ref_lp) <= RTOL * np.abs(ref_lp)), np.abs((a - ref_lp) / ref_lp).max()
            assert grad_ok(b, ref_g), np.abs(b - ref_g).max()
        # one leapfrog step in one launch = half step, drift, evaluation, half step
        C = 4
        dev = torch.device("cuda:0")
        eng.upload_state(i_raw[:C], w[:C])
        a = rng.normal(size=(17, 17))
        inv_mass = 1e-4 * (a @ a.T / 17 + np.eye(17))
        eps, p = rng.uniform(0.05, 0.2, size=C), rng.normal(size=(C, 17)) * 30
        ph = p + 0.5 * eps[:, None] * g[:C]
        qn = q[:C] + eps[:, None] * (ph @ inv_mass)
        lpn, gn = eng.logp_dlogp(qn)
        pn = ph + 0.5 * eps[:, None] * gn
        tq, tp, tg = (torch.from_numpy(v.copy()).to(dev) for v in (q[:C], p, g[:C]))
        te, tm = torch.from_numpy(eps).to(dev), torch.from_numpy(inv_mass).to(dev)
        tl = torch.zeros(C, dtype=torch.float64, device=dev)
        di, dw = eng.state_dev(C)
        eng.leapfrog_dev(C, 1, tq.data_ptr(), tp.data_ptr(), tg.data_ptr(), tl.data_ptr(), te.data_ptr(), tm.data_ptr(),
                         di, dw, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        eng.leapfrog_status(C)
        assert eng.last_plan()["compact_cells"] == 1
        fin = np.isfinite(lpn)
        assert fin.sum() >= 2
        np.testing.assert_allclose(tq.cpu().numpy(), qn, rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(tl.cpu().numpy()[fin], lpn[fin], rtol=1e-11)
        np.testing.assert_allclose(tp.cpu().numpy()[fin], pn[fin], rtol=1e-9, atol=1e-7)
    monkeypatch.delenv("ABD_B200_COMPACT_CELLS")
    # (2) the planner's own choice
    co = synthetic_cohort(12_500)
    q, i_raw, w = draw_points(np.random.default_rng(79), co.n_gaps, co.n_inds, 4)
    o = ora.Oracle(co, splits=(14, 20), dense=False)
    with Engine(co, splits=(14, 20)) as eng:
        lp, g = eng.logp_dlogp(q, i_raw, w)
        plan = eng.last_plan()
        assert plan["compact_cells"] == 1 and plan["tiles"] * plan["chain_groups"] <= 148 * plan["ctas_per_sm"], plan
        for k in range(4):
            rl, rg = o.logp_dlogp(q[k], i_raw[k], w[k])
            assert abs(lp[k] - rl) <= RTOL * abs(rl)
            assert grad_ok(g[k], rg)


def random_cohort(rng, G, N, rows_per_ind=6, p_empty=0.2):
    from abdpymc_b200.cohort import CohortArrays

    counts = rng.poisson(rows_per_ind, size=N) * (rng.random(N) > p_empty)
    ind = np.repeat(np.arange(N), counts)
    r = len(ind)
    perm = rng.permutation(r)  # rows need not be sorted at the boundary
    return CohortArrays(
        vacs=rng.random((N, G)) < 0.05, pcrpos=rng.random((N, G)) < 0.03, ind=ind[perm],
        gap=rng.integers(0, G, size=r), antigen=rng.integers(0, 2, size=r),
        x=rng.integers(0, 8, size=r).astype(float) + rng.random(r) * (rng.random(r) < 0.2),
        od=rng.normal(0.8, 0.5, size=r),
    )


@pytest.mark.parametrize("G,N,splits", [(26, 300, (14, 20)), (45, 70, (20,)), (5, 3, ())])
def test_streaming_posterior_sums_match_the_per_chain_deterministics(Engine, G, N, splits):
    """abd_deterministics_accum_dev adds, in one launch and straight from q17, the sum over chains of
    the three Deterministics to running totals: two calls on two states must equal the oracle's
    Deterministics summed over all chains of both states (i exactly, the titers to 1e-12)."""
    import torch

    rng = np.random.default_rng(G + N)
    co = random_cohort(rng, G, N)
    C = 5
    o = ora.Oracle(co, splits=splits, dense=False)
    want = [np.zeros((G, N)) for _ in range(3)]
    with Engine(co, splits=splits) as eng:
        sums = [torch.zeros(G, N, dtype=torch.float64, device="cuda") for _ in range(3)]
        for rep in range(2):
            q, i_raw, w = draw_points(rng, G, N, C)
            eng.upload_state(i_raw, w)
            di, dw = eng.state_dev(C)
            tq = torch.from_numpy(q).cuda()
            eng.deterministics_accum_dev(C, tq.data_ptr(), 1, di, dw, sums[0].data_ptr(), sums[1].data_ptr(), sums[2].data_ptr())
            torch.cuda.synchronize()
            for c in range(C):
                vals = ora.backward(q[c])[0]
                th = np.array([vals[n] for n in ora.THETA13])
                for acc, v in zip(want, o.deterministics(th, i_raw[c], w[c])):
                    acc += v
        got = [t.cpu().numpy() for t in sums]
    assert np.array_equal(got[0], want[0])
    np.testing.assert_allclose(got[1], want[1], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got[2], want[2], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("G,N,splits", [(40, 70, (14, 20)), (63, 33, (30,)), (31, 1, (14, 20)), (2, 5, ()),
                                        (32, 64, (0, 32)), (12, 300, (5, 5 + 3))])
def test_ragged_and_wide_cohorts(Engine, G, N, splits):
    """Edge cases: G > 31 (64-bit masks), N = 1, G = 1, individuals without rows, unsorted rows,
    empty chunks, non-integer dilutions."""
    rng = np.random.default_rng(G * 1000 + N)
    co = random_cohort(rng, G, N)
    q, i_raw, w = draw_points(rng, G, N, 8)
    o = ora.Oracle(co, splits=splits, dense=True)
    with Engine(co, splits=splits) as eng:
        lp, g = eng.logp_dlogp(q, i_raw, w)
        for k in range(len(q)):
            rl, rg = o.logp_dlogp(q[k], i_raw[k], w[k])
            assert abs(lp[k] - rl) <= RTOL * abs(rl), (k, lp[k], rl)
            assert grad_ok(g[k], rg), (k, g[k] - rg)
        vals = ora.backward(q[6])[0]
        th = np.array([vals[n] for n in ora.THETA13])
        i, mu_n, mu_s = eng.deterministics(th, i_raw[6], w[6])
        ri, rn, rs = o.deterministics(th, i_raw[6], w[6])
        assert np.array_equal(i, ri)
        np.testing.assert_allclose(mu_n, rn, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(mu_s, rs, rtol=1e-12, atol=1e-12)
        lo, lo_w = eng.cond_logodds(th, vals["p"], vals["ab_s_p_waner"], i_raw[6], w[6])
        rlo, rlo_w = o.cond_logodds(th, vals["p"], vals["ab_s_p_waner"], i_raw[6], w[6])
        assert np.all(np.abs(lo - rlo) <= 1e-9 * np.maximum(1, np.abs(rlo)))
        assert np.all(np.abs(lo_w - rlo_w) <= 1e-9 * np.maximum(1, np.abs(rlo_w)))


def test_empty_antigen_and_no_rows(Engine):
    from abdpymc_b200.cohort import CohortArrays

    rng = np.random.default_rng(5)
    co = random_cohort(rng, 20, 12)
    only_s = CohortArrays(vacs=co.vacs, pcrpos=co.pcrpos, ind=co.ind[co.antigen == 1], gap=co.gap[co.antigen == 1],
                          antigen=co.antigen[co.antigen == 1], x=co.x[co.antigen == 1], od=co.od[co.antigen == 1])
    none = CohortArrays(vacs=co.vacs, pcrpos=co.pcrpos, ind=[], gap=[], antigen=[], x=[], od=[])
    q, i_raw, w = draw_points(rng, 20, 12, 7)
    for c2 in (only_s, none):
        o = ora.Oracle(c2, splits=(7,))
        with Engine(c2, splits=(7,)) as eng:
            lp, g = eng.logp_dlogp(q, i_raw, w)
            for k in range(len(q)):
                rl, rg = o.logp_dlogp(q[k], i_raw[k], w[k])
                assert abs(lp[k] - rl) <= RTOL * abs(rl)
                assert grad_ok(g[k], rg)


def test_error_behaviour(Engine, cohorts):
    from abdpymc_b200._lib import AbdError

    co = cohorts["test_cohort"]
    with pytest.raises(AbdError, match="ascending"):  # abd.py:613-614
        Engine(co, splits=(20, 14))
    with pytest.raises(AbdError, match="unique"):  # abd.py:619-620
        Engine(co, splits=(14, 14))
    with pytest.raises(AbdError, match="positive"):  # abd.py:611-612
        Engine(co, splits=(-1,))
    with pytest.raises(AbdError, match="largest split"):  # abd.py:615-618
        Engine(co, splits=(27,))
    with pytest.raises(NotImplementedError):  # abd.py:881-882
        Engine(co, splits=(1, 2, 3))
    with pytest.raises(ValueError, match="ints"):  # abd.py:621-622
        Engine(co, splits=(1.5,))
    with Engine(co) as eng:
        with pytest.raises(ValueError):
            eng.logp_dlogp(np.zeros(16))
        with pytest.raises(AbdError):
            eng.gibbs_sweep(np.ones(13), 0.1, 0.5, np.zeros((26, 10)), np.zeros(10), mode=7)


# ------------------------------------------------------------------------------ Gibbs
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("splits", [(), (14, 20), (9,)])
def test_gibbs_sweep_bit_exact_vs_restated_device_sweep(Engine, cohorts, mode, splits):
    """Same Philox stream, same visiting order, same accept rule on the CPU: the states after
    three consecutive sweeps must be identical bit for bit.  (Mode 2, the per-chunk block draw,
    needs time chunks: without splits the library runs it as mode 1.)"""
    omode = 1 if (mode == 2 and not splits) else mode
    co = cohorts["test_cohort"]
    rng = np.random.default_rng(11 + mode)
    C = 3
    vals = [ora.sample_prior(rng, co.n_gaps) for _ in range(C)]
    th = np.array([[v[n] for n in ora.THETA13] for v in vals])
    p = np.array([v["p"] for v in vals])
    pw = np.array([v["ab_s_p_waner"] for v in vals])
    i_raw = (rng.random((C, co.n_gaps, co.n_inds)) < 0.05).astype(np.int8)
    w = (rng.random((C, co.n_inds)) < 0.5).astype(np.int8)
    seed = 0x1234_5678_9ABC_DEF0
    with Engine(co, splits=splits) as eng:
        gi, gw = i_raw, w
        ri, rw = i_raw.copy(), w.copy()
        for sweep in range(3):
            gi, gw, st = eng.gibbs_sweep(th, p, pw, gi, gw, seed=seed, sweep=sweep, mode=mode, transit_p=0.8)
            for c in range(C):
                ri[c], rw[c], rst = ora.device_gibbs_sweep(co, splits, False, th[c], p[c], pw[c], ri[c], rw[c],
                                                           seed, sweep, c, mode=omode, transit_p=0.8)
                assert list(st[c]) == rst, (sweep, c, st[c], rst)
            assert np.array_equal(gi, ri) and np.array_equal(gw, rw), sweep
        assert not np.array_equal(gi, i_raw)  # something moved


@pytest.mark.parametrize("G,N,splits", [(26, 10, (14, 20)), (45, 9, (20,)), (31, 257, ())])
def test_packed_resident_state_matches_the_int8_path(Engine, cohorts, G, N, splits):
    """The library keeps a packed copy of the resident chain state (bit masks + constrained infections) that
    evaluations and sweeps on the resident state read instead of the int8 arrays.  Both routes must give
    bitwise identical results, the sweeps must keep the two representations in step, and a caller that
    writes into the int8 arrays itself only has to call abd_state_touch."""
    import torch

    rng = np.random.default_rng(5 + G)
    co = cohorts["test_cohort"] if (G, N) == (26, 10) else random_cohort(rng, G, N, rows_per_ind=7)
    C = 3
    q, i_raw, w = draw_points(rng, G, N, C)
    vals = [ora.backward(q[k])[0] for k in range(C)]
    th = np.array([[v[n] for n in ora.THETA13] for v in vals])
    p = np.array([v["p"] for v in vals])
    pw = np.array([v["ab_s_p_waner"] for v in vals])
    with Engine(co, splits=splits) as eng:
        lp_h, g_h = eng.logp_dlogp(q, i_raw, w)            # host state: staged for one use, int8 route
        eng.upload_state(i_raw, w)                          # resident: packed route
        lp_r, g_r = eng.logp_dlogp(q)
        assert np.array_equal(lp_h, lp_r) and np.array_equal(g_h, g_r)
        for mode in (0, 1, 2):
            # in place on host arrays (int8 route) against upload + resident sweep + download (packed route)
            a_i, a_w = i_raw.copy(), w.copy()
            _, _, st_a = eng.gibbs_sweep(th, p, pw, a_i, a_w, seed=3, sweep=mode, mode=mode, inplace=True)
            b_i, b_w, st_b = eng.gibbs_sweep(th, p, pw, i_raw, w, seed=3, sweep=mode, mode=mode)
            assert np.array_equal(a_i, b_i) and np.array_equal(a_w, b_w) and np.array_equal(st_a, st_b)
            # the packed copy followed the sweep: evaluating the resident state == evaluating what was downloaded
            lp_res, g_res = eng.logp_dlogp(q)
            lp_dl, g_dl = eng.logp_dlogp(q, b_i, b_w)
            assert np.array_equal(lp_res, lp_dl) and np.array_equal(g_res, g_dl)
            assert not np.array_equal(b_i, i_raw)
        # a caller writing into the resident int8 arrays itself: stale until abd_state_touch
        eng.upload_state(i_raw, w)
        d_i, d_w = eng.state_dev(C)
        new_i = (rng.random((C, G, N)) < 0.1).astype(np.int8)
        ti = torch.from_numpy(new_i).cuda()
        import ctypes

        cudart = ctypes.CDLL("libcudart.so.12")  # already loaded by torch
        cudart.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        assert cudart.cudaMemcpy(d_i, ti.data_ptr(), new_i.nbytes, 3) == 0  # device to device
        torch.cuda.synchronize()
        eng.state_touch()
        lp_t, g_t = eng.logp_dlogp(q)
        lp_n, g_n = eng.logp_dlogp(q, new_i, w)
        assert np.array_equal(lp_t, lp_n) and np.array_equal(g_t, g_n)


@pytest.mark.parametrize("name,splits,ignore", [("cohort", (14, 20), False), ("test_cohort", (), True), ("wide_random", (20,), False)])
def test_cache_file_round_trip(Engine, cohorts, tmp_path, name, splits, ignore):
    """abd_save_cache / abd_create_from_cache: the engine read back from the file is the same engine
    (bitwise the same logp / gradient, the same Gibbs sweep), and it knows what it was built with.
    ("wide_random": 64-bit masks and continuous dilutions, i.e. the kernel variant without the factored exponential.)"""
    rng = np.random.default_rng(2)
    if name == "wide_random":
        co = random_cohort(rng, 45, 70, rows_per_ind=9)
        co.x[:] = co.x + rng.random(co.x.size)
    else:
        co = cohorts[name]
    q, i_raw, w = draw_points(rng, co.n_gaps, co.n_inds, 2)
    path = tmp_path / "c.abdcache"
    with Engine(co, splits=splits, ignore_pcrpos=ignore) as eng:
        eng.save_cache(path)
        lp, g = eng.logp_dlogp(q, i_raw, w)
        vals = [ora.backward(q[k])[0] for k in range(2)]
        th = np.array([[v[n] for n in ora.THETA13] for v in vals])
        p, pw = np.array([v["p"] for v in vals]), np.array([v["ab_s_p_waner"] for v in vals])
        gi, gw, st = eng.gibbs_sweep(th, p, pw, i_raw, w, seed=4, sweep=2)
    with Engine.from_cache(path) as e2:
        assert (e2.G, e2.N, e2.R_s + e2.R_n) == (co.n_gaps, co.n_inds, co.n_rows)
        assert e2.splits == tuple(splits) and e2.ignore_pcrpos == ignore
        lp2, g2 = e2.logp_dlogp(q, i_raw, w)
        gi2, gw2, st2 = e2.gibbs_sweep(th, p, pw, i_raw, w, seed=4, sweep=2)
    assert np.array_equal(lp, lp2) and np.array_equal(g, g2)
    assert np.array_equal(gi, gi2) and np.array_equal(gw, gw2) and np.array_equal(st, st2)
    with pytest.raises(ValueError):
        Engine.from_cache(path, splits=(3,))
    path.write_bytes(path.read_bytes()[:200])
    with pytest.raises(Exception, match="cache"):
        Engine.from_cache(path)


def test_chain_offset_keys_the_rng_streams(Engine, cohorts):
    """abd_set_chain_offset: chain c of an engine whose first chain is global chain k draws from the stream of
    global chain k + c (what processes that shard the chains of one run rely on)."""
    co = cohorts["test_cohort"]
    rng = np.random.default_rng(21)
    C, splits = 2, (14, 20)
    vals = [ora.sample_prior(rng, co.n_gaps) for _ in range(C)]
    th = np.array([[v[n] for n in ora.THETA13] for v in vals])
    p = np.array([v["p"] for v in vals])
    pw = np.array([v["ab_s_p_waner"] for v in vals])
    i_raw = (rng.random((C, co.n_gaps, co.n_inds)) < 0.05).astype(np.int8)
    w = (rng.random((C, co.n_inds)) < 0.5).astype(np.int8)
    with Engine(co, splits=splits) as eng:
        base_i, base_w, _ = eng.gibbs_sweep(th, p, pw, i_raw, w, seed=9, sweep=4)
        eng.set_chain_offset(5)
        off_i, off_w, st = eng.gibbs_sweep(th, p, pw, i_raw, w, seed=9, sweep=4)
        assert not np.array_equal(base_i, off_i)
        for c in range(C):
            ri, rw, rst = ora.device_gibbs_sweep(co, splits, False, th[c], p[c], pw[c], i_raw[c], w[c], 9, 4, 5 + c)
            assert np.array_equal(off_i[c], ri) and np.array_equal(off_w[c], rw) and list(st[c]) == rst
        with pytest.raises(Exception):
            eng.set_chain_offset(-1)


def test_gibbs_sweep_wide_mask_bit_exact(Engine):
    rng = np.random.default_rng(77)
    co = random_cohort(rng, 45, 9, rows_per_ind=10, p_empty=0.0)
    v = ora.sample_prior(rng, 45)
    th = np.array([v[n] for n in ora.THETA13])
    i_raw = (rng.random((45, 9)) < 0.05).astype(np.int8)
    w = (rng.random(9) < 0.5).astype(np.int8)
    with Engine(co, splits=(14, 20)) as eng:
        gi, gw, st = eng.gibbs_sweep(th, v["p"], v["ab_s_p_waner"], i_raw, w, seed=5, sweep=9, mode=0)
    ri, rw, rst = ora.device_gibbs_sweep(co, (14, 20), False, th, v["p"], v["ab_s_p_waner"], i_raw, w, 5, 9, 0)
    assert np.array_equal(gi, ri) and np.array_equal(gw, rw) and list(st) == rst


@pytest.mark.parametrize("with_pcr", [False, True])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_gibbs_stationary_distribution_small_model(Engine, mode, with_pcr):
    """Statistical parity: on a G = 5 cohort every individual has 2^6 binary states, so the exact
    conditional posterior given theta is enumerable with the oracle.  Both update rules must
    leave it invariant: compare empirical state frequencies over many sweeps x chains."""
    from abdpymc_b200.cohort import CohortArrays

    rng = np.random.default_rng(2024)
    G, N, C, sweeps = 5, 6, 256, 400
    co = random_cohort(rng, G, N, rows_per_ind=5, p_empty=0.0)
    pcrpos = np.zeros((N, G))
    if with_pcr:  # a PCR+ month overrides its whole chunk: the chunk's raw bits are then prior-only
        pcrpos[0, 1] = pcrpos[1, 3] = pcrpos[2, 0] = pcrpos[2, 4] = 1
    co = CohortArrays(vacs=co.vacs, pcrpos=pcrpos, ind=co.ind, gap=co.gap, antigen=co.antigen, x=co.x,
                      od=rng.normal(0.9, 0.4, size=co.n_rows))
    v = ora.sample_prior(rng, G)
    v.update(it_n_sigma=0.8, it_s_sigma=0.8, p=0.3, ab_s_p_waner=0.4)
    th = np.array([v[n] for n in ora.THETA13])
    splits = (2,)
    # exact distribution per individual
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
    thC = np.tile(th, (C, 1))
    pC, pwC = np.full(C, v["p"]), np.full(C, v["ab_s_p_waner"])
    weights = 1 << np.arange(G)
    with Engine(co, splits=splits) as eng:
        eng.upload_state(np.zeros((C, G, N), np.int8), np.zeros((C, N), np.int8))
        for sweep in range(sweeps + 50):
            gi, gw, _ = eng.gibbs_sweep(thC, pC, pwC, seed=99, sweep=sweep, mode=mode, transit_p=0.8)
            if sweep < 50:
                continue
            code = (gi.astype(np.int64) * weights[None, :, None]).sum(axis=1) + (gw.astype(np.int64) << G)  # (C, N)
            for n in range(N):
                counts[n] += np.bincount(code[:, n], minlength=exact.shape[1])
    freq = counts / counts.sum(axis=1, keepdims=True)
    # total-variation distance per individual; sampling noise at 102 400 (correlated) draws is ~1e-2
    tv = 0.5 * np.abs(freq - exact).sum(axis=1)
    assert np.all(tv < 0.05), tv


# ------------------------------------------------------------------------------ size-independent properties
def test_full_size_properties_10k(Engine):
    """BASELINE config sizes (10k individuals): additivity over shards of individuals (the
    multi-GPU decomposition), invariance to the order of individuals, chain-batch consistency,
    and the identity  logp_dlogp == finalize(sums)  through the device-pointer API."""
    import torch

    from abdpymc_b200.cohort import shard_bounds, synthetic_cohort

    co = synthetic_cohort(10_000)
    rng = np.random.default_rng(8)
    C = 4
    q, i_raw, w = draw_points(rng, co.n_gaps, co.n_inds, C)
    splits = (14, 20)
    with Engine(co, splits=splits) as eng:
        lp, g = eng.logp_dlogp(q, i_raw, w)
        # oracle on the first chain (recurrence form, a few seconds)
        o = ora.Oracle(co, splits=splits, dense=False)
        rl, rg = o.logp_dlogp(q[0], i_raw[0], w[0])
        assert abs(lp[0] - rl) <= RTOL * abs(rl)
        assert grad_ok(g[0], rg)

        # device-pointer API: sums -> finalize equals the fused single launch
        dev = torch.device("cuda:0")
        tq = torch.from_numpy(q).to(dev)
        ti = torch.from_numpy(i_raw).to(dev)
        tw = torch.from_numpy(w).to(dev)
        sums = torch.zeros(C, 16, dtype=torch.float64, device=dev)
        out = torch.zeros(C, dtype=torch.float64, device=dev)
        outg = torch.zeros(C, 17, dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        eng.sums_dev(C, tq.data_ptr(), 1, ti.data_ptr(), tw.data_ptr(), sums.data_ptr(), st)
        eng.finalize_logp_dev(C, tq.data_ptr(), sums.data_ptr(), out.data_ptr(), outg.data_ptr(), st)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), lp) and np.array_equal(outg.cpu().numpy(), g)
        whole = sums.cpu().numpy()

    # shards of individuals: raw sums add up to the whole cohort's
    totals = (co.n_inds, int((co.antigen == 1).sum()), int((co.antigen == 0).sum()))
    acc = np.zeros_like(whole)
    world = 3
    for r in range(world):
        lo, hi = shard_bounds(co.n_inds, r, world)
        with Engine(co.shard(r, world), splits=splits, totals=totals, ind_offset=lo) as eng:
            tis = torch.from_numpy(np.ascontiguousarray(i_raw[:, :, lo:hi])).to(dev)
            tws = torch.from_numpy(np.ascontiguousarray(w[:, lo:hi])).to(dev)
            s = torch.zeros(C, 16, dtype=torch.float64, device=dev)
            eng.sums_dev(C, tq.data_ptr(), 1, tis.data_ptr(), tws.data_ptr(), s.data_ptr(), st)
            torch.cuda.synchronize()
            acc += s.cpu().numpy()
            if r == world - 1:
                # finalising the all-reduced sums on any shard gives the whole-cohort answer
                ts = torch.from_numpy(acc).to(dev)
                eng.finalize_logp_dev(C, tq.data_ptr(), ts.data_ptr(), out.data_ptr(), outg.data_ptr(), st)
                torch.cuda.synchronize()
                assert np.all(np.abs(out.cpu().numpy() - lp) <= 1e-12 * np.abs(lp))
                assert grad_ok(outg.cpu().numpy(), g, 1e-11)
    np.testing.assert_allclose(acc, whole, rtol=1e-12, atol=1e-9)

    # permuting individuals permutes nothing in the answer
    perm = rng.permutation(co.n_inds)
    with Engine(co.take(perm), splits=splits) as eng:
        lp_p, g_p = eng.logp_dlogp(q, i_raw[:, :, perm], w[:, perm])
    assert np.all(np.abs(lp_p - lp) <= 1e-12 * np.abs(lp))
    assert grad_ok(g_p, g, 1e-11)


def test_gibbs_properties_10k(Engine):
    """At 10k individuals: a sweep only ever produces 0/1 states, is reproducible for a given
    (seed, sweep), differs across sweeps, is independent of chain batching and of sharding
    (ind_offset), and its flip statistics are plausible."""
    from abdpymc_b200.cohort import synthetic_cohort

    co = synthetic_cohort(10_000)
    rng = np.random.default_rng(9)
    C = 2
    vals = [ora.sample_prior(rng, co.n_gaps) for _ in range(C)]
    th = np.array([[v[n] for n in ora.THETA13] for v in vals])
    p = np.array([0.04, 0.03])
    pw = np.array([0.5, 0.6])
    i_raw = (rng.random((C, co.n_gaps, co.n_inds)) < 0.04).astype(np.int8)
    w = (rng.random((C, co.n_inds)) < 0.5).astype(np.int8)
    with Engine(co, splits=(14, 20)) as eng:
        a_i, a_w, st = eng.gibbs_sweep(th, p, pw, i_raw, w, seed=1, sweep=0)
        b_i, b_w, _ = eng.gibbs_sweep(th, p, pw, i_raw, w, seed=1, sweep=0)
        c_i, c_w, _ = eng.gibbs_sweep(th, p, pw, i_raw, w, seed=1, sweep=1)
        assert set(np.unique(a_i)) <= {0, 1} and set(np.unique(a_w)) <= {0, 1}
        assert np.array_equal(a_i, b_i) and np.array_equal(a_w, b_w)
        assert not np.array_equal(a_i, c_i)
        n_bits = co.n_gaps * co.n_inds + co.n_inds
        assert np.all(np.abs(st[:, 0] / n_bits - 0.8) < 0.01)  # transit_p
        assert np.all(st[:, 1] > 0) and np.all(st[:, 1] < st[:, 0])
        # chain 1 alone, as chain index 0, differs (streams are keyed by chain) but a shard with
        # the right offset reproduces the whole-cohort update of its individuals
    lo, hi = 2500, 5000
    with Engine(co.take(np.arange(lo, hi)), splits=(14, 20), ind_offset=lo) as eng:
        s_i, s_w, _ = eng.gibbs_sweep(th, p, pw, i_raw[:, :, lo:hi], w[:, lo:hi], seed=1, sweep=0)
    assert np.array_equal(s_i, a_i[:, :, lo:hi]) and np.array_equal(s_w, a_w[:, lo:hi])


def test_small_cohort_between_evaluations_of_a_large_one(Engine, cohorts):
    """The kernels' dynamic shared-memory limit is a per-process function attribute: creating and
    using an engine for a small cohort must not break an engine of a large one (regression)."""
    from abdpymc_b200.cohort import synthetic_cohort

    big = synthetic_cohort(10_000)
    rng = np.random.default_rng(12)
    q, i_raw, w = draw_points(rng, big.n_gaps, big.n_inds, 2)
    qs, i_s, w_s = draw_points(rng, 26, 10, 2)
    with Engine(big, splits=(14, 20)) as eb:
        lp0, g0 = eb.logp_dlogp(q, i_raw, w)
        with Engine(cohorts["test_cohort"], splits=(14, 20)) as es:
            es.logp_dlogp(qs, i_s, w_s)
            es.set_tuning(100, 1)
            es.logp_dlogp(qs, i_s, w_s)
        lp1, g1 = eb.logp_dlogp(q, i_raw, w)
        eb.set_tuning(1100, 2)
        eb.logp_dlogp(q, i_raw, w)
        eb.set_tuning(0, 0)
        lp2, g2 = eb.logp_dlogp(q, i_raw, w)
    assert np.array_equal(lp0, lp1) and np.array_equal(g0, g1) and np.array_equal(lp0, lp2) and np.array_equal(g0, g2)


def test_pinned_state_is_pulled_by_the_sms(Engine):
    """Chain state handed over in pinned host memory is fetched by a kernel (16-byte loads over
    PCIe) instead of the copy engine; pageable memory takes cudaMemcpyAsync.  Same bytes on the
    device either way (sizes that are not multiples of 16 included)."""
    import torch

    from abdpymc_b200.cohort import synthetic_cohort

    co = synthetic_cohort(1000)
    rng = np.random.default_rng(3)
    for C in (3, 4):
        q, i_raw, w = draw_points(rng, co.n_gaps, co.n_inds, C)
        assert (C * co.n_gaps * co.n_inds) % 16 == (8 if C == 3 else 0)
        pi, pw = torch.from_numpy(i_raw).pin_memory(), torch.from_numpy(w).pin_memory()
        with Engine(co, splits=(14, 20)) as eng:
            lp_a, g_a = eng.logp_dlogp(q, i_raw, w)                  # pageable
            back_i, back_w = eng.download_state(C)
            assert np.array_equal(back_i, i_raw) and np.array_equal(back_w, w)
            eng.upload_state(np.zeros_like(i_raw), np.zeros_like(w))
            lp_b, g_b = eng.logp_dlogp(q, pi.numpy(), pw.numpy())    # pinned: pulled
            back_i, back_w = eng.download_state(C)
            assert np.array_equal(back_i, i_raw) and np.array_equal(back_w, w)
        assert np.array_equal(lp_a, lp_b) and np.array_equal(g_a, g_b)
        # the in/out Gibbs call: pinned arrays are pulled, swept and pushed back in place
        vals = np.array([[ora.backward(q[k])[0][n] for n in ora.THETA13] for k in range(C)])
        p = np.array([ora.backward(q[k])[0]["p"] for k in range(C)])
        p_w = np.array([ora.backward(q[k])[0]["ab_s_p_waner"] for k in range(C)])
        with Engine(co, splits=(14, 20)) as eng:
            ri, rw, rst = eng.gibbs_sweep(vals, p, p_w, i_raw, w, seed=5, sweep=2)
            pi2, pw2 = pi.clone(), pw.clone()
            _, _, st = eng.gibbs_sweep(vals, p, p_w, pi2.numpy(), pw2.numpy(), seed=5, sweep=2, inplace=True)
            pg_i, pg_w = i_raw.copy(), w.copy()
            eng.gibbs_sweep(vals, p, p_w, pg_i, pg_w, seed=5, sweep=2, inplace=True)   # pageable, in place
        assert np.array_equal(pi2.numpy(), ri) and np.array_equal(pw2.numpy(), rw) and np.array_equal(st, rst)
        assert np.array_equal(pg_i, ri) and np.array_equal(pg_w, rw) and not np.array_equal(ri, i_raw)


def test_fast_math(Engine):
    """The kernels' table-based exp and Newton reciprocal against libm: <= 2 ulp over the whole
    range the OD-row code can produce (z is capped at 700 by the caller)."""
    import ctypes as C

    from abdpymc_b200 import _lib

    rng = np.random.default_rng(1)
    z = np.concatenate([rng.uniform(-745, 700, 400_000), rng.uniform(-40, 40, 400_000), rng.normal(0, 1e-3, 50_000),
                        np.array([0.0, -0.0, 700.0, -708.0, -745.0, -800.0, -1e6, -1e300, np.nan, 1e-300, -1e-300])])
    e = np.empty_like(z)
    r = np.empty_like(z)
    _lib.check(_lib.load().abd_debug_fast_math(0, len(z), z.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p),
                                               r.ctypes.data_as(C.c_void_p)))
    ref = np.exp(z)
    ok = np.isfinite(z) & (z > -700)
    assert np.max(np.abs(e[ok] - ref[ok]) / ref[ok]) < 4.5e-16
    assert np.all(e[np.isfinite(z) & (z <= -700)] <= 1e-300)  # flushed towards 0, never negative / NaN
    assert np.all(e[np.isfinite(z)] >= 0) and np.isnan(e[np.isnan(z)]).all()
    x = 1.0 + np.abs(z[np.isfinite(z)])
    assert np.max(np.abs(r[np.isfinite(z)] * x - 1.0)) < 4.5e-16


@pytest.mark.parametrize("n_inds,C,L", [(10_000, 4, 7), (1000, 2, 1), (1000, 3, 12), (60_000, 3, 1)])  # (the last: several waves, single step)
def test_persistent_leapfrog_matches_stepwise(Engine, n_inds, C, L):
    """abd_leapfrog_dev (one persistent launch, CTAs hand the position over through a generation
    counter) against the same L leapfrog steps done with one abd_logp_dlogp call per step."""
    import torch

    from abdpymc_b200.cohort import synthetic_cohort

    co = synthetic_cohort(n_inds)
    rng = np.random.default_rng(4 + L)
    q = np.stack([ora.forward(ora.sample_prior(rng, co.n_gaps)) for _ in range(C)])
    i_raw = (rng.random((C, co.n_gaps, co.n_inds)) < 0.04).astype(np.int8)
    w = (rng.random((C, co.n_inds)) < 0.5).astype(np.int8)
    a = rng.normal(size=(17, 17))
    inv_mass = 1e-4 * (a @ a.T / 17 + np.eye(17))
    eps = rng.uniform(0.05, 0.2, size=C)
    p = rng.normal(size=(C, 17)) * 30
    with Engine(co, splits=(14, 20)) as eng:
        eng.upload_state(i_raw, w)
        lp0, g0 = eng.logp_dlogp(q)
        # reference: host leapfrog, one evaluation per step
        qr, pr, gr = q.copy(), p.copy(), g0.copy()
        for _ in range(L):
            pr = pr + 0.5 * eps[:, None] * gr
            qr = qr + eps[:, None] * (pr @ inv_mass)
            lpr, gr = eng.logp_dlogp(qr)
            pr = pr + 0.5 * eps[:, None] * gr
        dev = torch.device("cuda:0")
        tq, tp, tg = (torch.from_numpy(v.copy()).to(dev) for v in (q, p, g0))
        te, tm = torch.from_numpy(eps).to(dev), torch.from_numpy(inv_mass).to(dev)
        tl = torch.zeros(C, dtype=torch.float64, device=dev)
        di, dw = eng.state_dev(C)
        for rep in range(2):  # twice: the generation counters are re-armed per launch
            tq.copy_(torch.from_numpy(q)), tp.copy_(torch.from_numpy(p)), tg.copy_(torch.from_numpy(g0))
            eng.leapfrog_dev(C, L, tq.data_ptr(), tp.data_ptr(), tg.data_ptr(), tl.data_ptr(), te.data_ptr(), tm.data_ptr(),
                             di, dw, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            eng.leapfrog_status(C)
            np.testing.assert_allclose(tq.cpu().numpy(), qr, rtol=1e-11, atol=1e-11)
            np.testing.assert_allclose(tp.cpu().numpy(), pr, rtol=1e-9, atol=1e-7)
            np.testing.assert_allclose(tl.cpu().numpy(), lpr, rtol=1e-11)
            assert grad_ok(tg.cpu().numpy(), gr, 1e-9)


def test_randomised_cohorts_against_the_oracle():
    """tools/fuzz_parity.py as a test: 40 random small cohorts (G in 2..63, ragged / empty rows,
    0-2 random splits, PCR+ ignored or not, integer and continuous dilutions), logp + gradient,
    Deterministics, conditional log-odds and one bit-exact Gibbs sweep each."""
    import subprocess
    import sys
    from pathlib import Path

    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    root = Path(__file__).resolve().parent.parent
    res = subprocess.run([sys.executable, str(root / "tools" / "fuzz_parity.py"), "40", "7"], capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-1000:]
    assert "40 of 40" in res.stdout


def test_full_size_properties_100k(Engine):
    """BASELINE configs[3] size (100k individuals, several waves of CTAs per launch): the raw sums of 8
    shards of individuals add up to the whole cohort's, finalising them on any shard reproduces the
    unsharded logp / gradient, chains are independent of how they are batched, and a Gibbs sweep of
    a shard equals the same individuals' part of the unsharded sweep."""
    import torch

    from abdpymc_b200.cohort import shard_bounds, synthetic_cohort

    co = synthetic_cohort(100_000)
    rng = np.random.default_rng(18)
    C, splits, world = 3, (14, 20), 8
    q, i_raw, w = draw_points(rng, co.n_gaps, co.n_inds, C)
    dev = torch.device("cuda:0")
    tq = torch.from_numpy(q).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    with Engine(co, splits=splits) as eng:
        lp, g = eng.logp_dlogp(q, i_raw, w)
        lp1, g1 = eng.logp_dlogp(q[1], i_raw[1], w[1])          # one chain on its own
        assert np.isfinite(lp).all() and np.isfinite(g).all()
        assert abs(lp1 - lp[1]) <= 1e-12 * abs(lp[1]) and grad_ok(g1, g[1], 1e-11)
        # the oracle at this size too (recurrence form, one chain: half a minute of NumPy)
        rl, rg = ora.Oracle(co, splits=splits, dense=False).logp_dlogp(q[0], i_raw[0], w[0])
        assert abs(lp[0] - rl) <= RTOL * abs(rl) and grad_ok(g[0], rg)
        eng.upload_state(i_raw, w)
        vals = np.array([[ora.backward(q[k])[0][n] for n in ora.THETA13] for k in range(C)])
        p = np.array([ora.backward(q[k])[0]["p"] for k in range(C)])
        pw = np.array([ora.backward(q[k])[0]["ab_s_p_waner"] for k in range(C)])
        gi, gw, gst = eng.gibbs_sweep(vals, p, pw, seed=3, sweep=1)
    totals = (co.n_inds, int((co.antigen == 1).sum()), int((co.antigen == 0).sum()))
    acc = np.zeros((C, 16))
    out = torch.zeros(C, dtype=torch.float64, device=dev)
    outg = torch.zeros(C, 17, dtype=torch.float64, device=dev)
    flips = 0
    for r in range(world):
        lo, hi = shard_bounds(co.n_inds, r, world)
        with Engine(co.shard(r, world), splits=splits, totals=totals, ind_offset=lo) as eng:
            tis = torch.from_numpy(np.ascontiguousarray(i_raw[:, :, lo:hi])).to(dev)
            tws = torch.from_numpy(np.ascontiguousarray(w[:, lo:hi])).to(dev)
            s = torch.zeros(C, 16, dtype=torch.float64, device=dev)
            eng.sums_dev(C, tq.data_ptr(), 1, tis.data_ptr(), tws.data_ptr(), s.data_ptr(), st)
            torch.cuda.synchronize()
            acc += s.cpu().numpy()
            if r in (0, world - 1):   # the sweep of a shard is the shard of the sweep (RNG keyed by global individual)
                si, sw, sst = eng.gibbs_sweep(vals, p, pw, i_raw[:, :, lo:hi], w[:, lo:hi], seed=3, sweep=1)
                assert np.array_equal(si, gi[:, :, lo:hi]) and np.array_equal(sw, gw[:, lo:hi])
                flips += int(sst[:, 1].sum())
            if r == world - 1:
                ts = torch.from_numpy(acc).to(dev)
                eng.finalize_logp_dev(C, tq.data_ptr(), ts.data_ptr(), out.data_ptr(), outg.data_ptr(), st)
                torch.cuda.synchronize()
    assert np.all(np.abs(out.cpu().numpy() - lp) <= 1e-12 * np.abs(lp))
    assert grad_ok(outg.cpu().numpy(), g, 1e-11)
    assert 0 < flips < int(gst[:, 1].sum())
