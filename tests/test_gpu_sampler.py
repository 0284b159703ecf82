"""GPU tests of the end-to-end sampling path (built-in HMC + GPU Gibbs, abdpymc-infer CLI)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from abdpymc_b200 import build

    build.build()
    return torch.device("cuda:0")


def test_builtin_sampler_recovers_simulated_truth(gpu):
    """Simulate a 400-individual cohort with the reference's default dynamics (ELISA b = -2.2,
    d = 1.6, sd = 0.1; init = -2), sample the model, and check convergence, the ELISA curve
    parameters (which the inference model shares with the simulator) and the infections."""
    from abdpymc_b200 import diagnostics as dg
    from abdpymc_b200.abd import infer_builtin
    from abdpymc_b200.cohort import synthetic_cohort

    co = synthetic_cohort(400)
    res, post, last = infer_builtin(co, (14, 20), False, tune=600, draws=600, chains=4, seed=1)
    assert np.isfinite(res.q).all() and np.isfinite(res.logp).all()
    assert 0.55 < res.accept.mean() < 0.995
    summ = dg.summary(post)
    # parameters that do not hinge on the slowly mixing infection indicators converge quickly ...
    for k in ("p", "it_n_b", "it_n_d", "it_s_d", "ab_n_init", "ab_s_init", "ab_n_perm", "ab_s_perm"):
        assert summ[k]["rhat"] < 1.1 and summ[k]["ess_bulk"] > 100, (k, summ[k])
    # ... the others (waning rates, sigmas: coupled to where the infections sit) at least agree roughly
    for k, v in summ.items():
        assert v["rhat"] < 2.0, (k, v)
    assert abs(summ["it_n_sigma"]["mean"] - 0.1) < 0.03 and abs(summ["it_s_sigma"]["mean"] - 0.1) < 0.05
    assert abs(summ["it_n_d"]["mean"] - 1.6) < 0.15 and abs(summ["it_n_b"]["mean"] + 2.2) < 0.5
    assert abs(summ["ab_n_init"]["mean"] + 2.0) < 0.6
    # tempinf / tempvac touch only their priors (abd.py:263-274 ignores temp): Gamma(mu 1, sigma 0.5)
    for k in ("ab_s_tempinf", "ab_s_tempvac"):
        assert abs(summ[k]["mean"] - 1.0) < 0.1 and abs(summ[k]["sd"] - 0.5) < 0.1
    # infections: PCR+ months are inferred with certainty (unless masked by the 3-gap rule), and
    # the posterior infection probability separates true infections from non-infections
    pi = res.means["i"]  # (G, N) posterior mean of the constrained infections
    truth = co.truth["infections"].T
    assert pi.shape == truth.shape
    assert pi[truth == 1].mean() > 0.5 and pi[truth == 0].mean() < 0.05
    assert set(np.unique(last["i_raw"])) <= {0, 1}


def test_cli_without_pymc_writes_posterior(gpu, tmp_path):
    from abdpymc_b200 import abd
    from abdpymc_b200.cohort import CohortArrays

    if abd.HAVE_PYMC:
        pytest.skip("PyMC present: the CLI takes the pm.sample path")
    CohortArrays.load("test_cohort").to_disk(tmp_path / "cohort_data")
    out = tmp_path / "post.npz"
    abd.main(["--tune", "60", "--draws", "40", "--ititers_data", str(tmp_path / "cohort_data"), "--split_delta",
              "--split_omicron", "--netcdf", str(out), "--chains", "2"])
    z = np.load(out)
    assert z["p"].shape == (2, 40) and z["ab_s_rho"].shape == (2, 40)
    assert z["mean_i"].shape == (26, 10) and z["mean_ab_n_mu"].shape == (26, 10)  # dims ("gap", "ind")
    assert z["last_i_raw"].shape == (2, 26, 10)
    # thinned draws of the Deterministics, dims (chain, draw, gap, ind) as in the InferenceData
    assert z["i"].shape == (2, 40, 26, 10) and z["i"].dtype == np.int8 and set(np.unique(z["i"])) <= {0, 1}
    assert z["ab_n_mu"].shape == z["ab_s_mu"].shape == (2, 40, 26, 10) and list(z["thinned_draw"]) == list(range(40))
    np.testing.assert_allclose(z["i"].mean(axis=(0, 1)), z["mean_i"], atol=1e-12)
    np.testing.assert_allclose(z["ab_n_mu"].mean(axis=(0, 1)), z["mean_ab_n_mu"], rtol=1e-5, atol=1e-5)
    assert np.all((z["ab_n_rho"] > 0) & (z["ab_n_rho"] < 1)) and np.all(z["it_n_sigma"] > 0)
    # no split flags: the OneTimeChunk model (abd.py:640-649), SURVEY 8d config 1 "splits (14, 20) and none"
    out1 = tmp_path / "post1.npz"
    abd.main(["--tune", "40", "--draws", "20", "--ititers_data", str(tmp_path / "cohort_data"), "--netcdf", str(out1),
              "--chains", "2", "--thinned", "5"])
    z1 = np.load(out1)
    assert z1["p"].shape == (2, 20) and z1["i"].shape == (2, 5, 26, 10) and np.isfinite(z1["ab_s_rho"]).all()
    # the block-draw update rule through the same front end
    out2 = tmp_path / "post2.npz"
    abd.main(["--tune", "40", "--draws", "20", "--ititers_data", str(tmp_path / "cohort_data"), "--split_delta",
              "--split_omicron", "--netcdf", str(out2), "--chains", "2", "--gibbs_mode", "2"])
    z2 = np.load(out2)
    assert z2["p"].shape == (2, 20) and set(np.unique(z2["last_i_raw"])) <= {0, 1}
    # --cache: the first run writes the preprocessed cohort, the second reads it instead of the CSV files and
    # reproduces the run exactly (same seed); a cache of other splits is refused
    cache = tmp_path / "cohort.abdcache"
    runs = []
    for k in range(2):
        outc = tmp_path / f"postc{k}.npz"
        abd.main(["--tune", "30", "--draws", "10", "--ititers_data", str(tmp_path / "cohort_data"), "--split_delta",
                  "--split_omicron", "--netcdf", str(outc), "--chains", "2", "--cache", str(cache), "--thinned", "0"])
        assert cache.exists()
        runs.append(np.load(outc))
    assert np.array_equal(runs[0]["ab_n_perm"], runs[1]["ab_n_perm"]) and np.array_equal(runs[0]["last_i_raw"], runs[1]["last_i_raw"])
    with pytest.raises(ValueError, match="splits"):
        abd.main(["--tune", "5", "--draws", "5", "--ititers_data", str(tmp_path / "cohort_data"), "--netcdf",
                  str(tmp_path / "x.npz"), "--chains", "2", "--cache", str(cache)])


def test_hmc_transition_kernels_against_host_formulas(gpu):
    """abd_hmc_begin_dev / abd_hmc_end_dev (momentum refresh, energy, accept, dual averaging) against
    the same formulas in torch / the sampler's host-side _DualAveraging."""
    import torch

    from abdpymc_b200.cohort import CohortArrays
    from abdpymc_b200.engine import AbdEngine
    from abdpymc_b200.sampler import _DualAveraging

    C, D = 4096, 17
    f64 = dict(dtype=torch.float64, device=gpu)
    g = torch.Generator(device=gpu).manual_seed(0)
    a = torch.randn(D, D, generator=g, **f64)
    inv_mass = (a @ a.T / D + torch.eye(D, **f64)).contiguous()          # Sigma
    chol = torch.linalg.cholesky(inv_mass)
    linv_t = torch.linalg.solve_triangular(chol.T.contiguous(), torch.eye(D, **f64), upper=True).contiguous()
    q, grad = torch.randn(C, D, generator=g, **f64), torch.randn(C, D, generator=g, **f64)
    logp = torch.randn(C, generator=g, **f64) * 10
    qw, pw, gw, h0 = torch.empty(C, D, **f64), torch.empty(C, D, **f64), torch.empty(C, D, **f64), torch.empty(C, **f64)
    with AbdEngine(CohortArrays.load("test_cohort"), splits=(14, 20)) as eng:
        eng.hmc_begin_dev(C, q.data_ptr(), grad.data_ptr(), logp.data_ptr(), linv_t.data_ptr(), 7, 3, qw.data_ptr(), pw.data_ptr(),
                          gw.data_ptr(), h0.data_ptr())
        torch.cuda.synchronize()
        assert torch.equal(qw, q) and torch.equal(gw, grad)
        kin = 0.5 * ((pw @ inv_mass) * pw).sum(dim=1)                    # p' Sigma p / 2 == z.z / 2
        assert torch.allclose(h0, -logp + kin, rtol=1e-12, atol=1e-12)
        # p ~ N(0, Sigma^-1): sample covariance over 4096 chains
        cov = (pw.T @ pw) / C
        want = torch.linalg.inv(inv_mass)
        assert (cov - want).abs().max() < 6 * want.diagonal().max() / C**0.5
        assert pw.mean(dim=0).abs().max() < 5 * want.diagonal().max().sqrt() / C**0.5
        # a different iteration / seed gives different momenta, the same one the same
        pw2 = torch.empty_like(pw)
        eng.hmc_begin_dev(C, q.data_ptr(), grad.data_ptr(), logp.data_ptr(), linv_t.data_ptr(), 7, 4, qw.data_ptr(), pw2.data_ptr(),
                          gw.data_ptr(), h0.data_ptr())
        torch.cuda.synchronize()
        assert not torch.equal(pw, pw2)

        # end of the transition: proposals with a known energy change
        dh = torch.linspace(-3, 3, C, **f64)                              # h0 - h1
        dh[::97] = float("nan")
        q_new, g_new = q + 1.0, grad - 1.0
        lpw = -(h0 - dh) + kin_of(pw2, inv_mass)                          # so that h1 = h0 - dh
        acc, eps = torch.empty(C, **f64), torch.full((C,), 0.1, **f64)
        da = torch.zeros(C, 4, **f64)
        da[:, 0] = torch.log(10 * eps)
        ref = _DualAveraging(eps.clone(), 0.8, gpu)
        qc, gc, lc = q.clone(), grad.clone(), logp.clone()
        for it in range(3):
            eng.hmc_end_dev(C, qc.data_ptr(), gc.data_ptr(), lc.data_ptr(), q_new.data_ptr(), pw2.data_ptr(), g_new.data_ptr(),
                            lpw.data_ptr(), inv_mass.data_ptr(), h0.data_ptr(), 7, 10 + it, acc.data_ptr(), da.data_ptr(),
                            eps.data_ptr(), 1, 0.8)
            torch.cuda.synchronize()
            want_acc = torch.exp(torch.clamp(torch.nan_to_num(dh, nan=-float("inf")), max=0.0))
            assert torch.allclose(acc, want_acc, rtol=1e-10, atol=1e-300)
            assert torch.allclose(eps, ref.update(want_acc), rtol=1e-12)
        assert torch.allclose(torch.exp(da[:, 2]), ref.final(), rtol=1e-12)
        taken = (qc == q_new).all(dim=1)
        assert torch.equal(taken, (gc == g_new).all(dim=1)) and torch.equal(lc[taken], lpw[taken])
        assert not taken[::97].any()                                      # non-finite energy: rejected
        assert taken[dh >= 0].all()                                       # never reject a downhill move
        frac = taken[(dh < 0) & torch.isfinite(dh)].double().mean().item()
        assert 0.50 < frac < 0.62                                         # E[1 - (1 - e^dh)^3] over dh in (-3, 0) = 0.562


def test_device_nuts_matches_host_nuts(gpu):
    """The No-U-Turn tree kept on the device (abd_nuts_*_dev, one launch per leaf) against the same algorithm as
    torch ops driven from the host, on the bundled test cohort WITH data: tree depths, acceptance statistic,
    divergence rate and posterior summaries of the 17 scalars agree (different random streams: statistically)."""
    import torch

    from abdpymc_b200 import diagnostics as dg
    from abdpymc_b200.cohort import CohortArrays
    from abdpymc_b200.engine import Q17_RV, AbdEngine, backward, forward
    from abdpymc_b200.sampler import AbdTarget, SamplerConfig, sample

    co = CohortArrays.load("test_cohort")
    G, N, C = co.n_gaps, co.n_inds, 8
    x0 = np.array([1.0 / G, 2, 1, 10 / 11, -2, 2, 10 / 11, 0.5, 1, 1, -2, -1, 2, 1, -1, 2, 1], dtype=np.float64)
    q0 = forward(x0)[None, :] + np.random.default_rng(2).uniform(-0.5, 0.5, size=(C, 17))
    out = {}
    for name, fused, tune, draws in (("device", True, 600, 3000), ("host", False, 300, 600)):
        with AbdEngine(co, splits=(14, 20)) as eng:
            tgt = AbdTarget(eng, C, np.zeros((C, G, N), np.int8), np.zeros((C, N), np.int8), seed=5)
            cfg = SamplerConfig(tune=tune, draws=draws, seed=7, kernel="nuts", persistent_trajectories=fused)
            res = sample(tgt, torch.from_numpy(q0).to(gpu), cfg)
        x = backward(res.q)
        out[name] = (res, dg.summary({n_: x[:, :, k] for k, (n_, _) in enumerate(Q17_RV)}))
    rd, sd_ = out["device"]
    rh, sh = out["host"]
    assert rd.wall_s / (600 + 3000) < 0.25 * rh.wall_s / (300 + 600)           # an order of magnitude fewer ms per draw
    assert abs(rd.stats["tree_depth"].mean() - rh.stats["tree_depth"].mean()) < 0.5
    assert abs(rd.accept.mean() - rh.accept.mean()) < 0.08
    assert abs(rd.stats["diverging"].mean() - rh.stats["diverging"].mean()) < 0.05
    for name in sd_:
        a, b = sd_[name], sh[name]
        se = np.hypot(a["sd"] / np.sqrt(max(a["ess_bulk"], 20.0)), b["sd"] / np.sqrt(max(b["ess_bulk"], 20.0)))
        assert abs(a["mean"] - b["mean"]) < 5 * se + 0.02 * b["sd"], (name, a, b)
        assert 0.6 < a["sd"] / b["sd"] < 1.6, (name, a, b)


def kin_of(p, inv_mass):
    return 0.5 * ((p @ inv_mass) * p).sum(dim=1)


@pytest.mark.parametrize("gibbs_mode,kernel", [(0, "hmc"), (2, "hmc"), (0, "nuts"), (0, "nuts_host")])
def test_sampler_recovers_the_priors_on_a_cohort_without_data(gpu, gibbs_mode, kernel):
    """Statistical check of the whole transition (priors, transforms + Jacobians, HMC on the device,
    Gibbs over the indicators) against distributions known in closed form: without OD rows the
    posterior of the 17 scalars IS their prior (scipy.stats moments), and the indicators are
    Bernoulli(p) / Bernoulli(p_waner) given those."""
    import torch
    from scipy import stats

    from abdpymc_b200 import diagnostics as dg
    from abdpymc_b200.cohort import CohortArrays
    from abdpymc_b200.engine import Q17_RV, AbdEngine, backward, forward
    from abdpymc_b200.sampler import AbdTarget, SamplerConfig, sample

    rng = np.random.default_rng(1)
    G, N, C = 10, 6, 8
    empty = np.zeros(0)
    co = CohortArrays(vacs=rng.random((N, G)) < 0.1, pcrpos=np.zeros((N, G)), ind=empty.astype(int), gap=empty.astype(int),
                      antigen=empty.astype(int), x=empty, od=empty)

    def gam(mu, sigma):
        return stats.gamma(a=mu * mu / sigma**2, scale=sigma**2 / mu)

    prior = {"p": stats.beta(1, G - 1), "ab_n_perm": gam(2, .5), "ab_n_temp": gam(1, .5), "ab_n_rho": stats.beta(10, 1),
             "ab_n_init": stats.norm(-2, 1), "ab_s_perm": gam(2, .5), "ab_s_rho": stats.beta(10, 1),
             "ab_s_p_waner": stats.beta(1, 1), "ab_s_tempinf": gam(1, .5), "ab_s_tempvac": gam(1, .5),
             "ab_s_init": stats.norm(-2, 1), "it_n_b": stats.norm(-1, .5), "it_n_d": stats.norm(2, .5),
             "it_n_sigma": stats.expon(), "it_s_b": stats.norm(-1, .5), "it_s_d": stats.norm(2, .5), "it_s_sigma": stats.expon()}
    with AbdEngine(co, splits=(4,)) as eng:
        tgt = AbdTarget(eng, C, np.zeros((C, G, N), np.int8), np.zeros((C, N), np.int8), seed=3, gibbs_mode=gibbs_mode)
        x0 = np.array([1.0 / G, 2, 1, 10 / 11, -2, 2, 10 / 11, 0.5, 1, 1, -2, -1, 2, 1, -1, 2, 1], dtype=np.float64)
        q0 = forward(x0)[None, :] + rng.uniform(-1, 1, size=(C, 17))
        # "nuts": the tree bookkeeping on the device (abd_nuts_*_dev); "nuts_host": the same algorithm as torch ops
        # driven from the host (~10 ms per draw)
        n_tune, n_draws = (300, 1000) if kernel == "nuts_host" else (1000, 6000)
        cfg = SamplerConfig(tune=n_tune, draws=n_draws, seed=3, kernel=kernel.split("_")[0],
                            persistent_trajectories=kernel != "nuts_host")
        res = sample(tgt, torch.from_numpy(q0).to(gpu), cfg)
        i_raw, waner = tgt.state()
    if kernel.startswith("nuts"):
        depth, div = res.stats["tree_depth"], res.stats["diverging"]
        # (no tight bounds here: the data-free posterior is the PRIOR in log / logodds space, stiff in its tails, where
        # trees run to the maximum depth or diverge; test_device_nuts_matches_host_nuts compares the two drivers)
        assert depth.shape == (C, n_draws) and depth.min() >= 1 and depth.max() <= cfg.max_treedepth
        assert div.mean() < 0.3 and 0.5 < res.accept.mean() < 0.99
    x = backward(res.q)
    summ = dg.summary({name: x[:, :, k] for k, (name, _) in enumerate(Q17_RV)})
    for name, dist in prior.items():
        v = summ[name]
        se = dist.std() / np.sqrt(max(v["ess_bulk"], 50.0))
        assert v["rhat"] < 1.05, (name, v)
        assert abs(v["mean"] - dist.mean()) < 5 * se + 0.01 * dist.std(), (name, v, dist.mean())
        assert abs(v["sd"] / dist.std() - 1.0) < 0.15, (name, v, dist.std())
        # a few quantiles of the pooled draws against the prior's
        draws = x[:, :, [k for k, (n_, _) in enumerate(Q17_RV) if n_ == name][0]].ravel()
        for qq in (0.1, 0.5, 0.9):
            assert abs((draws < dist.ppf(qq)).mean() - qq) < 0.05, (name, qq)
    assert set(np.unique(i_raw)) <= {0, 1} and set(np.unique(waner)) <= {0, 1}
