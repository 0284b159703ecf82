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
    assert np.all((z["ab_n_rho"] > 0) & (z["ab_n_rho"] < 1)) and np.all(z["it_n_sigma"] > 0)
