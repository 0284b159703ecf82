"""The PyMC-facing glue of abdpymc_b200/abd.py (PyTensor Ops, model(), GpuBinaryGibbs) executed
against a protocol stand-in for PyMC / PyTensor (tests/fake_pymc.py): PyMC itself is not installable
offline, and glue that never ran is glue with typos.  Checks that the Ops hand the right values to
the library in the right order and that the step method speaks the step protocol; the numbers are
compared with the CPU oracle.  tests/test_pymc_parity.py is the test against a real PyMC."""
import numpy as np
import pytest

from oracle import abd_oracle as ora

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def glue():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from abdpymc_b200 import build

    build.build()
    import fake_pymc

    return fake_pymc, fake_pymc.load_abd_with_fake_pymc()


def test_model_ops_and_step_method_against_the_oracle(glue, cohorts):
    fake, abd = glue
    from abdpymc_b200.engine import Q17, Q_OF_THETA, THETA13

    co, splits = cohorts["test_cohort"], (14, 20)
    m = abd.model(co, splits=splits)
    # same RVs, in the reference's declaration order, with PyMC's value-variable names
    assert [v.name for v in m.value_vars if v.name not in ("i_raw", "ab_s_waner")] == Q17
    assert m.dims["i_raw"] == ("gap", "ind") and m.dims["ab_s_waner"] == "ind" and m.dims["i"] == ("gap", "ind")
    assert list(m.coords["ind"]) == list(range(co.n_inds)) and list(m.coords["gap"]) == list(range(co.n_gaps))
    assert [d.name for d in m.deterministics] == ["i", "ab_n_mu", "ab_s_mu"] and [p.name for p in m.potentials] == ["loglik"]

    rng = np.random.default_rng(0)
    vals = ora.sample_prior(rng, co.n_gaps)
    i_raw = (rng.random((co.n_gaps, co.n_inds)) < 0.06).astype(np.int64)   # PyMC holds Bernoulli values as int64
    waner = (rng.random(co.n_inds) < 0.5).astype(np.int64)
    point = {**{k: np.float64(v) for k, v in vals.items()}, "i_raw": i_raw, "ab_s_waner": waner}
    o = ora.Oracle(co, splits=splits)
    th = np.array([vals[n] for n in THETA13])
    ll, g13 = o.loglik_grad(th, i_raw, waner)

    # the Potential: one Op evaluation = the two observed Normal log-densities of abd.py:445-469
    pot = m["loglik"]
    assert abs(float(fake.evaluate(pot, point)) - ll) <= 1e-10 * abs(ll)
    # its gradient: Op.grad returns gz * d loglik / d theta_k for the 13 scalars, undefined for the binaries
    node = pot.owner
    gz = fake.Variable(const=np.float64(2.0))
    grads = node.op.grad(node.inputs, [gz])
    assert len(grads) == 15 and all(isinstance(u, fake._Undefined) for u in grads[13:])
    got = np.array([float(fake.evaluate(gk, point)) for gk in grads[:13]])
    scale = np.maximum(np.abs(g13), 1e-3 * np.abs(g13).max())
    assert np.all(np.abs(got - 2.0 * g13) <= 2e-10 * scale)
    # value and gradient at one point come out of one launch (cache)
    launches = m.abd_engine.launch_count
    fake.evaluate(pot, point)
    [fake.evaluate(gk, point) for gk in grads[:13]]
    assert m.abd_engine.launch_count == launches

    # the three Deterministics
    ri, rn, rs = o.deterministics(th, i_raw, waner)
    assert np.array_equal(fake.evaluate(m["i"], point), ri)
    np.testing.assert_allclose(fake.evaluate(m["ab_n_mu"], point), rn, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(fake.evaluate(m["ab_s_mu"], point), rs, rtol=1e-12, atol=1e-12)

    # the step method: one step = one sweep, on a PyMC point (unconstrained values by value-variable name)
    step = abd.GpuBinaryGibbs(model=m, seed=11)
    assert [v.name for v in step.vars] == ["i_raw", "ab_s_waner"]
    q = ora.forward(vals)
    pm_point = {**{name: np.float64(q[k]) for k, name in enumerate(Q17)}, "i_raw": i_raw, "ab_s_waner": waner}
    new, stats = step.step(pm_point)
    want_i, want_w, want_st = ora.device_gibbs_sweep(co, splits, False, th, vals["p"], vals["ab_s_p_waner"], i_raw, waner,
                                                     step.seed, 0, 0)
    assert np.array_equal(new["i_raw"], want_i) and np.array_equal(new["ab_s_waner"], want_w)
    assert new["i_raw"].dtype == np.int64 and new["i_raw"].shape == i_raw.shape
    assert all(new[name] == pm_point[name] for name in Q17)          # the continuous values pass through
    assert stats == [{"p_jump": want_st[1] / max(want_st[0], 1), "tune": True}]
    step.stop_tuning()
    new2, stats2 = step.step(new)
    assert stats2[0]["tune"] is False and step.sweep == 2
    C = abd.Competence
    assert abd.GpuBinaryGibbs.competence(m["i_raw"], False) == C.COMPATIBLE
    assert abd.GpuBinaryGibbs.competence(m["p"], True) == C.INCOMPATIBLE

    # the binaries travel only when they changed: NUTS evaluates many leapfrogs between two sweeps
    cache = m.abd_cache
    up0 = cache.uploads
    for k in range(4):   # new continuous values, the same binaries (what `new2` holds is what the last sweep left on the device)
        pt_k = {**{name: np.float64(v * (1 + 0.01 * k)) for name, v in vals.items()}, "i_raw": new2["i_raw"],
                "ab_s_waner": new2["ab_s_waner"]}
        th_k = np.array([pt_k[n] for n in THETA13])
        ll_k, _ = o.loglik_grad(th_k, new2["i_raw"], new2["ab_s_waner"])
        assert abs(float(fake.evaluate(pot, pt_k)) - ll_k) <= 1e-10 * abs(ll_k)
    assert cache.uploads == up0
    other = {**point, "i_raw": 1 - i_raw}      # someone evaluates another point: uploaded, and correct
    ll_o, _ = o.loglik_grad(th, 1 - i_raw, waner)
    assert abs(float(fake.evaluate(pot, other)) - ll_o) <= 1e-10 * abs(ll_o) and cache.uploads == up0 + 1

    # model() validates splits like the reference (abd.py:604-622, :881-882)
    with pytest.raises(ValueError, match="ascending"):
        abd.model(co, splits=(20, 14))
    with pytest.raises(NotImplementedError):
        abd.model(co, splits=(1, 2, 3))
