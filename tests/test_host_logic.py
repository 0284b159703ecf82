"""CPU tests of the host-side logic: cohort containers and simulator, diagnostics, the sampler's
adaptation on a known target, and the multi-process (gloo, world_size 2) sharding plumbing."""
import os
import socket

import numpy as np
import pytest
import torch

from abdpymc_b200 import diagnostics as dg
from abdpymc_b200.cohort import CohortArrays, shard_bounds, simulate, synthetic_cohort
from oracle import abd_oracle as ora


# ------------------------------------------------------------------------------ cohort / simulator
def test_fixture_shapes(cohorts):
    # reference test_abd.py:666-674: the test cohort is 10 individuals x 26 gaps
    t = cohorts["test_cohort"]
    assert (t.n_inds, t.n_gaps, t.n_rows) == (10, 26, 288)
    c = cohorts["cohort"]
    assert (c.n_inds, c.n_gaps, c.n_rows) == (1520, 31, 35709)
    assert int((c.antigen == 1).sum()) == 19508 and int((c.antigen == 0).sum()) == 16201
    assert c.calculate_splits(True, True) == (14, 20)  # abd.py:215-219 with t0 = 2020-05
    assert c.calculate_splits(True, False) == (14,) and c.calculate_splits(False, True) == (20,)


def test_bootstrap_and_shard_keep_rows(cohorts):
    c = cohorts["cohort"]
    b = c.bootstrap(3000, seed=1)
    assert b.n_inds == 3000 and np.all(np.diff(b.ind) >= 0)
    counts = np.bincount(c.ind, minlength=c.n_inds)
    rng = np.random.default_rng(1)
    picked = rng.integers(0, c.n_inds, size=3000)
    assert np.array_equal(np.bincount(b.ind, minlength=3000), counts[picked])
    # shards are contiguous, cover everything exactly once and keep each individual's rows
    world = 3
    parts = [b.shard(r, world) for r in range(world)]
    assert sum(p.n_inds for p in parts) == b.n_inds and sum(p.n_rows for p in parts) == b.n_rows
    lo, hi = shard_bounds(b.n_inds, 1, world)
    m = (b.ind >= lo) & (b.ind < hi)
    assert np.array_equal(np.sort(parts[1].od), np.sort(b.od[m]))
    assert np.array_equal(parts[1].vacs, b.vacs[lo:hi])
    for n in (7, 8, 9, 10, 1520):
        bounds = [shard_bounds(n, r, 4) for r in range(4)]
        assert bounds[0][0] == 0 and bounds[-1][1] == n
        assert all(a[1] == b_[0] for a, b_ in zip(bounds, bounds[1:]))
        sizes = [b_ - a for a, b_ in bounds]
        assert max(sizes) - min(sizes) <= 1


def test_simulator_dynamics_match_reference(goldens, cohorts):
    """With lam0 = 0 the reference simulator (simulation.py:222-279) is deterministic given
    pcrpos / vacs: our vectorised restatement must reproduce its titers exactly."""
    z, _ = goldens
    for name in ("test_cohort", "cohort"):
        sim = simulate(cohorts[name], lam0=0.0, seed=0)
        take = z[f"sim/{name}/individuals"]
        assert np.array_equal(sim.truth["infections"][take], z[f"sim/{name}/infections"].astype(np.uint8))
        np.testing.assert_allclose(sim.truth["s_titer"][take], z[f"sim/{name}/s_titer"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(sim.truth["n_titer"][take], z[f"sim/{name}/n_titer"], rtol=0, atol=1e-12)


def test_simulated_od_follows_the_elisa_curve():
    co = synthetic_cohort(2000)
    assert co.n_inds == 2000 and co.n_gaps == 31
    titer = np.where(co.antigen == 1, co.truth["s_titer"][co.ind, co.gap], co.truth["n_titer"][co.ind, co.gap])
    resid = co.od - ora.logistic(co.x, titer, -2.2, 1.6)  # simulation.py:76-78 defaults
    assert abs(resid.mean()) < 2e-3 and abs(resid.std() - 0.1) < 2e-3
    # PCR+ months are always infections (simulation.py:248-249); infections are sparse otherwise
    assert np.all(co.truth["infections"][co.pcrpos == 1] == 1)
    assert 0.3 < co.truth["infections"].sum(axis=1).mean() < 2.0


def test_cohort_validation():
    with pytest.raises(ValueError, match="different shapes"):  # abd.py:196-197
        CohortArrays(vacs=np.zeros((3, 4)), pcrpos=np.zeros((3, 5)), ind=[], gap=[], antigen=[], x=[], od=[])
    with pytest.raises(ValueError, match="out of range"):
        CohortArrays(vacs=np.zeros((3, 4)), pcrpos=np.zeros((3, 4)), ind=[3], gap=[0], antigen=[0], x=[0.0], od=[0.0])


# ------------------------------------------------------------------------------ diagnostics
def test_ess_of_ar1_process():
    rng = np.random.default_rng(0)
    for phi in (0.0, 0.5, 0.9):
        x = np.zeros((4, 5000))
        e = rng.normal(size=x.shape)
        for t in range(1, x.shape[1]):
            x[:, t] = phi * x[:, t - 1] + e[:, t]
        want = x.size * (1 - phi) / (1 + phi)
        got = dg.ess_mean(x)
        assert 0.75 * want < got < 1.3 * want, (phi, got, want)
        assert dg.rhat(x) < 1.02
    shifted = rng.normal(size=(4, 500)) + np.array([0, 0, 0, 3.0])[:, None]
    assert dg.rhat(shifted) > 1.3
    s = dg.summary({"a": rng.normal(2.0, 3.0, size=(4, 2000))})
    assert abs(s["a"]["mean"] - 2.0) < 0.2 and abs(s["a"]["sd"] - 3.0) < 0.2 and s["a"]["ess_bulk"] > 4000


# ------------------------------------------------------------------------------ sampler
class GaussianTarget:
    """N(mu, diag(sd^2)) in 5 dimensions with very different scales (exercises the metric)."""

    def __init__(self):
        self.mu = torch.tensor([0.0, 3.0, -2.0, 10.0, 1.0], dtype=torch.float64)
        self.sd = torch.tensor([1.0, 0.1, 5.0, 0.5, 2.0], dtype=torch.float64)

    def logp_dlogp(self, q):
        z = (q - self.mu) / self.sd
        return -0.5 * (z * z).sum(dim=1), -z / self.sd


def test_sampler_recovers_gaussian_moments():
    from abdpymc_b200.sampler import SamplerConfig, _windows, sample

    assert _windows(1000)[-1] == 950 and _windows(1000)[0] == 100
    tgt = GaussianTarget()
    q0 = torch.zeros(8, 5, dtype=torch.float64)
    res = sample(tgt, q0, SamplerConfig(tune=400, draws=600, n_leapfrog=10, seed=3))
    x = res.q.reshape(-1, 5)
    assert np.all(np.abs(x.mean(axis=0) - tgt.mu.numpy()) < 0.25 * tgt.sd.numpy())
    assert np.all(np.abs(x.std(axis=0) / tgt.sd.numpy() - 1) < 0.15)
    # the adapted metric tracks the posterior variances and acceptance sits near the target
    assert np.all(np.abs(np.log(np.diag(res.inv_mass) / tgt.sd.numpy() ** 2)) < 0.7)
    assert 0.6 < res.accept.mean() < 0.97
    for k in range(5):
        assert dg.rhat(res.q[:, :, k]) < 1.05
        assert dg.ess_bulk(res.q[:, :, k]) > 400


def test_nuts_kernel_recovers_gaussian_moments():
    """The batched No-U-Turn transition (SamplerConfig(kernel="nuts")): moments of a badly scaled and
    of a strongly correlated Gaussian, tails, and the antithetic behaviour NUTS shows on Gaussians
    (bulk ESS above the number of draws)."""
    from abdpymc_b200.sampler import SamplerConfig, sample

    tgt = GaussianTarget()
    res = sample(tgt, torch.zeros(6, 5, dtype=torch.float64), SamplerConfig(tune=300, draws=400, seed=3, kernel="nuts"))
    x = res.q.reshape(-1, 5)
    assert np.all(np.abs(x.mean(axis=0) - tgt.mu.numpy()) < 0.15 * tgt.sd.numpy())
    assert np.all(np.abs(x.std(axis=0) / tgt.sd.numpy() - 1) < 0.1)
    assert 0.6 < res.accept.mean() < 0.97
    for k in range(5):
        assert dg.rhat(res.q[:, :, k]) < 1.03 and dg.ess_bulk(res.q[:, :, k]) > 1500

    class Correlated:  # rho = 0.95, scales 1 and 2
        cov = torch.tensor([[1.0, 1.9], [1.9, 4.0]], dtype=torch.float64)
        prec = torch.linalg.inv(cov)

        def logp_dlogp(self, q):
            g = -(q @ self.prec)
            return 0.5 * (q * g).sum(dim=1), g

    res = sample(Correlated(), torch.zeros(6, 2, dtype=torch.float64), SamplerConfig(tune=200, draws=700, seed=5, kernel="nuts"))
    x = res.q.reshape(-1, 2)
    np.testing.assert_allclose(np.cov(x.T), Correlated.cov.numpy(), rtol=0.12, atol=0.12)
    assert abs((np.abs(x[:, 0]) > 2).mean() - 0.0455) < 0.015


# ------------------------------------------------------------------------------ gloo, world_size 2
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, i_raw, w, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from abdpymc_b200 import distributed as D

    dist = D.init_process_group(backend="gloo")
    co = CohortArrays.load("cohort").bootstrap(301, seed=5)
    shard, totals, offset, sl = D.shard_cohort(co, rank, world)
    assert totals == D.cohort_totals(co) and offset == sl.start and shard.n_inds == sl.stop - sl.start
    # the shard's additive pieces (oracle as the stand-in for the GPU kernel on this CPU box)
    o = ora.Oracle(shard, splits=(14, 20), dense=False)
    vals = ora.backward(q)[0]
    th = np.array([vals[n] for n in ora.THETA13])
    ll, g13 = o.loglik_grad(th, i_raw[:, sl], w[sl])
    local = torch.zeros(1, 16, dtype=torch.float64)
    local[0, 0], local[0, 1:14] = ll, torch.from_numpy(g13)
    local[0, 14], local[0, 15] = float(i_raw[:, sl].sum()), float(w[sl].sum())
    D.allreduce_sums(local)
    cs = D.chain_slice(7, rank, world)
    out.put((rank, local.numpy().copy(), (cs.start, cs.stop)))
    dist.barrier()
    dist.destroy_process_group()


def test_individual_sharding_over_gloo_world2():
    """Two processes, gloo: each evaluates its contiguous block of individuals, one all-reduce of
    the additive pieces, and every rank ends up with the whole-cohort answer."""
    import torch.multiprocessing as mp

    rng = np.random.default_rng(12)
    co = CohortArrays.load("cohort").bootstrap(301, seed=5)
    q = ora.forward(ora.sample_prior(rng, co.n_gaps))
    i_raw = (rng.random((co.n_gaps, co.n_inds)) < 0.05).astype(np.int8)
    w = (rng.random(co.n_inds) < 0.5).astype(np.int8)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, i_raw, w, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = ora.Oracle(co, splits=(14, 20), dense=False)
    vals = ora.backward(q)[0]
    th = np.array([vals[n] for n in ora.THETA13])
    ll, g13 = o.loglik_grad(th, i_raw, w)
    for rank, red, cs in got:
        assert abs(red[0, 0] - ll) <= 1e-11 * abs(ll)
        np.testing.assert_allclose(red[0, 1:14], g13, rtol=1e-10, atol=1e-8)
        assert red[0, 14] == i_raw.sum() and red[0, 15] == w.sum()
    assert got[0][2] == (0, 4) and got[1][2] == (4, 7)
    assert np.array_equal(got[0][1], got[1][1])  # identical on both ranks


class OracleShardTarget:
    """The sampler-facing interface of distributed.ShardedTarget with the CPU oracle standing in for the
    kernels: this rank's block of individuals, one all-reduce (gloo) of the additive pieces per evaluation,
    local Gibbs sweeps keyed by GLOBAL individual.  world == 1 (no process group) is the unsharded run."""

    def __init__(self, co, splits, rank, world, i_raw, w, seed):
        from abdpymc_b200 import distributed as D

        self.D, self.world, self.splits, self.seed = D, world, splits, seed
        self.shard, self.totals, self.offset, sl = D.shard_cohort(co, rank, world)
        self.o = ora.Oracle(self.shard, splits=splits, dense=False)
        self.i_raw, self.w = i_raw[:, :, sl].copy(), w[:, sl].copy()
        self.G = co.n_gaps

    def logp_dlogp(self, q):
        C = q.shape[0]
        pieces = torch.zeros(C, 16, dtype=torch.float64)
        back = [ora.backward(q[c].numpy()) for c in range(C)]
        for c in range(C):
            th = np.array([back[c][0][n] for n in ora.THETA13])
            ll, g13 = self.o.loglik_grad(th, self.i_raw[c], self.w[c])
            pieces[c, 0], pieces[c, 1:14] = ll, torch.from_numpy(g13)
            pieces[c, 14], pieces[c, 15] = float(self.i_raw[c].sum()), float(self.w[c].sum())
        self.D.allreduce_sums(pieces)          # no-op without a process group
        n_tot = self.totals[0]
        lp, gr = torch.zeros(C, dtype=torch.float64), torch.zeros(C, 17, dtype=torch.float64)
        for c in range(C):   # the finaliser: priors, transforms, Bernoulli terms with GLOBAL totals (Oracle.logp_dlogp)
            vals, dvals, logj, dlogj = back[c]
            prior, dprior = ora.prior_logp(vals, self.G)
            k_i, n_i, k_w, n_w = float(pieces[c, 14]), self.G * n_tot, float(pieces[c, 15]), n_tot
            p, pw = vals["p"], vals["ab_s_p_waner"]
            bern = k_i * np.log(p) + (n_i - k_i) * np.log1p(-p) + k_w * np.log(pw) + (n_w - k_w) * np.log1p(-pw)
            g_con = dict(dprior)
            g_con["p"] += k_i / p - (n_i - k_i) / (1 - p)
            g_con["ab_s_p_waner"] += k_w / pw - (n_w - k_w) / (1 - pw)
            for k, gk in zip(ora.THETA13, pieces[c, 1:14].numpy()):
                g_con[k] += gk
            lp[c] = float(pieces[c, 0]) + bern + sum(prior.values()) + logj
            gr[c] = torch.from_numpy(np.array([g_con[n] * dvals[n] for n in ora.Q17]) + dlogj)
        return lp, gr

    def gibbs(self, q, sweep):
        for c in range(q.shape[0]):
            vals = ora.backward(q[c].numpy())[0]
            th = np.array([vals[n] for n in ora.THETA13])
            self.i_raw[c], self.w[c], _ = ora.device_gibbs_sweep(self.shard, self.splits, False, th, vals["p"], vals["ab_s_p_waner"],
                                                                 self.i_raw[c], self.w[c], self.seed, sweep, c,
                                                                 ind_offset=self.offset)


def _sharded_sampler_run(co, rank, world, q0, i_raw, w):
    from abdpymc_b200.sampler import SamplerConfig, sample

    tgt = OracleShardTarget(co, (14, 20), rank, world, i_raw, w, seed=4)
    res = sample(tgt, torch.from_numpy(q0), SamplerConfig(tune=4, draws=4, n_leapfrog=3, seed=2, init_step=0.01,
                                                          persistent_trajectories=False))
    return res.q, tgt.i_raw, tgt.w


def _sampler_worker(rank, world, port, q0, i_raw, w, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from abdpymc_b200 import distributed as D

    dist = D.init_process_group(backend="gloo")
    out.put((rank, *_sharded_sampler_run(CohortArrays.load("test_cohort"), rank, world, q0, i_raw, w)))
    dist.barrier()
    dist.destroy_process_group()


def test_compound_sampler_on_sharded_individuals_over_gloo_world2():
    """The sampler's host logic on an individual-sharded cohort, two gloo processes: the replicated HMC state
    stays identical on both ranks, the local Gibbs sweeps (RNG keyed by global individual) reproduce the
    unsharded sweeps, and the whole run equals the one-process run of the same seed."""
    import torch.multiprocessing as mp

    co = CohortArrays.load("test_cohort")
    rng = np.random.default_rng(31)
    C = 2
    q0 = np.stack([ora.forward(ora.sample_prior(rng, co.n_gaps)) for _ in range(C)])
    i_raw = (rng.random((C, co.n_gaps, co.n_inds)) < 0.05).astype(np.int8)
    w = (rng.random((C, co.n_inds)) < 0.5).astype(np.int8)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sampler_worker, args=(r, 2, port, q0, i_raw, w, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    q_one, i_one, w_one = _sharded_sampler_run(co, 0, 1, q0, i_raw, w)
    assert np.array_equal(got[0][1], got[1][1])                       # identical draws on both ranks
    np.testing.assert_allclose(got[0][1], q_one, rtol=1e-9, atol=1e-9)  # ... equal to the unsharded run
    assert np.array_equal(np.concatenate([got[0][2], got[1][2]], axis=2), i_one)
    assert np.array_equal(np.concatenate([got[0][3], got[1][3]], axis=1), w_one)
    assert not np.array_equal(i_one, i_raw)


def test_disk_roundtrip(tmp_path, cohorts):
    t = cohorts["test_cohort"]
    t.to_disk(tmp_path / "cohort_data")
    back = CohortArrays.from_disk(tmp_path / "cohort_data")
    assert np.array_equal(back.vacs, t.vacs) and np.array_equal(back.pcrpos, t.pcrpos) and back.t0 == t.t0
    for a in (0, 1):
        for u, v in zip(back.rows(a), t.rows(a)):
            np.testing.assert_array_equal(u, v)


def test_reference_interface_mirror_errors(cohorts):
    from abdpymc_b200 import abd

    t = cohorts["test_cohort"]
    for bad, msg in [((-1,), "positive"), ((20, 14), "ascending"), ((14, 14), "unique"), ((27,), "largest split"),
                     ((1.5,), "ints")]:
        with pytest.raises(ValueError, match=msg):  # abd.py:604-622
            abd.check_splits(bad, t)
    abd.check_splits(None, t)
    abd.check_splits((14, 20), t)
    if not abd.HAVE_PYMC:
        with pytest.raises(ImportError):
            abd.model(t)
    q = np.arange(17, dtype=float)
    assert np.array_equal(abd.point_to_q17(dict(zip(abd.Q17, q))), q)


def test_cohort_rejects_non_binary_exposures(cohorts):
    """vacs / pcrpos enter the reference numerically (exposure = i + v, abd.py:368,384); the device layout is one bit
    per (individual, gap), so anything but 0 / 1 must be refused, not silently binarised."""
    t = cohorts["test_cohort"]
    bad = t.vacs.copy().astype(np.int64)
    bad[0, 0] = 2
    with pytest.raises(ValueError, match="vacs"):
        CohortArrays(vacs=bad, pcrpos=t.pcrpos, ind=t.ind, gap=t.gap, antigen=t.antigen, x=t.x, od=t.od)
    with pytest.raises(ValueError, match="pcrpos"):
        CohortArrays(vacs=t.vacs, pcrpos=-t.pcrpos.astype(np.int64), ind=t.ind, gap=t.gap, antigen=t.antigen, x=t.x, od=t.od)
    from abdpymc_b200.cohort import splits_from_t0

    assert splits_from_t0("2020-05", True, True) == (14, 20) and splits_from_t0("2020-05", False, False) == ()


def test_op_cache_sends_the_binaries_only_when_they_changed():
    """abd._Cache (shared by the PyTensor Ops and the Gibbs step): object identity decides on the hot path, one content
    comparison when the objects are new, an upload only for a real difference; value + gradient from one call."""
    from abdpymc_b200 import abd

    class StubEngine:
        def __init__(self):
            self.uploads, self.evals, self.state = 0, 0, None

        def upload_state(self, i8, w8):
            self.uploads += 1
            self.state = (i8.copy(), w8.copy())

        def loglik_grad_resident(self, th_ptr, ll_ptr, g_ptr):
            import ctypes

            self.evals += 1
            th = np.ctypeslib.as_array(ctypes.cast(th_ptr, ctypes.POINTER(ctypes.c_double)), (13,))
            np.ctypeslib.as_array(ctypes.cast(ll_ptr, ctypes.POINTER(ctypes.c_double)), (1,))[0] = th.sum() + self.state[0].sum()
            np.ctypeslib.as_array(ctypes.cast(g_ptr, ctypes.POINTER(ctypes.c_double)), (13,))[:] = 2 * th

    eng = StubEngine()
    cache = abd._Cache(eng)
    rng = np.random.default_rng(0)
    i_raw, w = (rng.random((5, 4)) < 0.3).astype(np.int64), (rng.random(4) < 0.5).astype(np.int64)
    th = [np.float64(k) for k in range(13)]
    v1, g1 = cache.get(th + [i_raw, w])
    assert (eng.uploads, eng.evals) == (1, 1) and v1 == sum(range(13)) + i_raw.sum() and np.array_equal(g1, 2 * np.arange(13.0))
    cache.get(th + [i_raw, w])                                   # same point: neither an upload nor a launch
    assert (eng.uploads, eng.evals) == (1, 1)
    th2 = [np.float64(k + 0.5) for k in range(13)]
    cache.get(th2 + [i_raw, w])                                  # a leapfrog: new scalars, the same binary objects
    assert (eng.uploads, eng.evals) == (1, 2)
    cache.get(th2 + [i_raw.copy(), w.copy()])                    # new objects, equal contents (PyMC re-set the values)
    assert (eng.uploads, eng.evals) == (1, 2)
    other = 1 - i_raw
    v3, _ = cache.get(th2 + [other, w])                          # a real change: uploaded, evaluated
    assert (eng.uploads, eng.evals) == (2, 3) and v3 == sum(k + 0.5 for k in range(13)) + other.sum()
    new_i, new_w = (other != 0).astype(np.int8), (w != 0).astype(np.int8)
    cache.note_resident(new_i, new_w)                            # what the Gibbs step does after a sweep
    cache.get(th2 + [new_i.astype(np.int64), new_w.astype(np.int64)])
    assert eng.uploads == 2 and eng.evals == 4                   # no upload; evaluated again (the state may have changed)
    strict = abd._Cache(eng, strict=True)
    strict.get(th + [i_raw, w])
    i_raw[0, 0] ^= 1                                             # mutated in place: only the strict cache notices
    n = eng.uploads
    strict.get(th + [i_raw, w])
    assert eng.uploads == n + 1
