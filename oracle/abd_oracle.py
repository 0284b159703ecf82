"""CPU ORACLE for the abdpymc inference hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this module, and only as the checker or as the timed CPU
baseline.  The product path (``abdpymc_b200``) never imports it and has no CPU fallback.

It is a plain NumPy fp64 restatement of the algorithm in the reference's ``abdpymc/abd.py``
(all ``abd.py:NNN`` citations are relative to /root/reference/abdpymc/) and of the PyMC
distribution log-densities / transforms that file invokes (third-party ``pymc``, unpinned in
the reference's pyproject.toml:10; restated from PyMC 5's published formulas).

Parity pinning
--------------
* The integer prologue and the response functions are pinned against the golden vectors the
  reference's own tests hold (transcribed in tests/golden/reference_kats.json by
  tests/golden/make_golden.py, checked in tests/test_oracle.py).
* The joint log-density, its gradient, the Deterministics and the conditional log-odds are
  pinned against tests/golden/model_goldens.npz, produced by EXECUTING the reference's own
  ``abd.model()`` (its dense (G,G,N) formulation) in the build container on a NumPy stand-in
  for PyTensor/PyMC (tests/golden/refshim.py), with gradients by complex-step differentiation
  of that same reference code.
* NOT pinned by a live PyMC: the closed-form log-densities of Beta / Gamma / Normal /
  Exponential / Bernoulli and the log / logodds transforms (PyMC is not installable offline).
  For those pieces parity is "restated from documentation"; tests/test_pymc_parity.py runs the
  comparison against a real PyMC whenever one is importable.

Layout conventions (same as the reference graph): matrices over (gap, individual) are (G, N)
with the individual axis contiguous (abd.py:413-418 transposes the on-disk (N, G) arrays).
"""

from __future__ import annotations

import numpy as np
from scipy.special import gammaln

# ---------------------------------------------------------------------------------------
# parameter vectors
# ---------------------------------------------------------------------------------------
#: the 13 (constrained) scalars that reach the data likelihood, in the order the C-ABI uses
THETA13 = [
    "ab_n_perm", "ab_n_temp", "ab_n_rho", "ab_n_init",
    "ab_s_perm", "ab_s_rho", "ab_s_init",
    "it_n_b", "it_n_d", "it_n_sigma",
    "it_s_b", "it_s_d", "it_s_sigma",
]  # fmt: skip

#: PyMC value variables of the 17 continuous RVs in declaration order
#: (abd.py:424 -> 329-340 -> 367-388 -> 464-467) with their default transforms
VALUE_VARS = [
    ("p", "logodds"), ("ab_n_perm", "log"), ("ab_n_temp", "log"), ("ab_n_rho", "logodds"),
    ("ab_n_init", None), ("ab_s_perm", "log"), ("ab_s_rho", "logodds"),
    ("ab_s_p_waner", "logodds"), ("ab_s_tempinf", "log"), ("ab_s_tempvac", "log"),
    ("ab_s_init", None), ("it_n_b", None), ("it_n_d", None), ("it_n_sigma", "log"),
    ("it_s_b", None), ("it_s_d", None), ("it_s_sigma", "log"),
]  # fmt: skip
Q17 = [n for n, _ in VALUE_VARS]
#: position in the 17-vector of each of the 13 likelihood parameters
Q17_OF_THETA13 = [Q17.index(n) for n in THETA13]


# ---------------------------------------------------------------------------------------
# integer prologue: i_raw -> i   (abd.py:560-882)
# ---------------------------------------------------------------------------------------
def mask_multiple_infections(arr):
    """abd.py:792-818: zero every 1 that follows an earlier 1 in its column."""
    arr = np.asarray(arr)
    return np.where(arr.cumsum(axis=0) > 1, 0, arr)


def chunk_bounds(n_gaps, splits):
    """Row ranges of the time chunks: one per split + 1 (abd.py:685-686, 713-715)."""
    edges = [0, *(splits or ()), n_gaps]
    return list(zip(edges[:-1], edges[1:]))


def mask_multiple_infections_chunks(arr, splits):
    """abd.py:774-789 (3 chunks) and 821-862 (2 chunks)."""
    arr = np.asarray(arr)
    return np.concatenate([mask_multiple_infections(arr[a:b]) for a, b in chunk_bounds(len(arr), splits)])


def incorporate_pcrpos(i_raw, pcrpos):
    """abd.py:732-771: a column with any PCR+ is replaced by the PCR+ column."""
    i_raw, pcrpos = np.asarray(i_raw), np.asarray(pcrpos)
    return np.where(pcrpos.any(axis=0), pcrpos, i_raw)


def mask_future_infection(i0, im3, im2, im1):
    """abd.py:581-601."""
    return np.where(np.asarray(im3) | np.asarray(im2) | np.asarray(im1), 0, i0)


def mask_three_gaps(arr):
    """abd.py:560-578: scan over gaps with taps -3,-2,-1 on the OUTPUT, zero initial state,
    int8."""
    arr = np.asarray(arr).astype(np.int8)
    out = np.zeros((arr.shape[0] + 3,) + arr.shape[1:], dtype=np.int8)
    for t in range(arr.shape[0]):
        out[t + 3] = mask_future_infection(arr[t], out[t], out[t + 1], out[t + 2])
    return out[3:]


def constrain_infections(i_raw, pcrpos, splits=None):
    """OneTimeChunk (abd.py:640-649) when there are no splits, else MultipleTimeChunks
    (abd.py:658-667) with per-chunk PCR+ override (abd.py:691-697, 722-729)."""
    i_raw, pcrpos = np.asarray(i_raw), np.asarray(pcrpos)
    if not splits:
        return mask_three_gaps(np.where(i_raw + pcrpos > 0, 1, 0))
    masked = mask_multiple_infections_chunks(i_raw, splits)
    merged = np.concatenate(
        [incorporate_pcrpos(masked[a:b], pcrpos[a:b]) for a, b in chunk_bounds(len(i_raw), splits)]
    )
    return mask_three_gaps(merged)


# ---------------------------------------------------------------------------------------
# responses   (abd.py:224-306)
# ---------------------------------------------------------------------------------------
def decay_design(n_gaps):
    """abd.py:224-239."""
    k = np.arange(n_gaps)
    return np.maximum(0, k[None, :] - k[:, None])


def temp_response_dense(exposure, rho):
    """Dense (G,G,N) formulation shared by _temp_response_scalar_rho (abd.py:242-260, without
    the ``* temp`` factor) and _temp_response_vector_rho (abd.py:263-274, which never applies
    ``temp``).  ``rho`` is a scalar or an (N,) vector."""
    exposure = np.asarray(exposure)
    g = exposure.shape[0]
    design = decay_design(g)
    offset = np.tril(np.ones_like(design), -1)
    rho = np.asarray(rho)
    kernel = rho ** design[..., None] - offset[..., None]  # (G, G, N) or (G, G, 1)
    return (kernel * exposure[:, None, :]).sum(axis=0)


def dtemp_response_dense(exposure, rho):
    """d/d rho of temp_response_dense (the (G,G,N) route PyTensor's autodiff takes)."""
    exposure = np.asarray(exposure)
    g = exposure.shape[0]
    design = decay_design(g)[..., None]
    rho = np.asarray(rho)
    with np.errstate(divide="ignore", invalid="ignore"):
        kernel = np.where(design > 0, design * rho ** np.maximum(design - 1, 0), 0.0)
    return (kernel * exposure[:, None, :]).sum(axis=0)


def temp_response_scan(exposure, rho, temp=1.0):
    """The O(G N) recurrence of _temp_response_scan (abd.py:277-293): prev*rho + e*temp."""
    exposure = np.asarray(exposure, dtype=float)
    out = np.empty_like(exposure)
    prev = np.zeros(exposure.shape[1])
    for t in range(exposure.shape[0]):
        prev = prev * rho + exposure[t] * temp
        out[t] = prev
    return out


def perm_indicator(exposure):
    """abd.py:296-306 without the ``perm`` factor: 1 once cumsum(exposure) > 0."""
    return (np.cumsum(exposure, axis=0) > 0).astype(float)


def logistic(x, a, b, d):
    """abd.py:556-557."""
    return d / (1 + np.exp(-b * (x - a)))


# ---------------------------------------------------------------------------------------
# the model at a point
# ---------------------------------------------------------------------------------------
class Oracle:
    """Evaluates the reference model's data likelihood on one cohort.

    ``cohort`` needs: vacs, pcrpos (N, G); per-row ind, gap, antigen (0=N, 1=S), x, od.
    ``dense=True`` uses the reference's (G,G,N) formulation (the parity target and the timed
    "restated reference"); ``dense=False`` uses the O(G N) recurrence (the algorithmically
    fair CPU implementation).
    """

    def __init__(self, cohort, splits=None, ignore_pcrpos=False, dense=True):
        self.G, self.N = cohort.vacs.shape[1], cohort.vacs.shape[0]
        self.v = np.asarray(cohort.vacs, dtype=np.int64).T  # (G, N)  abd.py:413
        self.pcr = np.zeros_like(self.v) if ignore_pcrpos else np.asarray(cohort.pcrpos, dtype=np.int64).T
        self.splits = tuple(splits) if splits else ()
        self.dense = dense
        self.rows = {}
        for a, name in ((0, "n"), (1, "s")):
            m = np.asarray(cohort.antigen) == a
            self.rows[name] = dict(
                x=np.asarray(cohort.x, float)[m], od=np.asarray(cohort.od, float)[m],
                gap=np.asarray(cohort.gap)[m].astype(np.int64), ind=np.asarray(cohort.ind)[m].astype(np.int64),
            )

    # ---- pieces ----
    def constrain(self, i_raw):
        return constrain_infections(np.asarray(i_raw).reshape(self.G, self.N), self.pcr, self.splits)

    def _temp(self, exposure, rho):
        if self.dense:
            return temp_response_dense(exposure, rho)
        return temp_response_scan(exposure, rho)

    def _dtemp(self, exposure, rho):
        if self.dense:
            return dtemp_response_dense(exposure, rho)
        exposure = np.asarray(exposure, dtype=float)
        out = np.empty_like(exposure)
        u = np.zeros(exposure.shape[1])
        du = np.zeros(exposure.shape[1])
        for t in range(exposure.shape[0]):
            du = rho * du + u
            u = rho * u + exposure[t]
            out[t] = du
        return out

    def trajectories(self, th, i, w):
        """(ab_n_mu, ab_s_mu), each (G, N): abd.py:329-341 and 367-391."""
        w = np.asarray(w).astype(float)
        pn = perm_indicator(i)
        ps = perm_indicator(i + self.v)
        t_n = self._temp(i, th["ab_n_rho"])
        rho_ind = th["ab_s_rho"] * w + 1 - w  # abd.py:374
        u_s = self._temp(i, rho_ind) + self._temp(self.v, rho_ind)  # abd.py:378-386; temp unused
        mu_n = th["ab_n_perm"] * pn + th["ab_n_temp"] * t_n + th["ab_n_init"]
        mu_s = th["ab_s_perm"] * ps + u_s + th["ab_s_init"]
        return mu_n, mu_s

    def deterministics(self, theta13, i_raw, w):
        th = dict(zip(THETA13, theta13))
        i = self.constrain(i_raw)
        mu_n, mu_s = self.trajectories(th, i, w)
        return i.astype(np.int8), mu_n, mu_s

    def loglik_rows(self, theta13, i_raw, w):
        """Per-row Normal log-densities (abd.py:459-469), dict antigen -> (R_a,)."""
        th = dict(zip(THETA13, theta13))
        i = self.constrain(i_raw)
        mu = dict(zip("ns", self.trajectories(th, i, w)))
        out = {}
        for a in "ns":
            r = self.rows[a]
            b, d, sg = th[f"it_{a}_b"], th[f"it_{a}_d"], th[f"it_{a}_sigma"]
            pred = logistic(r["x"], mu[a][r["gap"], r["ind"]], b, d)
            z = (r["od"] - pred) / sg
            out[a] = -0.5 * z * z - np.log(np.sqrt(2 * np.pi)) - np.log(sg)
        return out

    def loglik(self, theta13, i_raw, w):
        ll = self.loglik_rows(theta13, i_raw, w)
        return float(ll["n"].sum() + ll["s"].sum())

    def loglik_per_individual(self, theta13, i_raw, w):
        ll = self.loglik_rows(theta13, i_raw, w)
        out = np.zeros(self.N)
        for a in "ns":
            np.add.at(out, self.rows[a]["ind"], ll[a])
        return out

    def loglik_grad(self, theta13, i_raw, w):
        """(loglik, d loglik / d theta13) -- analytic, SURVEY.md section 8a formulas."""
        th = dict(zip(THETA13, theta13))
        i = self.constrain(i_raw)
        wf = np.asarray(w).astype(float)
        pn, ps = perm_indicator(i), perm_indicator(i + self.v)
        rho_ind = th["ab_s_rho"] * wf + 1 - wf
        t_n = self._temp(i, th["ab_n_rho"])
        dt_n = self._dtemp(i, th["ab_n_rho"])
        e_s = i + self.v
        u_s = self._temp(i, rho_ind) + self._temp(self.v, rho_ind)
        du_s = self._dtemp(e_s, rho_ind) * wf
        mu = dict(
            n=th["ab_n_perm"] * pn + th["ab_n_temp"] * t_n + th["ab_n_init"],
            s=th["ab_s_perm"] * ps + u_s + th["ab_s_init"],
        )
        dmu = dict(
            n=dict(ab_n_init=1.0, ab_n_perm=pn, ab_n_temp=t_n, ab_n_rho=th["ab_n_temp"] * dt_n),
            s=dict(ab_s_init=1.0, ab_s_perm=ps, ab_s_rho=du_s),
        )
        total, grad = 0.0, dict.fromkeys(THETA13, 0.0)
        for a in "ns":
            r = self.rows[a]
            b, d, sg = th[f"it_{a}_b"], th[f"it_{a}_d"], th[f"it_{a}_sigma"]
            m = mu[a][r["gap"], r["ind"]]
            s = 1.0 / (1.0 + np.exp(-b * (r["x"] - m)))
            eps = (r["od"] - d * s) / sg
            total += float(np.sum(-0.5 * eps * eps - np.log(np.sqrt(2 * np.pi)) - np.log(sg)))
            grad[f"it_{a}_d"] = float(np.sum(eps / sg * s))
            grad[f"it_{a}_b"] = float(np.sum(eps / sg * d * s * (1 - s) * (r["x"] - m)))
            grad[f"it_{a}_sigma"] = float(np.sum((eps * eps - 1) / sg))
            dm = -(eps / sg) * d * s * (1 - s) * b
            for name, dd in dmu[a].items():
                dd_r = dd if np.isscalar(dd) else dd[r["gap"], r["ind"]]
                grad[name] = float(np.sum(dm * dd_r))
        return total, np.array([grad[k] for k in THETA13])

    # ---- joint density over the 17 unconstrained scalars + the binary variables ----
    def logp_dlogp(self, q17, i_raw, w):
        """Joint model logp in PyMC's unconstrained space and its gradient w.r.t. q17 (what
        ``model.logp_dlogp_function`` returns to NUTS, abd.py:922)."""
        q17 = np.asarray(q17, dtype=float)
        vals, dvals, logj, dlogj = backward(q17)
        prior, dprior = prior_logp(vals, self.G)
        th13 = np.array([vals[k] for k in THETA13])
        ll, dll = self.loglik_grad(th13, i_raw, w)
        k_i, n_i = int(np.count_nonzero(i_raw)), self.G * self.N
        k_w, n_w = int(np.count_nonzero(w)), self.N
        p, pw = vals["p"], vals["ab_s_p_waner"]
        bern = k_i * np.log(p) + (n_i - k_i) * np.log1p(-p) + k_w * np.log(pw) + (n_w - k_w) * np.log1p(-pw)
        g_con = dict(dprior)
        g_con["p"] += k_i / p - (n_i - k_i) / (1 - p)
        g_con["ab_s_p_waner"] += k_w / pw - (n_w - k_w) / (1 - pw)
        for k, gk in zip(THETA13, dll):
            g_con[k] += gk
        logp = ll + bern + sum(prior.values()) + logj
        grad = np.array([g_con[n] * dvals[n] for n in Q17]) + dlogj
        return float(logp), grad

    def logp(self, q17, i_raw, w):
        return self.logp_dlogp(q17, i_raw, w)[0]

    # ---- Gibbs ----
    def cond_logodds(self, theta13, p, p_w, i_raw, w):
        """Brute-force conditional log-odds of every binary variable:
        delta[t, n] = logp(i_raw[t,n]=1, rest) - logp(i_raw[t,n]=0, rest), (G, N), and
        delta_w[n] likewise for ab_s_waner[n].  One full per-individual likelihood per flip."""
        i_raw = np.array(i_raw).reshape(self.G, self.N)
        w = np.array(w)
        base = np.log(p) - np.log1p(-p)
        out = np.empty((self.G, self.N))
        for t in range(self.G):
            hi, lo = i_raw.copy(), i_raw.copy()
            hi[t, :] = 1
            lo[t, :] = 0
            out[t] = base + self.loglik_per_individual(theta13, hi, w) - self.loglik_per_individual(theta13, lo, w)
        base_w = np.log(p_w) - np.log1p(-p_w)
        out_w = base_w + (
            self.loglik_per_individual(theta13, i_raw, np.ones_like(w))
            - self.loglik_per_individual(theta13, i_raw, np.zeros_like(w))
        )
        return out, out_w


# ---------------------------------------------------------------------------------------
# PyMC pieces: transforms and priors
# ---------------------------------------------------------------------------------------
def backward(q17):
    """Unconstrained -> constrained.  Returns (values, d value/d q, sum log|J|, d logJ/d q).
    LogTransform: x = exp(y), log|J| = y.  LogOddsTransform: x = sigmoid(y),
    log|J| = log(sigmoid(y)) + log1p(-sigmoid(y))   (pymc/logprob/transforms.py)."""
    vals, dvals, logj = {}, {}, 0.0
    dlogj = np.zeros(len(VALUE_VARS))
    for k, ((name, tr), y) in enumerate(zip(VALUE_VARS, q17)):
        if tr == "log":
            vals[name] = np.exp(y)
            dvals[name] = vals[name]
            logj += y
            dlogj[k] = 1.0
        elif tr == "logodds":
            s = 1.0 / (1.0 + np.exp(-y))
            oms = 1.0 / (1.0 + np.exp(y))  # 1 - s without cancellation
            vals[name] = s
            dvals[name] = s * oms
            logj += np.log(s) + np.log(oms)
            dlogj[k] = oms - s
        else:
            vals[name] = y
            dvals[name] = 1.0
    return vals, dvals, logj, dlogj


def forward(values):
    """Constrained dict -> unconstrained 17-vector."""
    q = []
    for name, tr in VALUE_VARS:
        x = values[name]
        q.append(np.log(x) if tr == "log" else np.log(x) - np.log1p(-x) if tr == "logodds" else x)
    return np.array(q, dtype=float)


def _beta_lp(x, a, b):
    lp = (a - 1) * np.log(x) if a != 1 else 0.0
    lp += (b - 1) * np.log1p(-x) if b != 1 else 0.0
    d = ((a - 1) / x if a != 1 else 0.0) - ((b - 1) / (1 - x) if b != 1 else 0.0)
    return lp - (gammaln(a) + gammaln(b) - gammaln(a + b)), d


def _gamma_lp(x, mu, sigma):
    a, b = mu * mu / (sigma * sigma), mu / (sigma * sigma)
    return -gammaln(a) + a * np.log(b) - b * x + (a - 1) * np.log(x), -b + (a - 1) / x


def _normal_lp(x, mu, sigma):
    z = (x - mu) / sigma
    return -0.5 * z * z - np.log(np.sqrt(2 * np.pi)) - np.log(sigma), -z / sigma


def prior_logp(vals, n_gaps):
    """Priors of the 17 continuous RVs (constrained space) and their derivatives:
    abd.py:424 (p), :329-340 (N response), :367-388 (S response), :464-467 (sigmoids)."""
    spec = {
        "p": lambda x: _beta_lp(x, 1.0, float(n_gaps - 1)),
        "ab_n_perm": lambda x: _gamma_lp(x, 2.0, 0.5),
        "ab_n_temp": lambda x: _gamma_lp(x, 1.0, 0.5),
        "ab_n_rho": lambda x: _beta_lp(x, 10.0, 1.0),
        "ab_n_init": lambda x: _normal_lp(x, -2.0, 1.0),
        "ab_s_perm": lambda x: _gamma_lp(x, 2.0, 0.5),
        "ab_s_rho": lambda x: _beta_lp(x, 10.0, 1.0),
        "ab_s_p_waner": lambda x: _beta_lp(x, 1.0, 1.0),
        "ab_s_tempinf": lambda x: _gamma_lp(x, 1.0, 0.5),
        "ab_s_tempvac": lambda x: _gamma_lp(x, 1.0, 0.5),
        "ab_s_init": lambda x: _normal_lp(x, -2.0, 1.0),
        "it_n_b": lambda x: _normal_lp(x, -1.0, 0.5),
        "it_n_d": lambda x: _normal_lp(x, 2.0, 0.5),
        "it_n_sigma": lambda x: (-x, -1.0),  # Exponential(1): log(lam) - lam x
        "it_s_b": lambda x: _normal_lp(x, -1.0, 0.5),
        "it_s_d": lambda x: _normal_lp(x, 2.0, 0.5),
        "it_s_sigma": lambda x: (-x, -1.0),
    }
    lp, dlp = {}, {}
    for name, f in spec.items():
        lp[name], dlp[name] = f(vals[name])
    return lp, dlp


def sample_prior(rng, n_gaps):
    """One draw of the 17 continuous RVs from their priors (constrained dict)."""

    def gam(mu, sigma):
        return rng.gamma(mu * mu / sigma**2, sigma**2 / mu)

    return {
        "p": rng.beta(1, n_gaps - 1),
        "ab_n_perm": gam(2, 0.5), "ab_n_temp": gam(1, 0.5), "ab_n_rho": rng.beta(10, 1),
        "ab_n_init": rng.normal(-2, 1),
        "ab_s_perm": gam(2, 0.5), "ab_s_rho": rng.beta(10, 1), "ab_s_p_waner": rng.beta(1, 1),
        "ab_s_tempinf": gam(1, 0.5), "ab_s_tempvac": gam(1, 0.5), "ab_s_init": rng.normal(-2, 1),
        "it_n_b": rng.normal(-1, 0.5), "it_n_d": rng.normal(2, 0.5), "it_n_sigma": rng.exponential(1),
        "it_s_b": rng.normal(-1, 0.5), "it_s_d": rng.normal(2, 0.5), "it_s_sigma": rng.exponential(1),
    }  # fmt: skip


# ---------------------------------------------------------------------------------------
# PyMC's BinaryGibbsMetropolis, restated (pymc/step_methods/metropolis.py; invoked through
# assign_step_methods by abd.py:922).  Pure-Python loop: small cohorts only.
# ---------------------------------------------------------------------------------------
def binary_gibbs_metropolis_sweep(oracle, theta13, p, p_w, i_raw, w, rng, transit_p=0.8):
    """One ``astep``: shuffle all G*N + N binary dims; for each, with probability transit_p
    propose the flip and accept iff log(U) < logp_prop - logp_curr.  Uses the per-individual
    likelihood (exact: every other term cancels).  Returns (i_raw, w) updated copies."""
    G, N = oracle.G, oracle.N
    i_raw = np.array(i_raw).reshape(G, N).copy()
    w = np.array(w).copy()
    lo_i, lo_w = np.log(p) - np.log1p(-p), np.log(p_w) - np.log1p(-p_w)
    order = rng.permutation(G * N + N)
    cur = oracle.loglik_per_individual(theta13, i_raw, w)
    for idx in order:
        if rng.random() >= transit_p:
            continue
        if idx < G * N:
            t, n = divmod(idx, N)
            old = i_raw[t, n]
            i_raw[t, n] = 1 - old
            prior = lo_i if old == 0 else -lo_i
        else:
            n = idx - G * N
            old = w[n]
            w[n] = 1 - old
            prior = lo_w if old == 0 else -lo_w
        new = oracle.loglik_per_individual(theta13, i_raw, w)
        delta = prior + new[n] - cur[n]
        if np.isfinite(delta) and np.log(rng.random()) < delta:
            cur = new
        elif idx < G * N:
            i_raw[t, n] = old
        else:
            w[n] = old
    return i_raw, w


def binary_gibbs_metropolis_sweep_local(subs, G, N, theta13, p, p_w, i_raw, w, rng, transit_p=0.8):
    """The same ``astep`` as ``binary_gibbs_metropolis_sweep`` (same shuffle, same uniforms, same decisions for
    the same ``rng``), evaluating only the flipped bit's individual: ``subs[n]`` is the Oracle of individual n
    alone.  What makes the restated reference sampler affordable on cohorts of a few hundred individuals."""
    i_raw = np.array(i_raw).reshape(G, N).copy()
    w = np.array(w).copy()
    lo_i, lo_w = np.log(p) - np.log1p(-p), np.log(p_w) - np.log1p(-p_w)
    order = rng.permutation(G * N + N)
    cur = np.array([subs[n].loglik(theta13, i_raw[:, n:n + 1], w[n:n + 1]) for n in range(N)])
    for idx in order:
        if rng.random() >= transit_p:
            continue
        if idx < G * N:
            t, n = divmod(idx, N)
            old = i_raw[t, n]
            i_raw[t, n] = 1 - old
            prior = lo_i if old == 0 else -lo_i
        else:
            n = idx - G * N
            old = w[n]
            w[n] = 1 - old
            prior = lo_w if old == 0 else -lo_w
        new = subs[n].loglik(theta13, i_raw[:, n:n + 1], w[n:n + 1])
        delta = prior + new - cur[n]
        if np.isfinite(delta) and np.log(rng.random()) < delta:
            cur[n] = new
        elif idx < G * N:
            i_raw[t, n] = old
        else:
            w[n] = old
    return i_raw, w


# ---------------------------------------------------------------------------------------
# The device sweep, restated: same visiting order, same random numbers, same decisions as
# abd_gibbs_sweep (include/abd_b200.h), so a CUDA sweep can be checked bit-for-bit.
# ---------------------------------------------------------------------------------------
_M32 = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011).  ctr: 4 x u32, key: 2 x u32."""
    c = [int(v) & _M32 for v in ctr]
    k = [int(v) & _M32 for v in key]
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & _M32, p1 & _M32, ((p0 >> 32) ^ c[3] ^ k[1]) & _M32, p0 & _M32]
        k = [(k[0] + 0x9E3779B9) & _M32, (k[1] + 0xBB67AE85) & _M32]
    return c


def _u01(r):
    return (float(r) + 1.0) * 2.0**-32


def _blocked_individual(sub, ll, col, wn, cur, G, splits, p, lo_w, n_global, chain, sweep, key):
    """ABD_GIBBS_BLOCKED for one individual (include/abd_b200.h): per time chunk, the chunk's raw
    bits are redrawn together from their exact full conditional -- a categorical over the position
    of the chunk's first raw 1 (or none), prior p (1-p)^o / (1-p)^len, the bits behind the first 1
    (and every bit of a chunk overridden by PCR+) fresh Bernoulli(p) draws -- then the waner bit
    from its exact conditional.  Random numbers: Philox counter (j, individual, chain, sweep), word
    0 = fresh bit of gap j, word 1 = categorical uniform of chunk j, word 2 (j = 0) = waner."""
    rnd = [philox4x32_10((j, n_global, chain, sweep & _M32), key) for j in range(32)]
    fresh = np.array([1 if _u01(rnd[t][0]) <= p else 0 for t in range(G)], dtype=col.dtype)
    lp1, lp0 = np.log(p), np.log1p(-p)
    pcr = sub.pcr[:, 0]
    n_prop = n_acc = 0
    for k, (a, b) in enumerate(chunk_bounds(G, splits)):
        if a == b:
            continue
        n_prop += 1
        if pcr[a:b].any():
            col[a:b] = fresh[a:b]
            continue
        length = b - a
        lls = np.empty(length + 1)
        for o in range(length + 1):
            cand = col.copy()
            cand[a:b] = 0
            if o < length:
                cand[a + o] = 1
            lls[o] = ll(cand, wn)
        lw = lls + np.concatenate([np.arange(length) * lp0 + lp1, [length * lp0]])
        lw = np.where(np.isnan(lw), -np.inf, lw)
        old_first = int(np.argmax(col[a:b])) if col[a:b].any() else length
        if not np.isfinite(lw.max()):
            pick = length
        else:
            cum = np.cumsum(np.exp(lw - lw.max()))
            pick = int((cum[:length] < _u01(rnd[k][1]) * cum[length]).sum())
        col[a:b] = 0
        if pick < length:
            col[a + pick] = 1
            col[a + pick + 1:b] = fresh[a + pick + 1:b]
        n_acc += pick != old_first
        cur = lls[pick]
    n_prop += 1
    new = ll(col, 1 - wn)
    d10 = lo_w + ((cur - new) if wn else (new - cur))
    w_new = 1 if _u01(rnd[0][2]) <= 1.0 / (1.0 + np.exp(-d10)) else 0
    n_acc += w_new != wn
    return col, w_new, (n_prop, n_acc)


def device_gibbs_sweep(cohort, splits, ignore_pcrpos, theta13, p, p_w, i_raw, w, seed, sweep, chain,
                       mode=0, transit_p=0.8, ind_offset=0):
    """One sweep of one chain exactly as the CUDA kernel schedules it: for every individual n,
    proposals j = 0..G (j < G flips i_raw[j, n], j == G flips waner[n]) are visited in the
    order of their Philox keys; mode 0 = Metropolised flip with probability transit_p
    (BinaryGibbsMetropolis semantics), mode 1 = heat bath.  Returns (i_raw, w, [proposals,
    flips])."""
    G, N = cohort.n_gaps, cohort.n_inds
    i_raw = np.array(i_raw).reshape(G, N).astype(np.int8).copy()
    w = np.array(w).astype(np.int8).copy()
    lo_i, lo_w = np.log(p) - np.log1p(-p), np.log(p_w) - np.log1p(-p_w)
    key = (seed & _M32, ((seed >> 32) ^ (sweep >> 32)) & _M32)
    n_prop = n_flip = 0
    for n in range(N):
        sub = Oracle(cohort.take(np.array([n])), splits=splits, ignore_pcrpos=ignore_pcrpos, dense=False)

        def ll(col, wn):
            return sub.loglik(theta13, col.reshape(G, 1), np.array([wn]))

        col, wn = i_raw[:, n].copy(), int(w[n])
        cur = ll(col, wn)
        if mode == 2:
            col, wn, st = _blocked_individual(sub, ll, col, wn, cur, G, splits, p, lo_w, n + ind_offset, chain, sweep, key)
            n_prop += st[0]
            n_flip += st[1]
            i_raw[:, n], w[n] = col, wn
            continue
        rnd = [philox4x32_10((j, n + ind_offset, chain, sweep & _M32), key) for j in range(G + 1)]
        order = sorted(range(G + 1), key=lambda j: (rnd[j][0], j))
        for j in order:
            u_t, u_a = _u01(rnd[j][1]), _u01(rnd[j][2])
            if mode == 0 and not (u_t <= transit_p):
                continue
            col2, wn2 = col.copy(), wn
            if j == G:
                wn2 = 1 - wn
                cur_bit, lo = wn, lo_w
            else:
                col2[j] = 1 - col[j]
                cur_bit, lo = int(col[j]), lo_i
            new = ll(col2, wn2)
            d10 = (cur - new + lo) if cur_bit else (new - cur + lo)
            if mode == 0:
                delta = -d10 if cur_bit else d10
                flip = bool(np.isfinite(delta) and np.log(u_a) < delta)
            else:
                flip = (u_a <= 1.0 / (1.0 + np.exp(-d10))) != bool(cur_bit)
            n_prop += 1
            if flip:
                n_flip += 1
                col, wn, cur = col2, wn2, new
        i_raw[:, n], w[n] = col, wn
    return i_raw, w, [n_prop, n_flip]
