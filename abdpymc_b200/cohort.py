"""Cohort containers: the arrays the inference hot path consumes, and the synthetic cohorts
the benchmarks run on.

``CohortArrays`` is the flat, NumPy-only view of what the reference keeps in ``TiterData``
(abd.py:46-221): the (n_inds, n_gaps) ``vacs`` / ``pcrpos`` 0/1 matrices (abd.py:112-115) and,
per OD row, ``individual_i``, ``elapsed_months`` (the gather indices of abd.py:35-36),
``log_dilution`` and ``od`` (abd.py:462, 468) plus which antigen the row belongs to
(abd.py:82-98).  It is what ``abd_create`` uploads to the GPU once.

Synthetic cohorts follow SURVEY.md section 8d: the reference's ``simulation.Cohort``
(simulation.py:282-369) copies the *schedule* of an existing cohort and has no size knob, so
larger cohorts are bootstrapped (individuals drawn with replacement from the bundled
1520-individual cohort) and then forward-simulated with the reference's dynamics.
``simulate`` is a vectorised restatement of ``Individual.infection_responses``
(simulation.py:222-279) and ``Cohort.simulate_row`` (simulation.py:328-353).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

DATA_DIR = Path(__file__).resolve().parent / "data"
ANTIGEN_N, ANTIGEN_S = 0, 1
#: measurement codes in df.csv (abd.py:85, 94)
MEASUREMENT_S, MEASUREMENT_N = "10222020-S", "40588-V08B"


def splits_from_t0(t0: str, delta: bool, omicron: bool) -> tuple:
    """abd.py:204-221: gaps from t0 ("YYYY-MM") to 2021-07 (delta) and 2022-01 (omicron)."""
    y, m = (int(v) for v in str(t0).split("-")[:2])
    out = []
    if delta:
        out.append((2021 - y) * 12 + (7 - m))
    if omicron:
        out.append((2022 - y) * 12 + (1 - m))
    return tuple(out)


@dataclass
class CohortArrays:
    vacs: np.ndarray  # (N, G) uint8
    pcrpos: np.ndarray  # (N, G) uint8
    ind: np.ndarray  # (R,) int32   individual_i
    gap: np.ndarray  # (R,) int32   elapsed_months
    antigen: np.ndarray  # (R,) uint8   0 = N, 1 = S
    x: np.ndarray  # (R,) float64 log_dilution
    od: np.ndarray  # (R,) float64
    t0: str = "2020-05"
    truth: dict = field(default_factory=dict)  # simulated ground truth, if any

    def __post_init__(self):
        # The reference uses vacs / pcrpos NUMERICALLY (exposure = i + v, abd.py:368,384; i_raw + pcrpos, abd.py:643);
        # the device layout is one bit per (individual, gap), so anything but 0 / 1 would silently change the titers.
        for name in ("vacs", "pcrpos"):
            a = np.asarray(getattr(self, name))
            if a.size and not np.isin(a, (0, 1)).all():
                raise ValueError(f"{name} must hold only 0 / 1 (found values outside {{0, 1}}); "
                                 "several doses or positives in one gap are not supported by the bit-mask layout")
        self.vacs = np.ascontiguousarray(self.vacs, dtype=np.uint8)
        self.pcrpos = np.ascontiguousarray(self.pcrpos, dtype=np.uint8)
        if self.vacs.shape != self.pcrpos.shape:
            raise ValueError("vacs and pcrpos are different shapes")  # abd.py:196-197
        self.ind = np.ascontiguousarray(self.ind, dtype=np.int32)
        self.gap = np.ascontiguousarray(self.gap, dtype=np.int32)
        self.antigen = np.ascontiguousarray(self.antigen, dtype=np.uint8)
        self.x = np.ascontiguousarray(self.x, dtype=np.float64)
        self.od = np.ascontiguousarray(self.od, dtype=np.float64)
        r = len(self.ind)
        if not (len(self.gap) == len(self.antigen) == len(self.x) == len(self.od) == r):
            raise ValueError("row arrays have different lengths")
        if r and (self.ind.min() < 0 or self.ind.max() >= self.n_inds):
            raise ValueError("individual_i out of range")
        if r and (self.gap.min() < 0 or self.gap.max() >= self.n_gaps):
            raise ValueError("elapsed_months out of range")

    # ---- sizes -------------------------------------------------------------------------
    @property
    def n_inds(self) -> int:
        return int(self.vacs.shape[0])

    @property
    def n_gaps(self) -> int:
        return int(self.vacs.shape[1])

    @property
    def n_rows(self) -> int:
        return int(len(self.ind))

    def rows(self, antigen: int):
        """(x, od, gap, ind) of one antigen's rows, in file order (abd.py:82-98)."""
        m = self.antigen == antigen
        return self.x[m], self.od[m], self.gap[m], self.ind[m]

    # ---- construction ------------------------------------------------------------------
    @classmethod
    def load(cls, name_or_path="cohort") -> "CohortArrays":
        """Load a packed fixture: 'cohort' (1520 x 31), 'test_cohort' (10 x 26) or a path."""
        path = Path(name_or_path)
        if not path.suffix:
            path = DATA_DIR / f"{name_or_path}.npz"
        z = np.load(path)
        return cls(
            vacs=z["vacs"], pcrpos=z["pcrpos"], ind=z["ind"], gap=z["gap"],
            antigen=z["antigen"], x=z["x"], od=z["od"], t0=str(z["t0"]),
        )

    def save(self, path) -> None:
        np.savez_compressed(
            path, t0=np.array(self.t0), vacs=self.vacs, pcrpos=self.pcrpos, ind=self.ind,
            gap=self.gap, antigen=self.antigen, x=self.x, od=self.od,
        )

    @classmethod
    def from_disk(cls, directory) -> "CohortArrays":
        """Read df.csv / vacs.txt / pcrpos.txt / t0.txt the way TiterData.from_disk does
        (abd.py:171-202)."""
        import pandas as pd

        d = Path(directory)
        df = pd.read_csv(d / "df.csv", index_col=0)
        return cls.from_frame(
            df, np.loadtxt(d / "vacs.txt"), np.loadtxt(d / "pcrpos.txt"),
            (d / "t0.txt").read_text().strip(),
        )

    @classmethod
    def from_frame(cls, df, vacs, pcrpos, t0="2020-05") -> "CohortArrays":
        meas = df["measurement"].to_numpy()
        known = (meas == MEASUREMENT_S) | (meas == MEASUREMENT_N)
        df = df[known]
        return cls(
            vacs=np.atleast_2d(vacs), pcrpos=np.atleast_2d(pcrpos),
            ind=df["individual_i"].to_numpy(), gap=df["elapsed_months"].to_numpy(),
            antigen=(df["measurement"].to_numpy() == MEASUREMENT_S).astype(np.uint8),
            x=df["log_dilution"].to_numpy(float), od=df["od"].to_numpy(float), t0=str(t0),
        )

    def to_disk(self, directory) -> None:
        """Write df.csv / vacs.txt / pcrpos.txt / t0.txt in the layout TiterData.to_disk uses
        (abd.py:149-169), with the columns TiterData.__init__ reads (abd.py:75-126)."""
        import pandas as pd

        d = Path(directory)
        d.mkdir(parents=True, exist_ok=True)
        meas = np.where(self.antigen == ANTIGEN_S, MEASUREMENT_S, MEASUREMENT_N)
        sample = np.array([f"S{i:06d}-G{g:02d}" for i, g in zip(self.ind, self.gap)])
        df = pd.DataFrame(dict(
            sample=sample, measurement=meas, od=self.od, dilution=40.0 * 4.0**self.x, record_id=self.ind + 1000,
            elapsed_months=self.gap, individual_i=self.ind, log_dilution=self.x,
            sample_i=pd.factorize(sample)[0],
        ))
        df.to_csv(d / "df.csv")
        np.savetxt(d / "vacs.txt", self.vacs, fmt="%1.0f")
        np.savetxt(d / "pcrpos.txt", self.pcrpos, fmt="%1.0f")
        (d / "t0.txt").write_text(self.t0)

    def calculate_splits(self, delta: bool, omicron: bool) -> tuple:
        """abd.py:204-221: gaps from t0 to 2021-07 (delta) and 2022-01 (omicron)."""
        return splits_from_t0(self.t0, delta, omicron)

    # ---- re-sampling / sharding --------------------------------------------------------
    def take(self, individuals: np.ndarray) -> "CohortArrays":
        """New cohort whose individual k is a copy of ``individuals[k]`` (with its vacs,
        pcrpos and OD-row schedule); rows stay sorted by individual."""
        individuals = np.asarray(individuals, dtype=np.int64)
        order = np.argsort(self.ind, kind="stable")
        counts = np.bincount(self.ind, minlength=self.n_inds)
        starts = np.concatenate(([0], np.cumsum(counts)))[:-1]
        c = counts[individuals]
        new_ind = np.repeat(np.arange(len(individuals)), c)
        # position within each copied individual's block
        within = np.arange(c.sum()) - np.repeat(np.cumsum(c) - c, c)
        src = order[np.repeat(starts[individuals], c) + within]
        truth = {k: v[individuals] for k, v in self.truth.items() if getattr(v, "shape", (0,))[0] == self.n_inds}
        return CohortArrays(
            vacs=self.vacs[individuals], pcrpos=self.pcrpos[individuals], ind=new_ind,
            gap=self.gap[src], antigen=self.antigen[src], x=self.x[src], od=self.od[src],
            t0=self.t0, truth=truth,
        )

    def bootstrap(self, n_inds: int, seed: int = 20240518) -> "CohortArrays":
        """N individuals with replacement, default_rng(20240518)."""
        rng = np.random.default_rng(seed)
        return self.take(rng.integers(0, self.n_inds, size=n_inds))

    def shard(self, rank: int, world: int) -> "CohortArrays":
        """Contiguous block of individuals for rank ``rank`` of ``world``."""
        lo, hi = shard_bounds(self.n_inds, rank, world)
        return self.take(np.arange(lo, hi))


def shard_bounds(n: int, rank: int, world: int) -> tuple:
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


# ----------------------------------------------------------------------------------------
# forward simulation (vectorised restatement of simulation.py)
# ----------------------------------------------------------------------------------------
@dataclass
class SimParams:
    """Defaults of simulation.py:43-44 (Protection), :76-78 (Elisa), :104-108 (Dynamics);
    the same values for the S and the N antibody (simulation.py:162-173)."""

    init: float = -2.0
    perm_rise: float = 2.0
    temp_rise_i: float = 1.5
    temp_rise_v: float = 2.0
    temp_wane: float = 0.95
    elisa_b: float = -2.2
    elisa_d: float = 1.6
    elisa_sd: float = 0.1
    prot_a: float = 0.0
    prot_b: float = 1.0


def simulate(cohort: CohortArrays, lam0=0.04, seed: int = 42, params: SimParams | None = None) -> CohortArrays:
    """Forward-simulate infections, titers and OD readings on ``cohort``'s schedule.

    Dynamics follow simulation.py:242-277 month by month: an individual is infected when PCR+,
    or when exposed (w.p. lam0[t]) and protected by neither antibody (logistic protection on
    last month's titer, simulation.py:46-56, 175-184); temp response
    ``prev*wane + inf*rise_i + vac*rise_v`` (S) / ``prev*wane + inf*rise_i`` (N)
    (simulation.py:117-132); permanent rise after the first infection (or vaccination, S only)
    (simulation.py:134-152).  OD = Normal(logistic(log_dilution, titer, b, d), sd)
    (simulation.py:80-90; abd.py:556-557).  The RNG stream is NumPy ``default_rng`` and so
    differs from the reference's legacy global ``np.random`` stream; the dynamics are pinned
    against the reference in tests/test_cohort.py.
    """
    pr = params or SimParams()
    rng = np.random.default_rng(seed)
    n, g = cohort.n_inds, cohort.n_gaps
    lam0 = np.broadcast_to(np.asarray(lam0, dtype=float), (g,))
    inf = np.zeros((n, g), dtype=np.uint8)
    s_t = np.empty((n, g))
    n_t = np.empty((n, g))
    s_temp = np.zeros(n)
    n_temp = np.zeros(n)
    ever_inf = np.zeros(n, dtype=bool)
    ever_vac = np.zeros(n, dtype=bool)
    s_prev = np.full(n, pr.init)
    n_prev = np.full(n, pr.init)

    def p_prot(t):
        return 1.0 / (1.0 + np.exp(-pr.prot_b * (t - pr.prot_a)))

    for t in range(g):
        exposed = rng.uniform(size=n) < lam0[t]
        protected = (rng.uniform(size=n) < p_prot(s_prev)) | (rng.uniform(size=n) < p_prot(n_prev))
        infected = (cohort.pcrpos[:, t] == 1) | (exposed & ~protected)
        vacc = cohort.vacs[:, t] == 1
        inf[:, t] = infected
        s_temp = s_temp * pr.temp_wane + infected * pr.temp_rise_i + vacc * pr.temp_rise_v
        n_temp = n_temp * pr.temp_wane + infected * pr.temp_rise_i
        ever_inf |= infected
        ever_vac |= vacc
        s_prev = pr.init + s_temp + pr.perm_rise * (ever_inf | ever_vac)
        n_prev = pr.init + n_temp + pr.perm_rise * ever_inf
        s_t[:, t] = s_prev
        n_t[:, t] = n_prev

    titer = np.where(cohort.antigen == ANTIGEN_S, s_t[cohort.ind, cohort.gap], n_t[cohort.ind, cohort.gap])
    mean = pr.elisa_d / (1.0 + np.exp(-pr.elisa_b * (cohort.x - titer)))
    od = rng.normal(mean, pr.elisa_sd)
    return CohortArrays(
        vacs=cohort.vacs, pcrpos=cohort.pcrpos, ind=cohort.ind, gap=cohort.gap,
        antigen=cohort.antigen, x=cohort.x, od=od, t0=cohort.t0,
        truth=dict(infections=inf, s_titer=s_t, n_titer=n_t),
    )


def synthetic_cohort(n_inds: int, seed: int = 20240518, sim_seed: int = 42, lam0: float = 0.04) -> CohortArrays:
    """The benchmark cohort of SURVEY.md section 8d: bootstrap the bundled 1520-individual
    schedule to ``n_inds`` individuals, then simulate with the reference's default
    ``Antibodies()`` and lam0 = 0.04 (test_simulation.py:384)."""
    return simulate(CohortArrays.load("cohort").bootstrap(n_inds, seed=seed), lam0=lam0, seed=sim_seed)
