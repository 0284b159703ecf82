"""NumPy-backed stand-ins for ``pytensor`` / ``pymc`` / ``arviz`` so that the REFERENCE's own
``abdpymc/abd.py`` can be executed in this container, where PyMC is not installed.

TEST INFRASTRUCTURE ONLY.  This file is used by ``make_golden.py`` (run once, in the build
container, where ``/root/reference`` exists) to produce the committed fixtures in this
directory.  Nothing in the product path, the ``-m gpu`` tests or ``bench.py`` imports it.

What is and is not "the reference" when running under this shim
---------------------------------------------------------------
* Every line of ``abd.py`` that builds the model graph -- ``mask_multiple_infections*``,
  ``incorporate_pcrpos``, ``mask_three_gaps``/``mask_future_infection``, ``perm_response``,
  ``_temp_response_{scalar,vector}_rho`` (the dense (G,G,N) formulation), ``model_n_response``,
  ``model_s_response``, ``model_sigmoids``, ``logistic``, ``model`` -- is the reference's code,
  executed eagerly on ndarrays instead of symbolically.
* The tensor primitives (``at.where``, ``at.cumsum``, ``pt.scan`` ...) are NumPy equivalents of
  the PyTensor ops of the same name, and the distribution log-densities (``pm.Beta`` ...)
  are restated from PyMC's documented formulas (pymc/distributions/continuous.py, discrete.py;
  transforms in pymc/logprob/transforms.py).  Those restated pieces are validated by running
  the reference's own unit tests (``test_abd.py``) under the shim -- see ``make_golden.py``.

The shim evaluates a model *at a point*: ``with evaluate(point) as ev: abd.model(...)`` runs
the reference's ``model()`` with every free RV replaced by its value in ``point`` and collects
per-RV log-densities in ``ev.logp`` and Deterministics in ``ev.det``.  Values may be complex
(complex-step differentiation of the reference code).
"""

from __future__ import annotations

import importlib.util
import sys
import types
from contextlib import contextmanager
from pathlib import Path

import numpy as np
from scipy.special import gammaln

REFERENCE_ROOT = Path("/root/reference")


# --------------------------------------------------------------------------------------
# tensor stand-in
# --------------------------------------------------------------------------------------
class _Shape(tuple):
    def eval(self):
        return np.array(tuple(self))


class T(np.ndarray):
    """ndarray with ``.eval()`` and ``.shape.eval()`` like a PyTensor variable."""

    def eval(self):
        return np.asarray(self)

    @property
    def shape(self):
        return _Shape(np.ndarray.shape.__get__(self))

    def __array_wrap__(self, arr, context=None, return_scalar=False):
        return np.asarray(arr).view(T)

    def __getitem__(self, key):
        return np.asarray(np.ndarray.__getitem__(self, key)).view(T)

    def any(self, *a, **k):
        return _t(np.asarray(self).any(*a, **k))

    def sum(self, *a, **k):
        return _t(np.asarray(self).sum(*a, **k))

    def cumsum(self, *a, **k):
        return _t(np.asarray(self).cumsum(*a, **k))


def _t(x, dtype=None):
    return np.asarray(x, dtype=dtype).view(T)


def _make_tensor_module():
    at = types.ModuleType("pytensor.tensor")
    at.TensorLike = object
    at.as_tensor = at.as_tensor_variable = lambda x, *a, **k: _t(x)
    at.arange = lambda *a, **k: _t(np.arange(*a, **k))
    at.maximum = lambda a, b: _t(np.maximum(a, b))
    at.tril = lambda a, k=0: _t(np.tril(a, k))
    at.ones_like = lambda a: _t(np.ones_like(a))
    at.zeros_like = lambda a: _t(np.zeros_like(a))
    at.zeros = lambda shape, dtype="float64": _t(np.zeros(shape, dtype=dtype))
    at.cast = lambda a, dtype: _t(np.asarray(a).astype(dtype))
    at.cumsum = lambda a, axis=None: _t(np.cumsum(a, axis=axis))
    at.where = at.switch = lambda c, a, b: _t(np.where(c, a, b))
    at.concatenate = lambda arrs, axis=0: _t(np.concatenate([np.asarray(a) for a in arrs], axis=axis))
    at.exp = lambda a: _t(np.exp(a))
    at.log = lambda a: _t(np.log(a))
    return at


def _scan(fn, sequences=None, outputs_info=None, non_sequences=None, **_):
    """Eager ``pytensor.scan`` for the two call patterns in abd.py (abd.py:287-292, 572-576):
    argument order is sequences, then output taps oldest-first, then non_sequences."""
    if isinstance(sequences, dict):
        seqs = [np.asarray(sequences["input"])]
    else:
        seqs = [np.asarray(s) for s in (sequences if isinstance(sequences, (list, tuple)) else [sequences])]
    non_sequences = list(non_sequences or [])
    if isinstance(outputs_info, dict):
        taps = list(outputs_info["taps"])
        hist = [np.asarray(h) for h in np.asarray(outputs_info["initial"])]  # oldest first
    else:
        taps = [-1]
        hist = [np.asarray(outputs_info)]
    depth = max(-t for t in taps)
    assert len(hist) == depth
    out = []
    for k in range(len(seqs[0])):
        prev = [hist[len(hist) + t] for t in taps]
        val = np.asarray(fn(*[_t(s[k]) for s in seqs], *[_t(p) for p in prev], *non_sequences))
        hist.append(val)
        out.append(val)
    return _t(np.stack(out)), {}


# --------------------------------------------------------------------------------------
# pymc stand-in: evaluate the model at a point
# --------------------------------------------------------------------------------------
class _Eval:
    def __init__(self, point):
        self.point = point
        self.logp = {}
        self.det = {}
        self.coords = None

    @property
    def total(self):
        return sum(self.logp.values())


_CTX: list[_Eval] = []


@contextmanager
def evaluate(point):
    ev = _Eval(point)
    _CTX.append(ev)
    try:
        yield ev
    finally:
        _CTX.pop()


def _free(name):
    return _CTX[-1].point[name]


def _record(name, terms):
    _CTX[-1].logp[name] = np.sum(terms)


class _Model:
    def __init__(self, coords=None):
        if _CTX:
            _CTX[-1].coords = coords

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _beta(name, alpha, beta, **_):
    # pymc.distributions.continuous.Beta.logp
    x = _free(name)
    a, b = float(alpha), float(beta)
    lp = (0.0 if a == 1 else (a - 1) * np.log(x)) + (0.0 if b == 1 else (b - 1) * np.log1p(-x))
    _record(name, lp - (gammaln(a) + gammaln(b) - gammaln(a + b)))
    return _t(x)


def _gamma(name, mu, sigma, **_):
    # pymc Gamma(mu, sigma): alpha = mu^2/sigma^2, beta = mu/sigma^2
    x = _free(name)
    a, b = mu**2 / sigma**2, mu / sigma**2
    _record(name, -gammaln(a) + a * np.log(b) - b * x + (a - 1) * np.log(x))
    return _t(x)


def _normal(name, mu=0.0, sigma=1.0, observed=None, **_):
    x = np.asarray(observed) if observed is not None else _free(name)
    z = (x - np.asarray(mu)) / sigma
    _record(name, -0.5 * z * z - np.log(np.sqrt(2 * np.pi)) - np.log(sigma))
    return _t(x)


def _exponential(name, lam, **_):
    x = _free(name)
    _record(name, np.log(lam) - lam * x)
    return _t(x)


def _bernoulli(name, p, dims=None, **_):
    x = np.asarray(_free(name))
    _record(name, np.where(x != 0, np.log(p), np.log1p(-p)))
    return _t(x)


def _deterministic(name, value, dims=None):
    if _CTX:
        _CTX[-1].det[name] = np.array(value)
    return value


def install():
    """Insert the stand-in modules into ``sys.modules`` and load the reference's abd.py
    (and simulation.py) from where they lie.  Returns the loaded ``abd`` module."""
    if "abdpymc" in sys.modules and getattr(sys.modules["abdpymc"], "_shimmed", False):
        return sys.modules["abdpymc"].abd

    at = _make_tensor_module()
    pt = types.ModuleType("pytensor")
    pt.tensor = at
    pt.scan = _scan
    pm = types.ModuleType("pymc")
    pm.Model = _Model
    pm.Beta, pm.Gamma, pm.Normal = _beta, _gamma, _normal
    pm.Exponential, pm.Bernoulli, pm.Deterministic = _exponential, _bernoulli, _deterministic
    az = types.ModuleType("arviz")
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.update(
        {"pytensor": pt, "pytensor.tensor": at, "pymc": pm, "arviz": az, "matplotlib": mpl, "matplotlib.pyplot": plt}
    )

    def load(modname, fname):
        spec = importlib.util.spec_from_file_location(modname, REFERENCE_ROOT / "abdpymc" / fname)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    abd = load("abdpymc.abd", "abd.py")
    pkg = types.ModuleType("abdpymc")
    pkg.__file__ = str(REFERENCE_ROOT / "abdpymc" / "__init__.py")
    pkg._shimmed = True
    pkg.abd = abd
    for k, v in vars(abd).items():
        if not k.startswith("__"):
            setattr(pkg, k, v)
    sys.modules["abdpymc"] = pkg
    pkg.simulation = load("abdpymc.simulation", "simulation.py")
    return abd


# --------------------------------------------------------------------------------------
# PyMC's default transforms for the model's 17 continuous RVs (pymc/logprob/transforms.py):
#   LogOddsTransform: x = sigmoid(y), log|J| = log(sigmoid(y)) + log1p(-sigmoid(y))
#   LogTransform:     x = exp(y),     log|J| = y
# --------------------------------------------------------------------------------------
VALUE_VARS = [
    ("p", "logodds"),
    ("ab_n_perm", "log"),
    ("ab_n_temp", "log"),
    ("ab_n_rho", "logodds"),
    ("ab_n_init", None),
    ("ab_s_perm", "log"),
    ("ab_s_rho", "logodds"),
    ("ab_s_p_waner", "logodds"),
    ("ab_s_tempinf", "log"),
    ("ab_s_tempvac", "log"),
    ("ab_s_init", None),
    ("it_n_b", None),
    ("it_n_d", None),
    ("it_n_sigma", "log"),
    ("it_s_b", None),
    ("it_s_d", None),
    ("it_s_sigma", "log"),
]


def backward(q):
    """unconstrained 17-vector -> (dict of constrained values, sum of log|J|)."""
    vals, logj = {}, 0.0
    for (name, tr), y in zip(VALUE_VARS, q):
        if tr == "log":
            vals[name] = np.exp(y)
            logj = logj + y
        elif tr == "logodds":
            # log(sigmoid(y)) + log1p(-sigmoid(y)), with 1 - sigmoid(y) evaluated as sigmoid(-y)
            # (PyTensor's log1msigm -> -softplus stabilisation; matters only for |y| > ~15)
            s = 1.0 / (1.0 + np.exp(-y))
            vals[name] = s
            logj = logj + np.log(s) + np.log(1.0 / (1.0 + np.exp(y)))
        else:
            vals[name] = y
    return vals, logj
