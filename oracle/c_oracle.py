"""ctypes front of the plain-C oracle (oracle/abd_oracle_c.c) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__`` (build / smoke) and ``bench.py``'s CPU-baseline leg may import this
module.  ``COracle`` is ``abd_oracle.Oracle`` with the data log-likelihood and its gradient (the part that
costs time: constraints, titer recurrences, OD rows) computed by the compiled, OpenMP-threaded C
restatement; priors, transforms and the Bernoulli terms stay the NumPy oracle's.  It is checked against the
same reference-generated goldens as the NumPy oracle (tests/test_oracle.py).
"""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from . import abd_oracle as ora

HERE = Path(__file__).resolve().parent
SRC = HERE / "abd_oracle_c.c"
OUT = HERE / "_build" / "libabd_oracle_c.so"
_lib = None


def build(force: bool = False) -> Path:
    """gcc -O2 -fopenmp: compiles the C restatement in-tree (oracle/_build/, git-ignored, travels with gpurun)."""
    if not force and OUT.exists() and OUT.stat().st_mtime >= SRC.stat().st_mtime:
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", str(OUT), str(SRC), "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{res.stdout}\n{res.stderr}")
    return OUT


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        lib = C.CDLL(str(build()))
        vp, i32, i64 = C.c_void_p, C.c_int, C.c_long
        lib.abd_c_loglik_grad.restype = C.c_int
        lib.abd_c_loglik_grad.argtypes = [i32, i64, i32, vp, vp, vp, i64, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp,
                                          vp, i32]
        lib.abd_c_gibbs_sweep.restype = C.c_int
        lib.abd_c_gibbs_sweep.argtypes = [i32, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_double, C.c_double,
                                          vp, vp, C.c_uint64, C.c_uint64, C.c_uint32, i32, C.c_double, i64, vp, i32]
        lib.abd_c_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class COracle(ora.Oracle):
    """The NumPy oracle with `loglik_grad` (hence `logp_dlogp`) computed by the C restatement."""

    def __init__(self, cohort, splits=None, ignore_pcrpos=False, threads=0):
        super().__init__(cohort, splits=splits, ignore_pcrpos=ignore_pcrpos, dense=False)
        self._lib = load()
        self.threads = int(threads)
        self._pcr8 = np.ascontiguousarray(self.pcr, dtype=np.int8)
        self._vac8 = np.ascontiguousarray(self.v, dtype=np.int8)
        self._splits = np.asarray(self.splits, dtype=np.int32)
        self._rows = {}
        for a in "ns":
            r = self.rows[a]
            self._rows[a] = (np.ascontiguousarray(r["x"], dtype=np.float64), np.ascontiguousarray(r["od"], dtype=np.float64),
                             np.ascontiguousarray(r["gap"], dtype=np.int32), np.ascontiguousarray(r["ind"], dtype=np.int32))
        self._scratch = np.empty(7 * self.G * self.N, dtype=np.float64)
        self._out = np.empty(14, dtype=np.float64)

    def loglik_grad(self, theta13, i_raw, w):
        th = np.ascontiguousarray(theta13, dtype=np.float64)
        i8 = np.ascontiguousarray(np.asarray(i_raw).reshape(self.G, self.N) != 0, dtype=np.int8)
        w8 = np.ascontiguousarray(np.asarray(w) != 0, dtype=np.int8)
        (xn, odn, gn, in_), (xs, ods, gs, is_) = self._rows["n"], self._rows["s"]
        rc = self._lib.abd_c_loglik_grad(self.G, self.N, len(self._splits), _p(self._splits), _p(self._pcr8), _p(self._vac8),
                                         xn.size, _p(xn), _p(odn), _p(gn), _p(in_), xs.size, _p(xs), _p(ods), _p(gs), _p(is_),
                                         _p(th), _p(i8), _p(w8), _p(self._out), _p(self._scratch), self.threads)
        if rc:
            raise ValueError("abd_c_loglik_grad: invalid sizes")
        return float(self._out[0]), self._out[1:].copy()

    def _csr(self):
        """OD rows sorted by individual + row pointers per antigen (built on first use)."""
        if not hasattr(self, "_csr_rows"):
            self._csr_rows = {}
            for a in "ns":
                x, od, gap, ind = self._rows[a]
                order = np.argsort(ind, kind="stable")
                ptr = np.concatenate(([0], np.cumsum(np.bincount(ind, minlength=self.N)))).astype(np.int64)
                self._csr_rows[a] = (ptr, np.ascontiguousarray(x[order]), np.ascontiguousarray(od[order]),
                                     np.ascontiguousarray(gap[order]))
        return self._csr_rows

    def gibbs_sweep(self, theta13, p, p_w, i_raw, w, seed, sweep, chain, mode=0, transit_p=0.8, ind_offset=0):
        """One sweep of one chain as the CUDA kernel schedules it (same visiting order, random numbers and
        decisions as ``abd_oracle.device_gibbs_sweep`` and ``abd_gibbs_sweep``); modes 0 (Metropolis) and 1 (heat
        bath).  Returns (i_raw, w, [proposals, flips])."""
        th = np.ascontiguousarray(theta13, dtype=np.float64)
        i8 = np.ascontiguousarray(np.asarray(i_raw).reshape(self.G, self.N) != 0, dtype=np.int8).copy()
        w8 = np.ascontiguousarray(np.asarray(w) != 0, dtype=np.int8).copy()
        (pn, xn, odn, gn), (ps, xs, ods, gs) = self._csr()["n"], self._csr()["s"]
        stats = np.zeros(2, dtype=np.int64)
        rc = self._lib.abd_c_gibbs_sweep(self.G, self.N, len(self._splits), _p(self._splits), _p(self._pcr8), _p(self._vac8),
                                         _p(pn), _p(xn), _p(odn), _p(gn), _p(ps), _p(xs), _p(ods), _p(gs), _p(th), float(p),
                                         float(p_w), _p(i8), _p(w8), int(seed), int(sweep), int(chain), int(mode),
                                         float(transit_p), int(ind_offset), _p(stats), self.threads)
        if rc:
            raise ValueError("abd_c_gibbs_sweep: invalid sizes or mode")
        return i8, w8, [int(stats[0]), int(stats[1])]
