"""A minimal stand-in for the parts of the PyMC 5 / PyTensor API that abdpymc_b200/abd.py touches.

PyMC is not installable in the offline build image, so the PyMC-facing glue (the PyTensor Ops,
``model()``, the ``GpuBinaryGibbs`` step method) could otherwise never be executed here.  This
module implements just the protocol those pieces are written against -- ``Op.make_node / perform /
grad`` with ``Apply`` nodes, a ``Model`` context with named random variables, value variables with
PyMC's transform suffixes, ``Potential`` / ``Deterministic``, ``BlockedStep`` -- with eager
evaluation of the tiny graphs involved.  It is TEST infrastructure: it says nothing about PyMC's
numerics (those are restated in the oracle), only that our glue calls the library with the right
arguments in the right order.  ``tests/test_pymc_parity.py`` is the test against a real PyMC.
"""
from __future__ import annotations

import enum
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np


# ------------------------------------------------------------------------------------------ graph
class Variable:
    def __init__(self, name=None, dtype="float64", owner=None, index=0, const=None):
        self.name, self.dtype, self.owner, self.index, self.const = name, dtype, owner, index, const

    def astype(self, dtype):
        return self  # values are converted by the Ops themselves

    def __mul__(self, other):
        return _Mul()(self, other)

    __rmul__ = __mul__

    def __repr__(self):
        return f"Variable({self.name or self.owner})"


class Apply:
    def __init__(self, op, inputs, outputs):
        self.op, self.inputs, self.outputs = op, list(inputs), list(outputs)
        for k, o in enumerate(self.outputs):
            o.owner, o.index = self, k


class Op:
    __props__ = ()

    def __call__(self, *inputs):
        node = self.make_node(*inputs)
        return node.outputs[0] if len(node.outputs) == 1 else node.outputs


class _Mul(Op):
    def make_node(self, a, b):
        return Apply(self, [as_tensor_variable(a), as_tensor_variable(b)], [Variable()])

    def perform(self, node, inputs, output_storage):
        output_storage[0][0] = np.asarray(inputs[0]) * np.asarray(inputs[1])


class _Undefined:
    def __init__(self, op, idx, var):
        self.op, self.idx, self.var = op, idx, var


def grad_undefined(op, idx, var, comment=""):
    return _Undefined(op, idx, var)


def as_tensor_variable(v):
    return v if isinstance(v, Variable) else Variable(const=np.asarray(v))


def evaluate(var, point):
    """Value of ``var`` given ``point`` = {free variable name: value}."""
    if var.const is not None:
        return var.const
    if var.owner is None:
        return point[var.name]
    node = var.owner
    inputs = [evaluate(v, point) for v in node.inputs]
    storage = [[None] for _ in node.outputs]
    node.op.perform(node, inputs, storage)
    return storage[var.index][0]


# ------------------------------------------------------------------------------------------ model
_STACK = []


class Model:
    def __init__(self, coords=None):
        self.coords = dict(coords or {})
        self.free_RVs, self.value_vars, self.rvs_to_values = [], [], {}
        self.potentials, self.deterministics, self.named = [], [], {}
        self.dims = {}

    def __enter__(self):
        _STACK.append(self)
        return self

    def __exit__(self, *exc):
        _STACK.pop()

    def __getitem__(self, name):
        return self.named[name]

    def _rv(self, name, transform, dtype, dims):
        rv = Variable(name=name, dtype=dtype)
        value = Variable(name=name + {None: "", "log": "_log__", "logodds": "_logodds__"}[transform], dtype=dtype)
        self.free_RVs.append(rv)
        self.value_vars.append(value)
        self.rvs_to_values[rv] = value
        self.named[name] = rv
        self.dims[name] = dims
        return rv


def modelcontext(model=None):
    if model is not None:
        return model
    if not _STACK:
        raise TypeError("No model on context stack.")
    return _STACK[-1]


def _dist(transform, dtype="float64"):
    def make(name, *args, dims=None, **kwargs):
        return modelcontext()._rv(name, transform, dtype, dims)

    return make


def Potential(name, var):
    m = modelcontext()
    var.name = name
    m.potentials.append(var)
    m.named[name] = var
    return var


def Deterministic(name, var, dims=None):
    m = modelcontext()
    var.name = name
    m.deterministics.append(var)
    m.named[name] = var
    m.dims[name] = dims
    return var


class BlockedStep:
    """PyMC 5's BlockedStep allocates in __new__ and wants the variables there (pymc/step_methods/compound.py)."""

    def __new__(cls, *args, **kwargs):
        vars = kwargs.get("vars", args[0] if args else None)
        if vars is None or len(vars) == 0 or any(v is None for v in vars):
            raise ValueError("No free random variables to sample.")
        return super().__new__(cls)


class Competence(enum.IntEnum):
    INCOMPATIBLE = 0
    COMPATIBLE = 1
    PREFERRED = 2
    IDEAL = 3


# ------------------------------------------------------------------------------------------ install
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def load_abd_with_fake_pymc():
    """Import a private copy of abdpymc_b200/abd.py with the stand-in modules in place of pymc /
    pytensor; sys.modules is restored afterwards (the real abdpymc_b200.abd is left untouched)."""
    pt = _module("pytensor.tensor", as_tensor_variable=as_tensor_variable, dscalar=lambda: Variable(dtype="float64"),
                 dmatrix=lambda: Variable(dtype="float64"), bmatrix=lambda: Variable(dtype="int8"))
    fakes = {
        "pymc": _module("pymc", Model=Model, modelcontext=modelcontext, Potential=Potential, Deterministic=Deterministic,
                        Beta=_dist("logodds"), Gamma=_dist("log"), Exponential=_dist("log"), Normal=_dist(None),
                        Bernoulli=_dist(None, "int64")),
        "pytensor": _module("pytensor", tensor=pt),
        "pytensor.tensor": pt,
        "pytensor.gradient": _module("pytensor.gradient", grad_undefined=grad_undefined),
        "pytensor.graph": _module("pytensor.graph"),
        "pytensor.graph.basic": _module("pytensor.graph.basic", Apply=Apply),
        "pytensor.graph.op": _module("pytensor.graph.op", Op=Op),
        "pymc.step_methods": _module("pymc.step_methods"),
        "pymc.step_methods.arraystep": _module("pymc.step_methods.arraystep", BlockedStep=BlockedStep),
        "pymc.step_methods.compound": _module("pymc.step_methods.compound", Competence=Competence),
    }
    saved = {k: sys.modules.get(k) for k in fakes}
    sys.modules.update(fakes)
    try:
        path = Path(__file__).resolve().parent.parent / "abdpymc_b200" / "abd.py"
        spec = importlib.util.spec_from_file_location("abdpymc_b200._abd_with_fake_pymc", path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        assert mod.HAVE_PYMC
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
