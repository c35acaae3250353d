"""Stub unpickler for cpflow's dill result files.

The reference's stored `Results` pickles name cpflow/qiskit/hyperopt/jax classes that are
not installed here.  This loader replaces every unknown class by a recording stub so that
the plain data (angles, unitaries, gate lists, hyperopt trial results) can be read with the
standard library + numpy only.  Test infrastructure: used by `make_golden.py` (run once in
the build container where /root/reference exists) and by `cpflow_b200.legacy` is NOT
allowed to import this file.
"""
import io
import pickle

import numpy as np

_PASS = {"builtins", "collections", "copyreg", "_codecs", "functools", "fractions", "datetime"}
_DILL_HELPERS = {
    "_create_function", "_create_code", "_create_cell", "_load_type", "_get_attr",
    "_create_type", "_import_module", "_create_namedtuple", "_create_weakref",
    "_create_weakproxy", "_eval_repr", "_create_array", "_create_dtypemeta",
    "_create_lock", "_create_rlock", "_create_filehandle", "_create_stringi",
    "_create_stringo", "_getattr", "_dict_from_dictproxy", "_setattr", "_shims",
}


class Stub:
    """Recording placeholder for an unavailable class."""
    _stub_name = "?"

    def __init__(self, *args, **kwargs):
        self._args = args
        self._kwargs = kwargs

    def __setstate__(self, state):
        self._state = state

    def __repr__(self):
        return f"<Stub {self._stub_name}>"

    # hyperopt.Trials and friends get list/dict opcodes applied
    def append(self, x):
        self.__dict__.setdefault("_items", []).append(x)

    def extend(self, xs):
        self.__dict__.setdefault("_items", []).extend(xs)

    def __setitem__(self, k, v):
        self.__dict__.setdefault("_map", {})[k] = v


def _make_stub_class(module, name):
    return type(name, (Stub,), {"_stub_name": f"{module}.{name}", "__module__": module})


def _make_helper(name):
    def helper(*args, **kwargs):
        cls = type(f"dill_{name}", (Stub,), {"_stub_name": f"dill.{name}"})
        cls._fn = (args, kwargs)
        return cls
    helper.__name__ = name
    return helper


class StubUnpickler(pickle.Unpickler):
    _cache = {}

    def find_class(self, module, name):
        if module in _PASS:
            return super().find_class(module, name)
        if module.startswith("numpy"):
            module = module.replace("numpy.core", "numpy._core")
            return super().find_class(module, name)
        if module.startswith("dill") and name in _DILL_HELPERS:
            return _make_helper(name)
        key = (module, name)
        if key not in self._cache:
            self._cache[key] = _make_stub_class(module, name)
        return self._cache[key]


def load(path):
    with open(path, "rb") as f:
        return StubUnpickler(io.BytesIO(f.read())).load()


def jax_array(stub):
    """Decode a pickled jax DeviceArray (Appendix A: `_get_attr` stub with raw bytes state)."""
    if isinstance(stub, np.ndarray):
        return stub
    st = getattr(stub, "_state", None)
    if st is None and hasattr(stub, "_args") and stub._args:
        st = stub._args
    if isinstance(st, tuple) and len(st) >= 5:
        _, shape, dtype, fortran, raw = st[:5]
        a = np.frombuffer(raw, dtype=dtype).reshape(shape, order="F" if fortran else "C")
        return a.copy()
    raise TypeError(f"cannot decode array from {stub!r}")
