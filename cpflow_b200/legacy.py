"""Read result files written by the reference (idnm/cpflow) into this package's `Results`.

The reference saves `Results` objects with dill (main.py:448-456); the files name cpflow, qiskit, hyperopt and
jax classes and carry pickled lambdas (the loss closure).  None of those packages is needed here and nothing
in the file is executed: the reader below resolves every global outside an explicit (module, name) allow-list of
plain data constructors (list / dict / complex ..., OrderedDict, numpy's array reconstructors) to an inert
placeholder that only records constructor arguments and state, then copies the plain data out.  In particular
`builtins.eval / exec / __import__ / getattr`, `functools.partial`, `copyreg`, `os.*` all become placeholders, so a
crafted file cannot run code through this reader:

* every stored `Decomposition` -> `Decomposition` of this package: gate list -> `circuit.Circuit`, stored
  unitary / loss / type / CZ and T counts taken as stored (nothing is re-evaluated, so no GPU is needed);
* hyperopt `Trials` -> `main.Trials` (`results` = the stored objective dicts: num_cp_gates, r, random_seed,
  loss, cz_counts), so `Results.best_hyperparameters()` and a resumed `Synthesize.adaptive` see the old trials;
* the stored options objects -> `StaticOptions` / `AdaptiveOptions` where the fields match.

The loss function itself is a pickled closure and is not recovered: `Results.loss_function` is None unless the
stored decomposer carries a target unitary, in which case it is the Hilbert-Schmidt `Loss('hs', target)`.

    from cpflow_b200.legacy import load_reference_results
    res = load_reference_results('/path/to/cpflow/tutorial/results/toff3_chain')
    res.decompositions[0].circuit.qasm()
"""
import dataclasses
import io
import pickle

import numpy as np

# Explicit allow-list of globals a reference result file may really resolve: plain data constructors only.
_ALLOWED = {
    ("builtins", n) for n in ("list", "dict", "set", "frozenset", "tuple", "complex", "bytearray", "bytes", "slice",
                              "range", "int", "float", "bool", "str", "object")
} | {("collections", "OrderedDict"), ("collections", "defaultdict"), ("collections", "deque"),
     ("fractions", "Fraction"), ("_codecs", "encode"),
     ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy", "float32"), ("numpy", "float64"), ("numpy", "int32"),
     ("numpy", "int64"), ("numpy", "complex64"), ("numpy", "complex128"), ("numpy", "bool_"),
     ("numpy._core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "scalar"),
     ("numpy._core.numeric", "_frombuffer")}


class _Inert:
    """Placeholder for a class that is not importable here: records arguments and state, runs nothing."""
    _path = "?"

    def __init__(self, *args, **kwargs):
        self._args, self._kwargs = args, kwargs

    def __setstate__(self, state):
        self._state = state

    def append(self, x):                      # list / dict opcodes applied to container subclasses
        self.__dict__.setdefault("_items", []).append(x)

    def extend(self, xs):
        self.__dict__.setdefault("_items", []).extend(xs)

    def __setitem__(self, k, v):
        self.__dict__.setdefault("_map", {})[k] = v

    def __repr__(self):
        return f"<inert {self._path}>"


class _Reader(pickle.Unpickler):
    def __init__(self, f):
        super().__init__(f)
        self._classes = {}

    def find_class(self, module, name):
        if module.startswith("numpy.core"):                      # numpy 1 paths in the stored files
            module = module.replace("numpy.core", "numpy._core", 1)
        if (module, name) in _ALLOWED:
            return super().find_class(module, name)
        if module.startswith("dill"):
            # dill's reconstruction helpers (_create_function, _create_cell, _load_type, _get_attr, ...) are
            # CALLED by the pickle: return a factory that yields a fresh inert class holding the call's arguments
            def helper(*args, **kwargs):
                cls = type(f"dill_{name}", (_Inert,), {"_path": f"{module}.{name}"})
                cls._call = (args, kwargs)
                return cls
            return helper
        key = (module, name)
        if key not in self._classes:
            self._classes[key] = type(name, (_Inert,), {"_path": f"{module}.{name}", "__module__": module})
        return self._classes[key]


def _state(obj):
    st = getattr(obj, "_state", None)
    return st if isinstance(st, dict) else {}


def _array(x):
    """numpy array from a stored numpy array or a pickled jax DeviceArray (state = (_, shape, dtype, fortran, raw))."""
    if isinstance(x, np.ndarray):
        return x
    st = getattr(x, "_state", None)
    if st is None and getattr(x, "_args", None):
        st = x._args
    if isinstance(st, tuple) and len(st) >= 5:
        _, shape, dtype, fortran, raw = st[:5]
        return np.frombuffer(raw, dtype=dtype).reshape(shape, order="F" if fortran else "C").copy()
    raise TypeError(f"cannot decode an array from {x!r}")


def _scalar(x):
    if x is None or isinstance(x, (int, float)):
        return x
    if isinstance(x, np.generic):
        return float(x)
    try:
        return float(_array(x).reshape(()))
    except Exception:
        return None


def _circuit(qc):
    """qiskit QuantumCircuit placeholder -> circuit.Circuit (rz / rx / ry / cz / cx / cp / h gates)."""
    from .circuit import Circuit
    st = _state(qc)
    qubits = st["_qubits"]
    index = {id(q): i for i, q in enumerate(qubits)}
    gp = st.get("_global_phase", 0.0)
    try:
        gp = float(gp)
    except (TypeError, ValueError):
        gp = 0.0
    out = Circuit(len(qubits), global_phase=gp)
    for entry in st["_data"]:
        gate, qargs = entry[0], entry[1]
        gs = _state(gate)
        name = gs["_name"]
        params = [float(p) for p in (gs.get("_params") or [])]
        out.append(name, [index[id(q)] for q in qargs], params)
    return out


def _options(opt):
    """Stored StaticOptions / AdaptiveOptions -> this package's dataclass (unknown fields are dropped)."""
    from . import main as M
    if opt is None:
        return None
    st = _state(opt)
    cls = {"StaticOptions": M.StaticOptions, "AdaptiveOptions": M.AdaptiveOptions}.get(type(opt).__name__)
    if cls is None or not st:
        return None
    names = {f.name for f in dataclasses.fields(cls)}
    kw = {k: v for k, v in st.items() if k in names and isinstance(v, (int, float, str, bool, type(None)))}
    try:
        return cls(**kw)
    except TypeError:
        return None


def _decomposition(d, loss_function):
    from .circuit import gates_count, gates_depth
    from .main import Decomposition
    ds = _state(d)
    out = object.__new__(Decomposition)          # stored values are kept: nothing is re-evaluated on a GPU
    out.unitary_loss_func = loss_function
    out.circuit = _circuit(ds["circuit"])
    out.unitary = np.asarray(_array(ds["unitary"]), dtype=np.complex128)
    out.label = ds.get("label", "")
    out.loss = _scalar(ds.get("loss"))
    out.type = ds.get("type", "Approximate")
    out.cz_count = ds.get("cz_count", gates_count(["cz"], out.circuit))
    out.cz_depth = ds.get("cz_depth", gates_depth(["cz"], out.circuit))
    out.t_count, out.t_depth = ds.get("t_count"), ds.get("t_depth")
    out._cp_data = None                          # closures of the reference's jax functions: not recoverable
    out._static_options = _options(ds.get("_static_options"))
    out._adaptive_options = _options(ds.get("_adaptive_options"))
    out._decomposer = None
    return out


def _target_unitary(decs):
    for d in decs:
        dec = _state(d).get("_decomposer")
        tu = _state(dec).get("target_unitary") if dec is not None else None
        if tu is not None:
            try:
                return np.asarray(_array(tu), dtype=np.complex128)
            except TypeError:
                continue
    return None


def _plain(v):
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    return v


def load_reference_results(path):
    """Reference `Results` file -> `cpflow_b200.main.Results` (see the module docstring)."""
    from .engine import Loss
    from .main import Results, Trials
    with open(path, "rb") as f:
        raw = _Reader(io.BytesIO(f.read())).load()
    st = _state(raw)
    if "decompositions" not in st or "layer" not in st:
        raise ValueError(f"{path}: not a cpflow Results file")
    stored = list(st["decompositions"])
    target = _target_unitary(stored)
    loss_function = Loss("hs", target) if target is not None else None
    trials = None
    tr = st.get("trials")
    if tr is not None and "_trials" in _state(tr):
        trials = Trials()
        for t in _state(tr)["_trials"]:
            res = t["result"]
            trials.results.append({k: _plain(v) for k, v in res.items()})
    res = Results(loss_function, [list(p) for p in st["layer"]], label=st.get("label", ""), trials=trials,
                  decompositions=[_decomposition(d, loss_function) for d in stored],
                  save_to=st.get("save_to", ""))
    return res
