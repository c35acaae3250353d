"""Writes tests/golden/adam_c3.npz: outputs of the CPU oracle's Adam loop (oracle/cpflow_oracle.py:
adam_minimize_batched, restating optimization.py:28-94) on the C3 shape the bench times — 4 qubits, K = 40,
chain and star layers, 32 samples, complex128, T = 150 — plus the verification variant (projected CP angles frozen,
no penalty).  The GPU parity test compares the fused kernel with these arrays; tests/test_oracle_golden.py re-derives
one case from the oracle so the fixture cannot drift from it.

    python tests/golden/make_adam_c3.py          (about two minutes on 8 cores)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import parity_lib as P  # noqa: E402
from oracle import cpflow_oracle as O  # noqa: E402

CASES = [("chain", O.chain_layer(4), False), ("star", P.STAR4, False), ("chain", O.chain_layer(4), True),
         ("star", P.STAR4, True)]
B, T, K = 32, 150, 40

if __name__ == "__main__":
    out = {}
    tgt = O.toffoli_target(4).numpy()
    for name, layer, freeze in CASES:
        res = P.oracle_adam_case(4, layer, K, tgt, B, T, freeze)
        for k, v in res.items():
            out[f"{name}_{int(freeze)}_{k}"] = v
        print(name, freeze, res["best_regloss"][:4])
    np.savez_compressed(os.path.join(HERE, "adam_c3.npz"), **out)
