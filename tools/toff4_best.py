"""Full Synthesize.static() runs on the 4-qubit Toffoli for the five topologies of paper/CPFlow.tex:480 (best known CZ
counts: connected 14, kite 14, square 16, star 16, chain 18), at hyper-parameter points the reference's own stored
trials found productive (tests/golden/trials.json).  Prints prospective / verified counts.

    python tools/toff4_best.py [num_samples] > profiles/toff4_best_r2.txt
"""
import contextlib
import io
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cpflow_b200 as cp  # noqa: E402
from cpflow_b200.gates import u_toff4  # noqa: E402
from cpflow_b200.topology import chain_layer, connected_layer  # noqa: E402

CASES = [("star", [[0, 1], [0, 2], [0, 3]], 26, 0.000793, 16),
         ("chain", chain_layer(4), 26, 0.000256, 18),
         ("chain", chain_layer(4), 34, 0.000443, 18),
         ("kite", [[0, 1], [1, 2], [2, 3], [1, 3]], 25, 0.000648, 14),
         ("square", [[0, 1], [1, 2], [2, 3], [3, 0]], 24, 0.000528, 16),
         ("connected", connected_layer(4), 23, 0.000528, 14)]

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    print(f"# Toffoli-4, static(), {B} samples per case, seed 0, complex64; known best: paper/CPFlow.tex:480")
    for name, layer, K, r, best in CASES:
        syn = cp.Synthesize(layer, target_unitary=u_toff4, label=f"toff4_{name}")
        opts = cp.StaticOptions(num_cp_gates=K, r=r, accepted_num_cz_gates=best + 1, num_samples=B)
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            res = syn.static(opts, save_results=False)
        dt = time.perf_counter() - t0
        cz = sorted(d.cz_count for d in res.decompositions)
        pro = syn.last_prospective_cz_counts
        hist = {c: cz.count(c) for c in sorted(set(cz))}
        print(f"{name:10s} K={K:2d} r={r:.6f} known_best={best:2d} | prospective(cz<={best + 1})={len(pro):5d} "
              f"verified={len(cz):5d} min_cz={cz[0] if cz else None} hist={hist} max_loss="
              f"{max((d.loss for d in res.decompositions), default=None)} wall={dt:.1f}s", flush=True)
