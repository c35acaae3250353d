"""Exploration: prospective fraction / min CZ / score of our engine vs the reference's stored trials."""
import json, math, sys
import numpy as np, torch
sys.path.insert(0, '.')
import cpflow_b200 as cp
from cpflow_b200.gates import u_toff3, u_toff4
from cpflow_b200.topology import num_qubits_from_layer
t = json.load(open('tests/golden/trials.json'))
def score(cz, n): return -math.log2(sum(2.0 ** (-c) for c in cz) / n) if cz else float('inf')
for f, B_ref in [('tutorial/results/toff4_star', 500), ('paper/results/toff4_star_xyz', 1000), ('paper/results/toff4_chain_xyz', 1000),
                 ('paper/results/toff3_conn_xyz', 200), ('paper/results/toff3_chain_xyz', 200)]:
    rec = t[f]
    n = num_qubits_from_layer(rec['layer'])
    syn = cp.Synthesize(rec['layer'], target_unitary=u_toff3 if n == 3 else u_toff4)
    rows = [r for r in rec['trials'] if isinstance(r['cz_counts'], list)]
    rows = sorted(rows, key=lambda r: -len(r['cz_counts']))
    pick = rows[:3] + rows[len(rows)//2:len(rows)//2+3]
    for r in pick:
        B = 4 * B_ref
        o = cp.StaticOptions(num_cp_gates=r['num_cp_gates'], r=r['r'], accepted_num_cz_gates=10**6, num_samples=B, random_seed=r['random_seed'])
        anz, cand = syn._prospective(o)
        cz = [int(c) for c in cand[:, 1].tolist()]
        print(f"{f:34s} k={r['num_cp_gates']:3d} r={r['r']:.6f} ref: p={len(r['cz_counts'])/B_ref:.3f} min={min(r['cz_counts']) if r['cz_counts'] else None} score={r['score']:.2f} | ours: p={len(cz)/B:.3f} min={min(cz) if cz else None} score={score(cz,B):.2f}", flush=True)
