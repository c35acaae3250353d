"""Small C3 workload for ncu: 4q Toffoli, chain, K=40, one adam_run launch."""
import sys
import torch
sys.path.insert(0, '.')
from cpflow_b200.ansatz import Ansatz
from cpflow_b200.topology import fill_layers, chain_layer
from cpflow_b200.engine import Loss, Penalty
from cpflow_b200.penalty import make_regularization_function, RegularizationOptions
from cpflow_b200.gates import u_toff4
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
pf = make_regularization_function(RegularizationOptions)
anz = Ansatz(4, 'cp', fill_layers(chain_layer(4), 40))
prog = anz.program
st = prog.adam_state(prog.initial_angles(0, B))
prog.adam_run(st, Loss('hs', u_toff4), Penalty('piecewise', 0.001476, pf.segments, pf.period), 0.1, T)
torch.cuda.synchronize()
print('done', float(st.best_regloss.mean()))
