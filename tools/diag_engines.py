import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipy.stats import unitary_group
from cpflow_b200.ansatz import Ansatz
from cpflow_b200.engine import Loss, Penalty
from cpflow_b200.penalty import RegularizationOptions, make_regularization_function
from cpflow_b200.topology import chain_layer, connected_layer, fill_layers
PF = make_regularization_function(RegularizationOptions)
pen = lambda r=0.01: Penalty("piecewise", r, PF.segments, PF.period)
for n, layer, K, rg in [(4, chain_layer(4), 40, "xyz"), (4, [[0, 1], [0, 2], [0, 3]], 11, "xz"),
                        (3, connected_layer(3), 7, "xyz"), (5, connected_layer(5), 13, "xyz"), (4, connected_layer(4), 9, "zx")]:
    anz = Ansatz(n, "cp", fill_layers(layer, K), rg)
    V = unitary_group.rvs(2 ** n, random_state=2)
    for dt in (torch.float32, torch.float64):
        a = torch.tensor(np.random.default_rng(K).uniform(0, 6.28, (21, anz.num_angles)), dtype=dt, device="cuda")
        outs = []
        for flag in ("0", "1"):
            os.environ["CPF_NO_LAYERED"] = flag
            lo, rg_, gr = anz.program.loss_grad(a, Loss("hs", V), pen())
            hist = []
            st = anz.program.adam_state(a.clone())
            for t in range(7):
                anz.program.adam_run(st, Loss("hs", V), pen(), 0.1, 1)
                hist.append(st.angles.clone())
            outs.append((lo.clone(), gr.clone(), st.best_regloss.clone(), hist))
        os.environ["CPF_NO_LAYERED"] = "0"
        d = [float((x - y).abs().max()) for x, y in zip(outs[0][:3], outs[1][:3])]
        dh = [float((x - y).abs().max()) for x, y in zip(outs[0][3], outs[1][3])]
        gmin = float(outs[1][1].abs().min())
        print(n, K, rg, dt, "loss %.2e grad %.2e best %.2e" % tuple(d), "angles/step", " ".join("%.1e" % v for v in dh), "min|g| %.1e" % gmin)
