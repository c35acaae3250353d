"""Developer script: Adam-loop parity vs oracle + first throughput number."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle import cpflow_oracle as O
from cpflow_b200.ansatz import Ansatz
from cpflow_b200.topology import fill_layers, chain_layer, connected_layer
from cpflow_b200.engine import Loss, Penalty
from cpflow_b200.penalty import make_regularization_function, RegularizationOptions
from cpflow_b200.gates import u_toff3, u_toff4
dev = 'cuda'
pf = make_regularization_function(RegularizationOptions)
# ---- parity of the full loop ----
for dt in (torch.float64, torch.float32):
    n, layer, K = 3, chain_layer(3), 6
    anz = Ansatz(n, 'cp', fill_layers(layer, K)); oanz = O.cp_ansatz(layer, K); ops = O.ansatz_program(oanz)
    B, T = 8, 60
    a0 = torch.tensor(np.random.default_rng(0).uniform(0, 2*np.pi, (B, anz.num_angles)), dtype=dt)
    res = O.mynimize_repeated(n, ops, 'hs', torch.tensor(u_toff3), a0, 0.1, T, oanz.cp_mask, 0.002, O.make_regularization_function())
    for chunks in ([T], [1, 9, 50]):
        st = anz.program.adam_state(a0.to(dev).clone())
        for c in chunks:
            anz.program.adam_run(st, Loss('hs', u_toff3), Penalty('piecewise', 0.002, pf.segments, pf.period), 0.1, c)
        torch.cuda.synchronize()
        br = st.best_regloss.cpu().numpy(); obr = np.array([r['regloss'][1].item() for r in res])
        ir = st.init_regloss.cpu().numpy(); oir = np.array([r['regloss'][0].item() for r in res])
        bp = st.best_params.cpu().numpy(); obp = np.stack([r['params'][1].numpy() for r in res])
        print(dt, chunks, 'init', np.abs(ir-oir).max(), 'best', np.abs(br-obr).max(), 'best_params', np.abs(bp-obp).max(), 'reg', np.abs(st.best_reg.cpu().numpy()-np.array([r['reg'][1].item() for r in res])).max())
    # history mode
    resh = O.mynimize_repeated(n, ops, 'hs', torch.tensor(u_toff3), a0, 0.1, T, oanz.cp_mask, 0.002, O.make_regularization_function(), keep_history=True)
    st = anz.program.adam_state(a0.to(dev).clone(), hist_len=T)
    anz.program.adam_run(st, Loss('hs', u_toff3), Penalty('piecewise', 0.002, pf.segments, pf.period), 0.1, T)
    hp = st.hist_params.cpu().numpy(); hl = st.hist_regloss.cpu().numpy()
    print('   history params', np.abs(hp - np.stack([r['params'].numpy() for r in resh])).max(), 'regloss', np.abs(hl - np.stack([r['regloss'].numpy() for r in resh])).max())
# ---- throughput: 4q Toffoli, K=40 ----
for layer_name, layer in [('chain', chain_layer(4)), ('star', [[0,1],[0,2],[0,3]])]:
    anz = Ansatz(4, 'cp', fill_layers(layer, 40))
    prog = anz.program
    flops, byts = prog.eval_cost()
    for B in (16*148*2, 16*148*4, 12500):
        a0 = prog.initial_angles(0, B)
        st = prog.adam_state(a0)
        pen = Penalty('piecewise', 0.001476, pf.segments, pf.period); loss = Loss('hs', u_toff4)
        prog.adam_run(st, loss, pen, 0.1, 20)
        torch.cuda.synchronize()
        T = 200
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); prog.adam_run(st, loss, pen, 0.1, T); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ev = B * T / (ms * 1e-3)
        print(layer_name, 'B', B, 'ms', round(ms, 2), 'evals/s %.3e' % ev, 'TFLOP/s(alg) %.2f' % (ev * flops / 1e12), 'frac of 70.96: %.3f' % (ev * flops / 70.96e12))
