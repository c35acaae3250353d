"""Developer script: first GPU parity check of the engine against the oracle."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle import cpflow_oracle as O
from cpflow_b200.ansatz import Ansatz
from cpflow_b200.topology import fill_layers, chain_layer, connected_layer
from cpflow_b200.engine import Loss, Penalty
from cpflow_b200.penalty import make_regularization_function, RegularizationOptions
from scipy.stats import unitary_group

torch.manual_seed(0)
dev = 'cuda'
pf = make_regularization_function(RegularizationOptions)
def rel(a, b): return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))
for n, layer, K, rg in [(3, chain_layer(3), 5, 'xyz'), (4, [[0,1],[0,2],[0,3]], 10, 'xyz'), (2, [[0,1]], 3, 'xz'),
                        (5, connected_layer(5), 12, 'xyz'), (4, [[3,1],[2,0]], 7, 'zyx'), (4, chain_layer(4), 40, 'xyz')]:
    anz = Ansatz(n, 'cp', fill_layers(layer, K), rg)
    oanz = O.cp_ansatz(layer, K, rg); ops = O.ansatz_program(oanz)
    assert ops == [tuple(o) for o in anz.ops]
    P = anz.num_angles; N = 2 ** n; B = 37
    rng = np.random.default_rng(n)
    a64 = rng.uniform(0, 2 * np.pi, (B, P))
    V = unitary_group.rvs(N, random_state=1)
    for dt, tol in [(torch.float64, 1e-12), (torch.float32, 2e-5)]:
        a = torch.tensor(a64, dtype=dt, device=dev)
        u = anz.program.unitary(a).cpu().numpy()
        uo = O.program_unitary_batched(n, ops, torch.tensor(a64))[:].numpy()
        print(n, rg, K, dt, 'unitary err', np.abs(u - uo).max())
        for kind, tgt in [('hs', V), ('relphase', V), ('state', V[:, 0].copy())]:
            pen = Penalty('piecewise', 0.01, pf.segments, pf.period)
            lo, rg_, gr = anz.program.loss_grad(a, Loss(kind, tgt), pen)
            ol, orr, og = O.loss_and_grad_batched(n, ops, torch.tensor(a.cpu().numpy().astype(np.float64)), kind, torch.tensor(tgt),
                                                  oanz.cp_mask, 0.01, O.make_regularization_function())
            e1 = rel(lo.cpu().numpy(), ol.numpy()); e2 = rel(rg_.cpu().numpy(), orr.numpy())
            gn = np.linalg.norm(gr.cpu().numpy() - og.numpy(), axis=1) / np.linalg.norm(og.numpy(), axis=1)
            print('   ', kind, 'loss', e1, 'reg', e2, 'grad(normwise max)', gn.max(), 'OK' if max(e1, e2, gn.max()) < tol else 'FAIL')
        # cotangent
        cot = torch.tensor(unitary_group.rvs(N, random_state=2)[None].repeat(B, 0), dtype={torch.float32: torch.complex64, torch.float64: torch.complex128}[dt], device=dev).contiguous()
        g = anz.program.adjoint_from_cotangent(a, cot).cpu().numpy()
        at = torch.tensor(a.cpu().numpy().astype(np.float64), requires_grad=True)
        U = O.program_unitary_batched(n, ops, at)
        # L = 2 Re sum conj(cot) U  => dL/dconj(U) = cot
        (2 * (torch.tensor(cot.cpu().numpy().astype(np.complex128)).conj() * U).real.sum()).backward()
        print('    cotangent', rel(g, at.grad.numpy()))
