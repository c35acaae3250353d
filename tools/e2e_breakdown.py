"""Where the host-API (e2e) time of one C3 step goes: stage timings with a synchronize after each."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpflow_b200.ansatz import Ansatz
from cpflow_b200.engine import Loss, Penalty
from cpflow_b200.gates import u_toff4
from cpflow_b200.optimization import ProgramLoss, mynimize_repeated, run_adam_batch, _as_device_batch
from cpflow_b200.penalty import RegularizationOptions, make_regularization_function
from cpflow_b200.topology import chain_layer, fill_layers

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
anz = Ansatz(4, "cp", fill_layers(chain_layer(4), 40)); prog = anz.program
pf = make_regularization_function(RegularizationOptions); pen = Penalty("piecewise", 0.001476, pf.segments, pf.period)
loss = Loss("hs", u_toff4)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
a0 = prog.initial_angles(0, B); a0_host = a0.cpu().pin_memory()
pl = ProgramLoss(prog, loss)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = sync()
    init = _as_device_batch(a0_host, torch.float32, "cuda"); t1 = sync()
    raw = run_adam_batch(prog, loss, pen, init, 0.1, T); t2 = sync()
    r = raw.numpy(); t3 = sync()
    print(f"rep {rep}: h2d {1e3*(t1-t0):.1f} ms  run {1e3*(t2-t1):.1f} ms  d2h+numpy {1e3*(t3-t2):.1f} ms")
    t0 = sync(); res = mynimize_repeated(pl, anz.num_angles, learning_rate=0.1, num_iterations=T, initial_params_batch=a0_host,
                            regularization_func=pen, keep_history=False); t1 = sync()
    print(f"   mynimize_repeated total {1e3*(t1-t0):.1f} ms")
