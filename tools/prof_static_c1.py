import cProfile, pstats, time, numpy as np, sys
sys.path.insert(0, '/root/repo')
import cpflow_b200 as cp
from cpflow_b200.topology import chain_layer
ccz = np.diag([1, 1, 1, 1, 1, 1, 1, -1]).astype(complex)
def run(label):
    syn = cp.Synthesize(chain_layer(3), target_unitary=ccz, label=label)
    return syn.static(cp.StaticOptions(num_cp_gates=12, accepted_num_cz_gates=10, num_samples=10))
import tempfile, os
os.chdir(tempfile.mkdtemp())
run('warm')
t0 = time.perf_counter(); run('a'); print('second call', time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); run('b'); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
