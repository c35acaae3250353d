"""Synthesize.static() under the launcher's world size (1 process, or `torch.distributed.run --nproc-per-node N`, one
GPU per rank, NCCL) -> a JSON summary of the Results on rank 0.  tests/test_multi_gpu.py runs it with 1 and N ranks and
checks that sharding the samples over GPUs changes nothing (parallel.gather_rows / gather_round_robin on device
tensors).

    python tools/static_ranks.py OUT.json [--case c2|c3small]
"""
import argparse
import contextlib
import hashlib
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--case", default="c2")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        saved = os.dup(1)           # NCCL prints its version on stdout at communicator creation
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.all_reduce(torch.zeros(1, device="cuda"))
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    import cpflow_b200 as cp
    from cpflow_b200.gates import u_toff3, u_toff4
    from cpflow_b200.topology import chain_layer, connected_layer
    if a.case == "c2":      # BASELINE configs[1]: Toffoli-3, all-to-all, 10^4 samples
        syn = cp.Synthesize(connected_layer(3), target_unitary=u_toff3, label="t3")
        opts = cp.StaticOptions(num_cp_gates=7, r=0.00131, accepted_num_cz_gates=6, num_samples=10000)
    else:                   # a slice of configs[2]: Toffoli-4 chain, K = 40
        syn = cp.Synthesize(chain_layer(4), target_unitary=u_toff4, label="t4")
        opts = cp.StaticOptions(num_cp_gates=40, r=0.001476, accepted_num_cz_gates=23, num_samples=20001)
    with contextlib.redirect_stdout(io.StringIO()):
        syn.static(opts, save_results=False)                       # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = syn.static(opts, save_results=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    if rank == 0:
        decs = res.decompositions
        h = hashlib.sha256()
        for d in decs:
            h.update(np.ascontiguousarray(d.unitary).tobytes())
            h.update(np.ascontiguousarray(d._cp_data[2]).tobytes())
        json.dump({"world": world, "wall_s": dt, "prospective_cz": syn.last_prospective_cz_counts,
                   "cz_counts": [d.cz_count for d in decs], "cz_depths": [d.cz_depth for d in decs],
                   "losses": [d.loss for d in decs], "digest": h.hexdigest()}, open(a.out, "w"))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
