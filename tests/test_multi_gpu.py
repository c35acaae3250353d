"""The real multi-GPU driver on hardware: `Synthesize.static()` under torchrun with one rank per
GPU over NCCL must give exactly the Results of the single-rank run — samples are keyed by their global index, the
candidate gather keeps the reference's CZ-sorted order (cp_utils.py:200) and verification candidates are dealt
round-robin and gathered back.  Needs >= 2 GPUs (skipped on a one-GPU box; the run log of
`gpurun --gpus 2` is kept in profiles/r2_multi_gpu_static.txt)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, out, case):
    script = os.path.join(ROOT, "tools", "static_ranks.py")
    if world == 1:
        cmd = [sys.executable, script, out, "--case", case]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), script, out, "--case", case]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.load(open(out))


@pytest.mark.parametrize("case", ["c2", "c3small"])
def test_static_is_identical_on_one_and_many_ranks(tmp_path, case):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    one = _run(1, str(tmp_path / "one.json"), case)
    for world in sorted({2, min(n, 4)}):
        many = _run(world, str(tmp_path / f"w{world}.json"), case)
        assert many["world"] == world
        for key in ("prospective_cz", "cz_counts", "cz_depths", "digest"):
            assert many[key] == one[key], (world, key)
        assert many["losses"] == one["losses"]
    assert len(one["cz_counts"]) > 0
