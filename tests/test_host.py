"""CPU-side checks: the C-ABI library loads and exports every symbol include/cpflow_b200.h
declares, host logic (program compilation, topology, ansatz layout, penalty table) agrees with
the oracle.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import cpflow_oracle as O
from cpflow_b200 import _lib as L
from cpflow_b200 import ansatz as A
from cpflow_b200 import topology as T
from cpflow_b200 import penalty as PN
from cpflow_b200 import gates as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "cpflow_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cpf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    syms = header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/cpflow_b200.h but not exported"
    assert set(syms) == set(L.EXPORTS), "ctypes table and header disagree"
    assert lib.cpf_version() == 200


def test_struct_layouts_match_header():
    assert C.sizeof(L.CpfOp) == 24
    assert C.sizeof(L.CpfProgramInfo) == 32
    assert C.sizeof(L.CpfAdamSpec) == 32
    assert C.sizeof(L.CpfPenaltySpec) == 8 + 16 + 4 * 8 * L.MAX_SEGMENTS + 8
    assert C.sizeof(L.CpfAdamBuffers) == 14 * 8
    assert C.sizeof(L.CpfLaunchInfo) == 8 * 4 + 3 * 8


def _create(lib, n, ops, P):
    arr = (L.CpfOp * max(1, len(ops)))(*[L.CpfOp(*o) for o in ops])
    h = C.c_void_p()
    rc = lib.cpf_program_create(n, len(ops), arr, P, C.byref(h))
    return rc, h


def test_program_create_validation(lib):
    rc, h = _create(lib, 3, [(L.RX, 0, -1, 0, 0.0), (L.CP, 0, 1, 1, 0.0)], 2)
    assert rc == 0 and h.value
    info = L.CpfProgramInfo()
    assert lib.cpf_program_get_info(h, C.byref(info)) == 0
    assert (info.n_qubits, info.n_params, info.n_ops, info.n_rotations, info.n_phase) == (3, 2, 2, 1, 1)
    f, b = C.c_double(), C.c_double()
    assert lib.cpf_eval_cost(h, L.LOSS_HS, L.F32, C.byref(f), C.byref(b)) == 0
    assert f.value == 8 * 8 * (16 * 1 + 4 * 1 + 8) and b.value == 6 * 2 * 4 + 8
    lib.cpf_program_destroy(h)
    # errors: too many qubits, parameter reused, bad pair, bad kind, out-of-range param
    for n, ops, P, code in [
        (9, [(L.RX, 0, -1, 0, 0.0)], 1, -2),
        (3, [(L.RX, 0, -1, 0, 0.0), (L.RY, 1, -1, 0, 0.0)], 1, -1),
        (3, [(L.CP, 1, 1, 0, 0.0)], 1, -1),
        (3, [(17, 0, -1, 0, 0.0)], 1, -1),
        (3, [(L.RX, 0, -1, 5, 0.0)], 1, -1),
        (3, [(L.CZ, 0, 1, 0, 0.0)], 1, -1),
    ]:
        rc, h = _create(lib, n, ops, P)
        assert rc == code, (ops, rc)
        assert not h.value
        assert len(lib.cpf_last_error()) > 0


def test_fusion_counts_for_cp_ansatz(lib):
    """Surface rounds fuse to n SU(2) gates and every block to 2 (main.py:77-80, 122-124)."""
    for n, layer, K, rg in [(4, T.chain_layer(4), 40, "xyz"), (3, T.connected_layer(3), 7, "xz"),
                            (5, T.connected_layer(5), 23, "xyz")]:
        anz = A.Ansatz(n, "cp", T.fill_layers(layer, K), rg)
        rc, h = _create(lib, n, anz.ops, anz.num_angles)
        assert rc == 0
        info = L.CpfProgramInfo()
        lib.cpf_program_get_info(h, C.byref(info))
        assert info.n_rotations == 3 * n + 2 * len(rg) * K
        assert info.n_phase == K
        assert info.n_fused == n + 2 * K
        assert info.n_sched == n + 3 * K
        lib.cpf_program_destroy(h)


@pytest.mark.parametrize("n,layer,K,rg", [(3, [[0, 1], [1, 2]], 12, "xyz"), (4, [[0, 1], [0, 2], [0, 3]], 40, "xyz"),
                                           (4, [[3, 1], [2, 0]], 7, "zyx"), (5, T.connected_layer(5), 13, "xz")])
def test_ansatz_layout_matches_oracle(n, layer, K, rg):
    anz = A.Ansatz(n, "cp", T.fill_layers(layer, K), rg)
    oanz = O.cp_ansatz(layer, K, rg)
    assert anz.num_angles == oanz.num_angles == 3 * n + (2 * len(rg) + 1) * K
    assert np.array_equal(anz.cp_mask, oanz.cp_mask)
    assert [tuple(o) for o in anz.ops] == O.ansatz_program(oanz)


def test_cz_ansatz_layout():
    anz = A.Ansatz(3, "cz", T.fill_layers(T.chain_layer(3), 5), "xyz")
    oanz = O.Ansatz(3, "cz", O.fill_layers(O.chain_layer(3), 5), "xyz")
    assert anz.num_angles == oanz.num_angles == 9 + 6 * 5
    assert [tuple(o) for o in anz.ops] == O.ansatz_program(oanz)
    with pytest.raises(TypeError):
        A.Ansatz(3, "iswap", T.fill_layers(T.chain_layer(3), 5))


def test_topology_matches_reference_semantics():
    assert T.connected_layer(3) == [[0, 1], [0, 2], [1, 2]]
    assert T.chain_layer(4) == [[0, 1], [1, 2], [2, 3]]
    assert T.fill_layers([[0, 1], [1, 2]], 5) == {"layers": [[[0, 1], [1, 2]], 2], "free": [[0, 1]]}
    assert T.num_qubits_from_layer([[0, 3], [1, 2]]) == 4


def test_penalty_table_matches_oracle():
    import torch
    pf = PN.make_regularization_function(PN.RegularizationOptions)
    R = O.make_regularization_function()
    x = np.concatenate([np.linspace(-7, 14, 4001), [0, np.pi, 2 * np.pi, np.pi / 2 - 0.05, 0.05]])
    assert np.abs(pf(x) - R(torch.tensor(x)).numpy()).max() < 1e-12
    assert len(pf.segments) <= L.MAX_SEGMENTS
    l1 = PN.make_regularization_function(PN.RegularizationOptions(function="L1"))
    assert np.array_equal(l1(x), np.abs(x))


def test_toffoli_targets():
    assert np.array_equal(G.u_toff3, O.toffoli_target(3).numpy())
    assert np.array_equal(G.u_toff4, O.toffoli_target(4).numpy())
    assert np.array_equal(G.u_toff5, O.toffoli_target(5).numpy())


def test_product_path_has_no_cpu_fallback():
    """Ops must fail loudly on CPU tensors instead of silently computing elsewhere."""
    import torch
    from cpflow_b200.engine import Program
    anz = A.Ansatz(3, "cp", T.fill_layers(T.chain_layer(3), 4))
    with pytest.raises(L.CpflowError):
        anz.program.unitary(torch.zeros(2, anz.num_angles))
    # and the package never imports the oracle
    import cpflow_b200, pkgutil
    for m in pkgutil.iter_modules(cpflow_b200.__path__):
        src = open(os.path.join(cpflow_b200.__path__[0], m.name + ".py")).read() if not m.ispkg else ""
        assert "oracle" not in src.replace("oracle restates", ""), m.name


def test_launch_plan_geometry_rules(lib):
    """cpf_launch_plan (heis_geometry, csrc/heis_impl.cuh) without a device: every plan covers the batch, fits the
    thread, register and shared-memory limits of an SM, and the documented choices hold (two co-resident CTAs only
    when each keeps 8 warps and the SM as many samples as with one)."""
    import torch
    from cpflow_b200.engine import Program  # noqa: F401
    cases = [(2, [[0, 1]], 3), (3, T.chain_layer(3), 12), (3, T.connected_layer(3), 7), (4, T.chain_layer(4), 40),
             (4, [[0, 1], [0, 2], [0, 3]], 40), (4, T.connected_layer(4), 61), (5, T.chain_layer(5), 60),
             (5, T.connected_layer(5), 60)]
    for n, layer, K in cases:
        anz = A.Ansatz(n, "cp", T.fill_layers(layer, K))
        n_su2, n_cp, nbl = n + 2 * K, K, len(layer)
        for dt, rs in ((torch.float32, 4), (torch.float64, 8)):
            for B in (0, 1, 7, 147, 148, 149, 1000, 12500, 100000, 1000003):
                # (5-qubit complex128: one warp per sample at up to 255 registers, blocks of at most 256 threads)
                for regs in ((200, 255) if dt == torch.float64 and n == 5 else (96, 118, 128)):
                    p = anz.program.launch_plan(B, dtype=dt, n_sm=148, regs_per_thread=regs)
                    assert p["engine"] == 1
                    tps, spc, blk, ctas = p["threads_per_sample"], p["samples_per_cta"], p["block_threads"], p["ctas_per_sm"]
                    words = 8 * n_su2 + 3 * n_cp + 8 * max(2 * nbl, n)
                    assert p["words_per_sample"] >= words and p["words_per_sample"] - words < 8
                    assert (p["words_per_sample"] // 4) % 2 == 1                      # odd number of 16-byte groups
                    assert blk % 32 == 0 and 32 <= blk <= p["max_block_threads"] and spc * tps <= blk < spc * tps + 32
                    assert p["grid"] * spc >= B and (p["grid"] - 1) * spc < max(B, 1)
                    assert ctas in (1, 2) and ctas * (p["smem_bytes"] + 1024) <= 227 * 1024
                    regs_warp = (regs * 32 + 255) // 256 * 256
                    assert ctas * (blk // 32) * regs_warp <= 65536
                    if ctas == 2:
                        assert blk >= 256
                    assert p["smem_bytes"] >= spc * p["words_per_sample"] * rs
    c3 = A.Ansatz(4, "cp", T.fill_layers(T.chain_layer(4), 40)).program
    p = c3.launch_plan(100000, n_sm=148, regs_per_thread=118)
    assert (p["ctas_per_sm"], p["block_threads"], p["samples_per_cta"]) == (2, 256, 31)      # the bench launch
    p = c3.launch_plan(12500, n_sm=148, regs_per_thread=118)
    assert (p["ctas_per_sm"], p["block_threads"], p["samples_per_cta"]) == (1, 352, 43)
    assert c3.launch_plan(100, loss_kind=L.LOSS_STATE)["engine"] == 0


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the CUDA arm): one JSON line on stdout with
    the contract's keys; ranks other than 0 print nothing."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--cpu-samples", "8", "--cpu-iters", "2"]
    env = dict(os.environ, RANK="0")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["metric"].startswith("loss+grad evals/sec") and d["config"]["workload"].startswith("C3")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1"))
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_built_kernels_fit_the_planned_occupancy(lib):
    """Resource usage of the built sm_100a kernels (cuobjdump -res-usage on the in-tree library): the Heisenberg
    kernels of the complex64 templates up to 4 qubits must stay within 128 registers (two co-resident CTAs of 8
    warps, heis_geometry) without a spill frame beyond the sincos fallback's 32 bytes; every engine kernel targets
    sm_100a."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-res-usage", L.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
    assert "sm_100a" in out
    usage = {}
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s*REG:(\d+) STACK:(\d+)", line)
        if m and name:
            usage[name] = (int(m.group(1)), int(m.group(2)))
    heis32 = {k: v for k, v in usage.items() if "heis_kernelIfLi" in k}
    assert len(heis32) >= 8
    for k, (regs, stack) in heis32.items():
        assert regs <= 128, (k, regs)
        if "heis_kernelIfLi5" not in k:                       # 5 qubits: one warp per sample, small spill accepted
            assert stack <= 32, (k, stack)
    c3 = [v for k, v in heis32.items() if "HeisSweepIfLi4ELi2ELi3ELy528ELy801" in k]
    assert c3 and c3[0][0] <= 128        # the bench kernel: 2 CTAs x 256 threads x 128 regs = the 64 K register file


def test_tabulate_penalty_callable():
    """A callable cp_regularization_func (reference main.py:536-539) becomes the kernels' segment table, or is
    rejected with the fit error."""
    import math
    import torch
    from cpflow_b200.penalty import RegularizationOptions, make_regularization_function, tabulate_penalty
    pf0 = make_regularization_function(RegularizationOptions)
    pf = tabulate_penalty(lambda a: pf0(a))
    x = np.random.default_rng(0).uniform(-20, 20, 50000)
    assert len(pf.segments) == 9 and np.abs(pf(x) - pf0(x)).max() < 1e-12
    assert np.abs(pf(x) - O.make_regularization_function()(torch.tensor(x)).numpy()).max() < 1e-12
    tri = lambda a: np.abs((np.mod(a, 2 * math.pi) / math.pi) - 1)          # noqa: E731
    assert len(tabulate_penalty(tri).segments) == 2
    ramp = lambda a: float(min(math.fmod(a, 2 * math.pi), 1.0))             # noqa: E731  scalar-only callable
    pr = tabulate_penalty(ramp, grid=1 << 12)
    assert len(pr.segments) == 2 and abs(pr(0.5) - 0.5) < 1e-9 and abs(pr(4.0) - 1.0) < 1e-9
    with pytest.raises(ValueError):
        tabulate_penalty(lambda a: (np.mod(a, 2 * math.pi) > 2.0) * 1.0)    # a jump inside the period
    for bad in (lambda a: np.sin(a) ** 2, lambda a: np.abs(a)):
        with pytest.raises(ValueError, match="periodic|piecewise|pieces"):
            tabulate_penalty(bad)


def test_oracle_normal_sampler():
    """cp_dist='normal' in the oracle (cp_utils.py:38-40): XLA's float32 erf_inv polynomial against scipy, moments of
    the draws, and the non-CP angles untouched."""
    from scipy.special import erfinv
    x = np.linspace(-0.999999, 0.999999, 20001).astype(np.float32)
    big = np.abs(x) > 1e-3
    assert np.abs(O.erf_inv_f32(x)[big] / erfinv(x.astype(np.float64))[big] - 1).max() < 1e-6
    mask = (np.arange(93) % 7 == 2).astype(int)
    a = O.generate_initial_angles(0, 93, mask, cp_dist="normal", batch_size=1500)
    u = O.generate_initial_angles(0, 93, mask, cp_dist="uniform", batch_size=1500)
    cp_draws = a[:, mask == 1]
    assert abs(cp_draws.mean()) < 0.05 and abs(cp_draws.std() - 1.5) < 0.05
    assert np.array_equal(a[:, mask == 0], u[:, mask == 0])


def test_workspace_query_without_a_device(lib):
    """cpf_workspace_bytes needs no device: 16 bytes per parameter and sample of packed optimiser state on the
    Heisenberg kernel (32 in complex128), a staged target only on the state-adjoint kernels."""
    import torch
    anz = A.Ansatz(4, "cp", T.fill_layers(T.chain_layer(4), 40))
    P = anz.num_angles
    w32 = anz.program.workspace_bytes(1000)
    w64 = anz.program.workspace_bytes(1000, dtype=torch.float64)
    # (the packed state is lane-interleaved: positions are padded to the lanes' pair slots, heis_impl.cuh: heis_pk_stride)
    assert 16 * 1000 * P <= w32 <= 16 * 1000 * (P + 48) + 16 * 1000 * 88 + 8192
    assert 32 * 1000 * P <= w64 <= 2.6 * w32 + 8192      # complex128: 16 lanes per sample, more padding
    assert anz.program.workspace_bytes(1000, L.LOSS_STATE) < 4096
    assert anz.program.workspace_bytes(0) >= 0
