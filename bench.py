#!/usr/bin/env python
"""Headline benchmark: loss+grad(+Adam step) evaluations per second of the multi-start synthesis
loop on BASELINE.json's metric configuration (4-qubit Toffoli, 40 CP gates, complex64).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --steps K --warmup W    # CPU arm (oracle port of cpflow)

A *step* is one complete stage-1 run of `Synthesize.static()` for this rank's shard of samples:
B = 10^5 independent random initialisations (the sample count of BASELINE configs[2]) x T = 2000
Adam iterations, every iteration = forward sweep + loss + penalty + adjoint sweep + Adam update +
best tracking, all inside ONE launch of the fused engine kernel.  One "eval" = one such iteration
of one sample.  Samples shard over ranks with no data-path collective ("weak" scaling: per-GPU
batch fixed); timing is CUDA events on the launching stream, max over ranks.

Keys beyond the base contract:
  roofline     FP32 CUDA-core roofline of the engine kernel (the path is gate
               arithmetic, not HBM or tensor-core bound).  achieved = algorithmic flops per launch
               (C*N*(16*G1+4*K+8) per eval, the adjoint-sweep count; uncompute not credited) / mean
               launch time measured live with CUDA events.  The Heisenberg-picture kernel
               (csrc/heis_impl.cuh) needs fewer executed flops than that count; peak = FP32 FMA issue peak measured live by tools/fp32_peak
               on the same GPU (MEASURED_PEAKS.json carries no FP32 figure).  The HBM side is
               reported next to it (hbm_*), it is not the bound.
  cpu_baseline the CPU oracle (torch restatement of the reference's JAX loop; JAX is not in the
               image) timed on this host, bounded sample.
  e2e          same metric through the public host API (`mynimize_repeated` with HOST buffers):
               pinned H2D of the initial angles and D2H of the result arrays inside the timing.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "loss+grad evals/sec (4q Toffoli, 40 CP gates)"
UNIT = "evals/s"
R_WEIGHT = 0.001476          # the reference's stored K=40 trial point
LR = 0.1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # C3 quotes 10^5 samples: at N = 1 the whole configuration runs on one GPU; N > 1 keeps the per-GPU work (weak)
    ap.add_argument("--samples-per-gpu", type=int, default=100000)
    ap.add_argument("--iters", type=int, default=2000, help="Adam iterations per step (num_gd_iterations)")
    ap.add_argument("--layer", default="chain", choices=["chain", "star"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-static", action="store_true", help="skip the Synthesize.static() wall-time extras")
    ap.add_argument("--cpu-samples", type=int, default=1024)
    ap.add_argument("--cpu-iters", type=int, default=30)
    return ap.parse_args()


def workload(layer_name):
    from cpflow_b200.topology import chain_layer
    layer = chain_layer(4) if layer_name == "chain" else [[0, 1], [0, 2], [0, 3]]
    return layer, 40


def config_dict(args, n_gpus, extra=None):
    c = {"workload": f"C3: 4q Toffoli (C3X), {args.layer} connectivity, 40 CP gates, 'xyz' rotations, P=292, "
                     f"HS loss + linear CP penalty r={R_WEIGHT}, Adam lr={LR}, complex64",
         "samples_per_gpu": args.samples_per_gpu, "global_samples": args.samples_per_gpu * n_gpus,
         "adam_iterations_per_step": args.iters, "parallelism": f"samples sharded over {n_gpus} GPU(s), no collective"}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference loop (torch CPU, complex64, all host threads)
# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(layer, K, samples, iters, repeats=1):
    """evals/s of the oracle's Adam loop (value_and_grad + optax-Adam restatement) on this host."""
    import torch
    from oracle import cpflow_oracle as O
    anz = O.cp_ansatz(layer, K)
    ops = O.ansatz_program(anz)
    tgt = O.toffoli_target(4, torch.complex64)
    R = O.make_regularization_function()
    a0 = torch.tensor(O.generate_initial_angles(0, anz.num_angles, anz.cp_mask, batch_size=samples))
    best = 0.0
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.mynimize_repeated(4, ops, "hs", tgt, a0, LR, iters, anz.cp_mask, R_WEIGHT, R)
        dt = time.perf_counter() - t0
        best = max(best, samples * (iters + 1) / dt)   # the oracle evaluates theta_0 twice like the reference
    return best, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    layer, K = workload(args.layer)
    for _ in range(args.warmup):
        cpu_oracle_rate(layer, K, min(args.cpu_samples, 256), 2)
    t0 = time.perf_counter()
    rates = []
    for _ in range(args.steps):
        r, threads = cpu_oracle_rate(layer, K, args.cpu_samples, args.cpu_iters)
        rates.append(r)
    dt = time.perf_counter() - t0
    evals = args.steps * args.cpu_samples * (args.cpu_iters + 1)
    value = evals / dt
    sample = (f"{args.cpu_samples} samples x {args.cpu_iters} Adam iterations per step of the same C3 program "
              f"(bounded sample of the {args.samples_per_gpu}x{args.iters} step)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (complex64)",
            "data": "synthetic", "config": config_dict(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "note": "CPU oracle = torch restatement of cpflow's jit(vmap(value_and_grad)) + optax "
                                     "Adam loop; JAX/optax are not installable in this image"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cores": os.cpu_count()}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# helpers for the CUDA arm
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.stop_flag = threading.Event()
        self.sm, self.reasons, self.power = [], set(), []
        self.sm_max = None
        self.err = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                     nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
                     nv.nvmlClocksThrottleReasonSyncBoost: "sync_boost",
                     nv.nvmlClocksThrottleReasonApplicationsClocksSetting: "applications_clocks_setting"}
            while not self.stop_flag.is_set():
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:  # NVML missing: report, do not fail the bench
            self.err = repr(e)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=2)
        d = {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max,
             "reasons": sorted(self.reasons), "samples": len(self.sm)}
        if self.power:
            d["power_w_median"] = statistics.median(self.power)
        if self.err:
            d["error"] = self.err
        return d


def measure_fp32_peak():
    """FP32 FMA issue peak of this GPU, measured now by tools/fp32_peak (built by build())."""
    exe = os.path.join(ROOT, "tools", "fp32_peak")
    src = exe + ".cu"
    try:
        if not os.path.exists(exe):
            subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", exe, src],
                           check=True, capture_output=True)
        out = subprocess.run([exe, "--quick"], capture_output=True, text=True, timeout=120).stdout
        best = {}
        for ln in out.splitlines():
            if ln.startswith("{"):
                d = json.loads(ln)
                best[d["bench"]] = max(best.get(d["bench"], 0.0), d["rate_per_s"])
        peak = max(best.get("ffma_flops", 0.0), best.get("ffma2_flops", 0.0))
        if peak > 0:
            return peak / 1e12, "measured live by tools/fp32_peak (max of FFMA / FFMA2 chains)", best
    except Exception as e:
        return 74.4, f"nominal 148 SM x 128 lanes x 2 x 1.965 GHz (fp32_peak failed: {e!r})", {}
    return 74.4, "nominal 148 SM x 128 lanes x 2 x 1.965 GHz (fp32_peak gave no result)", {}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def static_wall_times():
    """Second half of BASELINE.json's metric: wall time of the whole Synthesize.static() call (sampling, the
    fused Adam stage, selection, batched verification, result objects) on BASELINE configs[0] (README example,
    'one to five minutes' on the reference's CPU path) and configs[1] (Toffoli-3, all-to-all, 10^4 samples)."""
    import contextlib
    import io
    import numpy as np
    import cpflow_b200 as cp
    from cpflow_b200.gates import u_toff3
    from cpflow_b200.topology import chain_layer, connected_layer
    out = {}
    ccz = np.diag([1, 1, 1, 1, 1, 1, 1, -1]).astype(complex)
    cases = [("C1_ccz_chain_K12_B10", chain_layer(3), ccz,
              dict(num_cp_gates=12, accepted_num_cz_gates=10, num_samples=10)),
             ("C2_toffoli3_connected_K7_B10000", connected_layer(3), u_toff3,
              dict(num_cp_gates=7, r=0.00131, accepted_num_cz_gates=6, num_samples=10000))]
    for name, layer, target, kw in cases:
        try:
            syn = cp.Synthesize(layer, target_unitary=target, label=name)
            opts = cp.StaticOptions(**kw)
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                syn.static(opts, save_results=False)          # warm-up (module import, allocator, device program)
                t0 = time.perf_counter()
                res = syn.static(opts, save_results=False)
                dt = time.perf_counter() - t0
            cz = sorted(d.cz_count for d in res.decompositions)
            out[name] = {"wall_s": dt, "decompositions": len(cz), "min_cz": cz[0] if cz else None,
                         "prospective": len(syn.last_prospective_cz_counts)}
        except Exception as e:   # an extra must never take the headline down
            out[name] = {"error": repr(e)}
    return out


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CUDA engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL announces its version there when the communicator is created
        # ("NCCL version ..." with NCCL_DEBUG=VERSION in the image), so creation runs with fd 1 pointed at stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    from cpflow_b200.ansatz import Ansatz
    from cpflow_b200.engine import Loss, Penalty
    from cpflow_b200.gates import u_toff4
    from cpflow_b200.optimization import ProgramLoss, mynimize_repeated
    from cpflow_b200.penalty import RegularizationOptions, make_regularization_function
    from cpflow_b200.topology import fill_layers

    layer, K = workload(args.layer)
    anz = Ansatz(4, "cp", fill_layers(layer, K))
    prog = anz.program
    pf = make_regularization_function(RegularizationOptions)
    pen = Penalty("piecewise", R_WEIGHT, pf.segments, pf.period)
    loss = Loss("hs", u_toff4)
    B, T = args.samples_per_gpu, args.iters
    flops_eval, bytes_eval = prog.eval_cost()

    # FP32 peak of this GPU (before the timed region; rank 0's GPU stands for the box)
    peak_tf, peak_src, peak_detail = (measure_fp32_peak() if rank == 0 else (None, None, None))

    # inputs resident in HBM: this rank's shard of the global batch, keyed by global sample index
    a0 = prog.initial_angles(0, B * world, first=rank * B, count=B, device=dev)
    st = prog.adam_state(a0.clone())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        st.angles.copy_(a0)
        st.step = 0
        prog.adam_run(st, loss, pen, LR, T)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush.zero_()
        step()
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        flush.zero_()
        st.angles.copy_(a0)
        st.step = 0
        ev[k][0].record()
        prog.adam_run(st, loss, pen, LR, T)
        ev[k][1].record()
    e1.record()
    barrier()
    clocks = sampler.summary()
    ms_total = e0.elapsed_time(e1)
    launch_ms = [a.elapsed_time(b) for a, b in ev]
    t = torch.tensor([ms_total, max(launch_ms), sum(launch_ms) / len(launch_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, _, launch_mean = t.tolist()
    evals_total = world * B * T * args.steps
    value = evals_total / (ms_total * 1e-3)

    # sanity of the work actually done in the timed region (not timed)
    best = st.best_regloss
    assert bool(torch.isfinite(best).all()) and bool((best <= st.init_regloss).all())
    frac_converged = float((best - st.best_reg < 1e-3).float().mean())

    # ---- e2e: the public host API with HOST buffers (pinned H2D of inputs, D2H of results) ----
    a0_host = a0.cpu().pin_memory()
    pl = ProgramLoss(prog, loss)

    def e2e_step():
        res = mynimize_repeated(pl, anz.num_angles, learning_rate=LR, num_iterations=T,
                                initial_params_batch=a0_host, regularization_func=pen, keep_history=False,
                                device=dev)
        return res

    for _ in range(2):          # untimed: first use of the pinned staging buffers and of the allocator's blocks
        res = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = evals_total / te.item()
    h2d = a0_host.numel() * a0_host.element_size()
    d2h = int(res.params.nbytes + res.regloss.nbytes + res.reg.nbytes)
    assert np.allclose(res.regloss[:, 1], best.cpu().numpy())   # same computation through both paths

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_pk, hbm_src = hbm_peak()
    achieved_tf = flops_eval * B * T / (launch_mean * 1e-3) / 1e12
    traffic = None
    rp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(rp):
        try:
            traffic = next((e["dram_bytes_per_launch"] for e in json.load(open(rp))["launches"]
                            if e["samples"] == B and e["iters"] == T), None)   # ncu capture of this launch shape
        except Exception:
            traffic = None
    roofline = {"bound": "fp32", "kernel": "cpf::heis_kernel<float,4,2,HeisSweep<chain>>",
                "note": "FP32 CUDA-core bound (north_star: no tensor cores, not HBM). 'achieved' credits the adjoint-"
                        "sweep flop count of SURVEY.md 8(d); the Heisenberg-picture kernel executes ~0.6 M flop per "
                        "eval (real Pauli-basis backward sweep), so frac can exceed the FMA-pipe utilisation ncu "
                        "reports (profiles/).",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "peak_source": peak_src, "flops_per_eval": flops_eval, "evals_per_launch": B * T,
                "launch_ms_mean": launch_mean, "traffic": traffic,
                "hbm_algorithmic_bytes_per_eval": bytes_eval,
                "hbm_achieved_gbs": bytes_eval * B * T / (launch_mean * 1e-3) / 1e9,
                "hbm_peak_gbs": hbm_pk, "hbm_peak_source": hbm_src,
                "hbm_frac": bytes_eval * B * T / (launch_mean * 1e-3) / 1e9 / hbm_pk,
                "peak_detail": peak_detail}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, threads = cpu_oracle_rate(layer, K, args.cpu_samples, args.cpu_iters)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.cpu_samples} samples x {args.cpu_iters} Adam iterations of the same C3 program "
                         f"(torch CPU oracle, complex64)", "host_cores": os.cpu_count()}

    static_wall = None
    if world == 1 and not args.no_static:
        static_wall = static_wall_times()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (complex64 amplitudes)", "data": "synthetic",
            "config": config_dict(args, world, {"l2": "flushed between steps (256 MiB write)",
                                                "fraction_of_samples_below_entry_loss": frac_converged}),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "cpflow_b200.optimization.mynimize_repeated(host arrays)", "timer": "wall clock"},
            "gpu_launches": 2 * args.steps, "clocks": clocks, "static_wall_s": static_wall}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
