"""Measured parity maxima of the CUDA engine against the CPU oracle (run on the GPU box):

    python tools/parity_report.py > profiles/parity_r2.txt

Same measurement functions as tests/test_gpu_parity.py (tests/parity_lib.py); the tests assert the contract's
tolerances (1e-5 complex64, 1e-12 complex128), this prints what was actually measured per configuration."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import parity_lib as P  # noqa: E402
from cpflow_b200.gates import u_toff4  # noqa: E402
from cpflow_b200.topology import chain_layer  # noqa: E402


def main():
    print("# loss / reg: max |x - oracle| / max |oracle| over 37 samples; grad: max over samples of |g - g_o| / |g_o|")
    print(f"# device: {torch.cuda.get_device_name(0)}")
    print(f"{'config':44s} {'dtype':5s} {'kind':8s} {'engine':6s} {'loss':>9s} {'reg':>9s} {'grad':>9s} {'unitary':>9s}")
    worst = {torch.float32: 0.0, torch.float64: 0.0}
    for n, layer, K, rg in P.CONFIGS:
        for dt in (torch.float32, torch.float64):
            m = P.measure_loss_grad(n, layer, K, rg, dt)
            anz, _, _ = P.setup(n, layer, K, rg)
            for kind in ("hs", "relphase", "state"):
                e = m[kind]
                eng = anz.program.launch_plan(37, loss_kind={"hs": 0, "state": 1, "relphase": 2}[kind], dtype=dt)["engine"]
                name = f"n={n} K={K} {rg} {layer}"
                print(f"{name[:44]:44s} {'c64' if dt == torch.float32 else 'c128':5s} {kind:8s} "
                      f"{'heis' if eng == 1 else 'adj':6s} {e['loss']:9.2e} {e['reg']:9.2e} {e['grad']:9.2e} {m['unitary']:9.2e}")
                worst[dt] = max(worst[dt], e["loss"], e["reg"], e["grad"])
    print(f"# worst complex64: {worst[torch.float32]:.3e} (contract 1e-5)   worst complex128: {worst[torch.float64]:.3e} (contract 1e-12)")
    print("#\n# fused Adam loop vs the oracle loop, C3 shape (n=4, K=40, 32 samples, complex128, T=150): max abs errors")
    for lname, layer in (("chain", chain_layer(4)), ("star", P.STAR4)):
        for freeze in (False, True):
            m = P.measure_adam_loop(4, layer, 40, u_toff4, B=32, T=150, freeze=freeze)
            print(f"{lname:6s} freeze={int(freeze)} " + " ".join(f"{k}={v:.2e}" if isinstance(v, float) else f"{k}={v}"
                                                                 for k, v in m.items()))
    m = P.measure_adam_loop(4, chain_layer(4), 40, u_toff4, B=32, T=10, dt=torch.float32)
    print("chain  complex64 T=10 " + " ".join(f"{k}={v:.2e}" if isinstance(v, float) else f"{k}={v}" for k, v in m.items()))


if __name__ == "__main__":
    main()
