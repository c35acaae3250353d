"""profiles/r2_roofline.json from the round's ncu captures (read by bench.py for the roofline block):

    python tools/ncu_roofline.py REP EVALS TRAFFIC_CSV SAMPLES ITERS OUT.json

REP: `ncu --set full` capture of heis_kernel on EVALS = samples x steps evaluations (tools/gpu_ncu_r2.sh);
TRAFFIC_CSV: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` of one bench-shaped run
(SAMPLES x ITERS; all heis_kernel launches of the run are summed).  Executed flops per evaluation are counted from the
per-instruction thread-level execution counts of the capture (FFMA2 = 4, FFMA = FMUL2 = FADD2 = 2, FMUL = FADD = 1)."""
import collections
import csv
import io
import json
import subprocess
import sys

rep, evals, traffic_csv, samples, iters, out = sys.argv[1], float(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))


def num(k):
    return float(d[k].replace(",", "")) if k in d and d[k] not in ("", "n/a") else None


src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], [r for r in rows[2:] if len(r) > 6]
iS, iT, iE = hdr.index("Source"), hdr.index("Thread Instructions Executed"), hdr.index("Instructions Executed")
FLOP = {"FFMA2": 4, "FFMA": 2, "FMUL2": 2, "FADD2": 2, "FMUL": 1, "FADD": 1, "DFMA": 2, "DMUL": 1, "DADD": 1}
flops = 0.0
mix = collections.Counter()
for r in data:
    s = r[iS].strip()
    if s.startswith("@"):
        s = s.split(None, 1)[1]
    op = s.split()[0].split(".")[0]
    mix[op] += int(r[iE])
    flops += FLOP.get(op, 0) * int(r[iT])
tot = sum(mix.values())
fp = sum(mix[o] for o in FLOP)
launches = []
t_rows = [r for r in csv.reader(open(traffic_csv)) if len(r) > 5]
if t_rows:
    h = t_rows[0]
    iK, iM, iV = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    agg = collections.Counter()
    n_launch = set()
    for r in t_rows[1:]:
        if "heis_kernel" in r[iK] and "pack_target" not in r[iK]:
            agg[r[iM]] += float(r[iV].replace(",", ""))
            n_launch.add(r[h.index("ID")])
    launches.append({"kernel": "cpf::heis_kernel<float,4,2,HeisSweep<chain>>", "samples": samples, "iters": iters,
                     "kernel_launches": len(n_launch), "dram_bytes_read": agg["dram__bytes_read.sum"],
                     "dram_bytes_write": agg["dram__bytes_write.sum"],
                     "dram_bytes_per_run": agg["dram__bytes_read.sum"] + agg["dram__bytes_write.sum"],
                     "dram_bytes_per_eval": (agg["dram__bytes_read.sum"] + agg["dram__bytes_write.sum"]) / (samples * iters),
                     "gpu_time_ns_under_ncu": agg["gpu__time_duration.sum"]})
res = {"capture": rep, "evals_in_capture": evals, "kernel": d.get("Kernel Name"),
       "grid": d.get("Grid Size"), "block": d.get("Block Size"), "registers_per_thread": num("launch__registers_per_thread"),
       "pipe_fma_cycles_active_pct": num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
       "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
       "warp_instructions_per_eval": num("smsp__inst_executed.sum") / evals,
       "executed_flops_per_eval_from_opcode_mix": flops / evals,
       "fp_instruction_share": fp / tot,
       "unfused_multiply_share_of_fp_instructions": (mix["FMUL2"] + mix["FMUL"]) / fp,
       "opcode_mix_warp_instr_per_eval": {o: c / evals for o, c in mix.most_common(16)},
       "launches": launches}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k not in ("opcode_mix_warp_instr_per_eval",)}, indent=1))
