"""Attribute the stall samples / executed instructions of an ncu capture of heis_kernel to the phases of the
kernel body (update, merged diagonals, forward, pivot, loss/h, SO(3), backward), using the line table of the
in-tree cubin (nvdisasm -gi: outermost inlining frame = line of the kernel body).
usage: python tools/ncu_regions.py REP EVALS [mangled-name-substring] [out.txt]"""
import bisect, collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, evals = sys.argv[1], float(sys.argv[2])
want = sys.argv[3] if len(sys.argv) > 3 else "heis_kernelIfLi4ELi2ENS_9HeisSweepIfLi4ELi2ELi3ELy528ELy801"
out = open(sys.argv[4], "w") if len(sys.argv) > 4 else sys.stdout


def P(*a): print(*a, file=out)


# ---- phase markers: first line of each phase inside heis_kernel, found by text ----
src = open(os.path.join(ROOT, "cpflow_b200", "csrc", "heis_impl.cuh")).read().split("\n")
def find(txt, start=0):
    for i in range(start, len(src)):
        if txt in src[i]:
            return i + 1
    raise KeyError(txt)
k0 = find("heis_kernel(const KParams<R> p)")
marks = [(k0, "prologue/pack"),
         (find("for (int it = 0; it <= p.nsteps", k0), "update: su2 gates"),
         (find("for (int k = m; k < p.n_cp; k += TPS) {", k0), "update: cp gates"),
         (find("// merged diagonals of the forward sweep", k0), "merged diagonals"),
         (find("V yr[N], yi[N];", k0), "forward sweep"),
         (find("SWP::gather_wht", k0), "pivot (gather+WHT)"),
         (find("const R tr = __shfl_sync", k0), "loss / best / h init"),
         (find("SWP::backward", k0), "backward sweep"),
         (find("if (p.mode == M_ADAM && active) {", find("SWP::backward", k0)), "epilogue")]
starts = [m[0] for m in marks]

# ---- line table of the kernel ----
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", os.environ.get("NCU_REGIONS_CUBIN", "inst_heis_f32"), os.path.join(ROOT, "cpflow_b200", "lib", "libcpflow_b200.so")],
                   cwd=td, check=True, capture_output=True)
    # the kernels are spread over several translation units (inst_heis_f32_p*.cu): take the cubin that has this one
    sass = []
    for cub in sorted(f for f in os.listdir(td) if f.endswith(".cubin")):
        txt = subprocess.run(["nvdisasm", "-gi", os.path.join(td, cub)], capture_output=True, text=True).stdout
        if want in txt:
            sass = txt.split("\n")
            break
ins = []   # (opcode text, kernel-body line)
inside = False
last_outer = None
for ln in sass:
    if ln.startswith(".text."):
        inside = want in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        if "inlined at" not in m.group(3) and m.group(1).endswith("heis_impl.cuh"):
            last_outer = int(m.group(2))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m:
        ins.append((m.group(2).strip(), last_outer))

csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(csvtxt)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) > 5]
iS, iE, iSm, iNi = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), \
    hdr.index("Warp Stall Sampling (Not-issued Samples)")
if len(data) != len(ins):
    P(f"WARNING: {len(data)} profiled instructions vs {len(ins)} in the local cubin (rebuilt since the capture?)")
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
ops = collections.defaultdict(collections.Counter)
mism = 0
for r, (txt, line) in zip(data, ins):
    o1 = r[iS].strip().split(None, 1)
    o1 = (o1[1] if o1[0].startswith("@") else " ".join(o1)).split()[0]
    o2 = (txt.split(None, 1)[1] if txt.startswith("@") else txt).split()[0]
    mism += o1 != o2
    reg = marks[max(0, bisect.bisect_right(starts, line or 0) - 1)][1]
    a = agg[reg]
    e = int(r[iE]); a[0] += e; a[1] += int(r[iSm]); a[2] += int(r[iNi]); a[3] += 1
    ops[reg][o2.split(".")[0]] += e
if mism:
    P(f"WARNING: {mism} opcode mismatches between the capture and the local cubin")
# per-instruction stall reasons (pc sampling), same instruction order
STALLS = ["wait", "long_scoreboard", "short_scoreboard", "barrier", "math_pipe_throttle", "not_selected", "selected",
          "branch_resolving", "no_instructions", "dispatch_stall", "mio_throttle", "lg_throttle", "membar", "sleeping", "misc"]
stxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--metrics",
                       ",".join("smsp__pcsamp_warps_issue_stalled_" + x for x in STALLS)], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(stxt)))
shdr = srows[1]; sdata = [r for r in srows[2:] if len(r) > 5]
stall = collections.defaultdict(collections.Counter)
for r, (txt, line) in zip(sdata, ins):
    reg = marks[max(0, bisect.bisect_right(starts, line or 0) - 1)][1]
    for name, v in zip(shdr[2:], r[2:]):
        stall[reg][name] += int(v)
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
P(f"{'phase':24s} {'static':>7s} {'instr/eval':>10s} {'%instr':>7s} {'%samples':>9s} {'issued/sample':>13s}   top opcodes (warp instr per eval)")
for _, name in marks:
    if name not in agg: continue
    a = agg[name]
    top = ", ".join(f"{o} {c / evals:.0f}" for o, c in ops[name].most_common(int(os.environ.get("NCU_REGIONS_TOP", "6"))))
    P(f"{name:24s} {a[3]:7d} {a[0] / evals:10.1f} {100 * a[0] / ti:6.1f}% {100 * a[1] / ts:8.1f}% {1 - a[2] / max(a[1], 1):13.2f}   {top}")
P(f"total instr/eval {ti / evals:.1f}")
P("-- stall samples per phase (% of the phase's samples; 'selected' = issuing) --")
for _, name in marks:
    if name not in stall: continue
    tot = sum(stall[name].values()) or 1
    P(f"{name:24s} " + "  ".join(f"{k.replace('stall_', '')} {100 * v / tot:.0f}" for k, v in stall[name].most_common(7)))
