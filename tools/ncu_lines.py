"""Executed warp instructions of an ncu capture of heis_kernel by opcode and INNERMOST source line (nvdisasm -gi line
table of the in-tree cubin), for one phase of the kernel body (the outermost frame's line range).
usage: python tools/ncu_lines.py REP EVALS PHASE [fp|nonfp|all] [top]      PHASE: update | forward | backward"""
import collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, evals, phase = sys.argv[1], float(sys.argv[2]), sys.argv[3]
kind = sys.argv[4] if len(sys.argv) > 4 else "nonfp"
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
want = os.environ.get("NCU_LINES_KERNEL", "heis_kernelIfLi4ELi2ENS_9HeisSweepIfLi4ELi2ELi3ELy528ELy801")
src = open(os.path.join(ROOT, "cpflow_b200", "csrc", "heis_impl.cuh")).read().split("\n")
def find(txt):
    return next(i + 1 for i, l in enumerate(src) if txt in l)
lo = {"update": find("for (int it = 0; it <= p.nsteps"), "forward": find("V yr[N], yi[N];"), "backward": find("SWP::backward(p, lb")}[phase]
hi = {"update": find("V yr[N], yi[N];"), "forward": find("SWP::gather_wht(yr, yi)"), "backward": find("if (p.mode == M_ADAM && active) {\n") if False else lo + 1}[phase]
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
ins, grp, cur, new, inside = [], [], [], False, False
for ln in sass:
    if ln.startswith(".text."):
        inside = want in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        if not new:
            grp, new = [], True
        grp.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m:
        if new:
            cur, new = list(grp), False
        ins.append((m.group(2).strip(), cur))
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) > 5]
iE, iN = hdr.index("Instructions Executed"), hdr.index("# Samples")
assert len(data) == len(ins), (len(data), len(ins))
FP = {"FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD"}
agg = collections.Counter(); smp = collections.Counter(); tot = 0; tots = 0
for r, (txt, chain) in zip(data, ins):
    if not chain or chain[-1][0] != "heis_impl.cuh" or not (lo <= chain[-1][1] < hi):
        continue
    op = (txt.split()[1] if txt.startswith("@") else txt.split()[0]).split(".")[0]
    if kind == "fp" and op not in FP or kind == "nonfp" and op in FP:
        continue
    inner = next((c for c in chain if c[0] == "heis_impl.cuh"), chain[0])
    agg[(op, inner[1])] += int(r[iE]); smp[(op, inner[1])] += int(r[iN]); tot += int(r[iE]); tots += int(r[iN])
print(f"{phase}: {tot / evals:.1f} {kind} warp instr per eval, {tots} samples")
for (op, line), n in agg.most_common(top):
    print(f"{n / evals:8.1f}  {smp[(op, line)]:6d}  {op:8s} {line:5d}  {src[line - 1].strip()[:100]}")
