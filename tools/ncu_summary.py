"""Summarise an .ncu-rep of the engine kernel: headline metrics, stall breakdown, dynamic opcode mix.
usage: python tools/ncu_summary.py gpurun_out/prof_engine.ncu-rep EVALS [out.txt]"""
import collections, csv, io, re, subprocess, sys
rep, evals = sys.argv[1], float(sys.argv[2])
out = open(sys.argv[3], 'w') if len(sys.argv) > 3 else sys.stdout
def P(*a): print(*a, file=out)
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
P('kernel:', d.get('Kernel Name'), 'grid', d.get('Grid Size'), 'block', d.get('Block Size'))
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__shared_mem_per_block_dynamic']
for k in keys:
    if k in d: P(f'{k:85s} {d[k]}')
P('warp instructions per eval:', float(d['smsp__inst_executed.sum'].replace(',', '')) / evals)
P('-- stalls (warps per issue) --')
st = {k: float(v) for k, v in d.items() if re.match(r'smsp__average_warps_issue_stalled_.*_per_issue_active.ratio', k)}
for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:12]:
    P(f"  {k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):25s} {v:.3f}")
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
op = collections.Counter(); smp = collections.Counter(); tot = 0
for r in data:
    if len(r) <= iE: continue
    s = r[iS].strip()
    if s.startswith('@'): s = s.split(None, 1)[1]
    o = s.split()[0].split('.')[0]
    e = int(r[iE]); op[o] += e; smp[o] += int(r[iSm]); tot += e
P('-- dynamic opcode mix (warp instr per eval, % of instr, % of stall samples) --')
ts = sum(smp.values())
for o, c in op.most_common(22):
    P(f'  {o:8s} {c/evals:9.1f} {100*c/tot:5.1f}% {100*smp[o]/ts:5.1f}%')
