"""Summarise an ncu report: headline metrics, stall breakdown and the hottest source lines.
   python scripts/ncu_summary.py gpurun_out/x.ncu-rep [nlines]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
for k in keys:
    print(f'{k:90s} {d.get(k)}')
print('-- stalls per issue')
st = [(float(v), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
      for h, v in d.items() if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
for v, h in sorted(st, reverse=True)[:10]:
    print(f'   {h:28s} {v:.3f}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
while rows and (not rows[0] or rows[0][0] != 'Address'): rows.pop(0)
hdr = rows[0]
def col(name):
    for i, h in enumerate(hdr):
        if h.strip() == name: return i
    return None
def colp(prefix):
    for i, h in enumerate(hdr):
        if h.strip().startswith(prefix): return i
    return None
ci, cs, cx = col('Source'), colp('Warp Stall Sampling (All'), col('Instructions Executed')
tot = sum(float(r[cs] or 0) for r in rows[1:] if len(r) > cs)
totx = sum(float(r[cx] or 0) for r in rows[1:] if len(r) > cx)
print('-- total samples', tot, 'instr', totx)
byop = collections.Counter(); byopx = collections.Counter()
for r in rows[1:]:
    if len(r) <= cs: continue
    op = r[ci].split()[0] if r[ci].split() else ''
    if op.startswith('@'): op = r[ci].split()[1]
    byop[op.split('.')[0]] += float(r[cs] or 0); byopx[op.split('.')[0]] += float(r[cx] or 0)
print('-- by opcode (samples %, instr %)')
for op, v in byop.most_common(24):
    print(f'   {op:14s} {100*v/tot:6.2f} {100*byopx[op]/totx:6.2f}')
print('-- hottest SASS lines')
top = sorted(rows[1:], key=lambda r: -float(r[cs] or 0) if len(r) > cs else 0)[:nl]
for r in top:
    print(f'   {100*float(r[cs] or 0)/tot:5.2f}%  {r[ci][:110]}')
