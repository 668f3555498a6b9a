import csv, sys, subprocess, collections, io
rep = sys.argv[1]; units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, un, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__thread_inst_executed_per_inst_executed.ratio','sass__inst_executed_local_loads','sass__inst_executed_local_stores','lts__t_bytes.sum','l1tex__t_bytes.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__sass_branch_targets_threads_divergent','sm__sass_branch_targets']
d = dict(zip(hdr, vals))
for h,u,v in zip(hdr,un,vals):
    if h in want or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and float(v.replace(',','') or 0) > 0.3): print(f'{h:85s} {u:12s} {v}')
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; ci = {h:i for i,h in enumerate(hdr)}; data = rows[2:]
ops = collections.Counter(); samp = collections.Counter()
for r in data:
    toks = r[ci['Source']].split(); op = toks[1] if toks[0].startswith('@') else toks[0]; op = op.split('.')[0]
    ops[op] += int(r[ci['Instructions Executed']]); samp[op] += int(r[ci['# Samples']])
tot = sum(ops.values()); print('total warp inst', tot, ('per unit %.1f' % (tot/units)) if units else '')
for op,c in ops.most_common(16): print(f'  {op:10s} {c/(units or 1):10.2f}  samples {samp[op]}')
top = sorted(data, key=lambda r:-int(r[ci['# Samples']]))[:10]
for r in top: print(r[ci['# Samples']], r[ci['Source']][:60], '| long_sb', r[ci['stall_long_sb']], 'wait', r[ci['stall_wait']], 'short', r[ci['stall_short_sb']], 'branch', r[ci['stall_branch_resolving']], 'barrier', r[ci['stall_barrier']])
