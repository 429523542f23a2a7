import csv, sys, subprocess, json, collections, re
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h,u,v = rows[0],rows[1],rows[2]
want = ['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
 'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed.sum.per_cycle_active','smsp__inst_executed.sum',
 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
 'dram__bytes_read.sum','dram__bytes_write.sum','dram__bytes_read.sum.per_second','dram__bytes_write.sum.per_second','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__cycles_elapsed.max','smsp__thread_inst_executed_per_inst_executed.ratio']
d = {}
for i,n in enumerate(h):
    if n in want: d[n] = (v[i], u[i])
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(src.splitlines()))
hh = rows[1]; idx = {n:i for i,n in enumerate(hh)}
byop = collections.Counter(); stall = collections.Counter(); tot=0; samples=0
scols=[n for n in hh if n.startswith('stall_') and 'Not Issued' not in n]
for r in rows[2:]:
    if len(r) < len(hh): continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[idx['Source']]); op = m.group(2) if m else '?'
    n = int(r[idx['Instructions Executed']] or 0); byop[op]+=n; tot+=n
    samples += int(r[idx['# Samples']] or 0)
    for c in scols:
        if r[idx[c]]: stall[c]+=int(r[idx[c]])
with open(out,'w') as f:
    f.write('# ncu summary of %s (kernel %s)\n\n' % (rep.split('/')[-1], rows[0][1][:80]))
    f.write('| metric | value | unit |\n|---|---|---|\n')
    for k in want:
        if k in d: f.write('| %s | %s | %s |\n' % (k, d[k][0], d[k][1]))
    f.write('\nwarp instructions executed: %d\n\n| opcode | warp instr | share |\n|---|---|---|\n' % tot)
    for op,n in byop.most_common(14): f.write('| %s | %d | %.1f%% |\n' % (op,n,100*n/tot))
    f.write('\nwarp stall samples (share of %d samples):\n\n' % samples)
    f.write(', '.join('%s %.1f%%' % (k.replace('stall_',''), 100*x/samples) for k,x in stall.most_common(9)) + '\n')
print(open(out).read())
