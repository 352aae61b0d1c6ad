import csv, sys, subprocess, collections
rep=sys.argv[1]; items=float(sys.argv[2]) if len(sys.argv)>2 else 430000
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
d={h:(v,u) for h,u,v in zip(hdr,units,vals)}
for w in want:
    if w in d: print(f"{w} = {d[w][0]} {d[w][1]}")
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines())); hdr=rows[1]; data=[r for r in rows[2:] if len(r)>=len(hdr)]
ix={h:i for i,h in enumerate(hdr)}
byop=collections.Counter(); samp=collections.Counter(); wf=collections.Counter(); wfex=collections.Counter(); tot=0
for r in data:
    s=r[ix['Source']].strip().split(); op=(s[1] if s[0].startswith('@') else s[0]).split('.')[0]
    n=int(r[ix['Instructions Executed']]); tot+=n; byop[op]+=n; samp[op]+=int(r[ix['# Samples']]); wf[op]+=int(r[ix['L1 Wavefronts Shared']]); wfex[op]+=int(r[ix['L1 Wavefronts Shared Excessive']])
print("warp-inst/item", tot/items, "samples", sum(samp.values()))
for op,n in byop.most_common(22): print(f"  {op:10s} {n/items:7.1f}/item samples {samp[op]:6d} smem wf {wf[op]/items:6.1f} excess {wfex[op]/items:6.1f}")
open('/tmp/last_src.csv','w').write(src)
