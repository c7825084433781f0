import csv,sys,subprocess
rep=sys.argv[1]; samples=float(sys.argv[2]) if len(sys.argv)>2 else 1e6; T=564
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
h=rows[0]; u=rows[1]; v=rows[2]
d={n:(v[i],u[i]) for i,n in enumerate(h)}
keys=['gpu__time_duration.sum','sm__inst_executed.avg.per_cycle_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__warps_eligible.avg.per_cycle_active','smsp__inst_executed.sum','sm__cycles_elapsed.avg.per_second','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sectors_op_red.sum','lts__t_sectors_op_atom.sum']
for k in keys:
    if k in d: print(f"{k:75s} {d[k][0]:>18s} {d[k][1]}")
inst=float(d['smsp__inst_executed.sum'][0]); print("warp-instr per warp-day", inst/(samples*T/32))
for n in h:
    if 'issue_stalled' in n and n.endswith('.ratio') and 'not_issued' not in n and float(d[n][0])>0.05:
        print(f"{n:90s} {d[n][0]}")
