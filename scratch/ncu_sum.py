import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
h=rows[0]; v=rows[2]
keep=('smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','gpu__time_duration.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warps_active.avg.per_cycle_active','sm__inst_executed_pipe_lsu.sum','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed','sm__icc_requests_lookup_miss.sum','dram__bytes_read.sum','dram__bytes_write.sum')
for k,x in zip(h,v):
    if ('issue_stalled' in k and 'per_issue_active' in k) or k in keep or 'icc' in k or 'idc' in k:
        print(k.replace('smsp__average_warps_issue_stalled_','stall ').replace('_per_issue_active.ratio',''),x)
