ncu --metrics sm__icc_request_hit_rate.pct,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_shared_mem --clock-control none -k regex:fast_logmel -c 6 --csv --log-file gpurun_out/cfg_metrics.csv python benchmarks/bench_configs.py --only C3,C5 --steps 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/cfg_metrics.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value'); ii=h.index('ID')
cur=None
for r in rows[1:]:
    if r[ii]!=cur:
        cur=r[ii]; print('---',r[ik][40:150])
    print('   ',r[im][:70],r[iv])
PY
