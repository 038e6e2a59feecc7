mkdir -p gpurun_out
CMD="python bench.py --clips 512 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
for sk in 0 4 1 2 7; do
B2A_SKIP=$sk ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:fast_logmel -s 2 -c 1 --csv --log-file gpurun_out/ncu_skip_$sk.csv $CMD > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncu_skip_$sk.csv")) if len(r)>10]
h=rows[0]; 
out={}
for r in rows[1:]:
    out[r[h.index("Metric Name")]]=r[h.index("Metric Value")]
print("skip=$sk", out)
PY
done
