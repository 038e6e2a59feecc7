python benchmarks/bench_configs.py --only C3 --steps 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c3_launches.csv python benchmarks/bench_configs.py --only C3 --steps 2 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/c3_launches.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
for r in rows[-14:]:
    print(r[ik][:90], r[iv])
PY
