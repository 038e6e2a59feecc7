mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; tail -c 300 gpurun_out/bench_r01_final.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r01_reference.json 2> gpurun_out/bench_r01_reference.err
python benchmarks/bench_configs.py --out gpurun_out/configs_r01.json > /dev/null 2>&1
CMD="python bench.py --clips 512 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fast_logmel -s 2 -c 1 -f -o gpurun_out/prof_r01_fast_logmel $CMD > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/ncu.log
# the other kernels of the path: one --set full capture each (largest-batch launch of the per-config script)
C3="python benchmarks/bench_configs.py --only C3 --steps 1"
$C3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fast_logmel -s 6 -c 1 -f -o gpurun_out/prof_r01_fast_logmel_512 $C3 > gpurun_out/ncu_c3.log 2>&1
C5="python benchmarks/bench_configs.py --only C5 --steps 1"
$C5 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fast_logmel -s 14 -c 1 -f -o gpurun_out/prof_r01_fast_logmel_1024 $C5 > gpurun_out/ncu_c5f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fast_istft -s 10 -c 1 -f -o gpurun_out/prof_r01_fast_istft_1024 $C5 > gpurun_out/ncu_c5i.log 2>&1
C4="python benchmarks/bench_configs.py --only C4 --steps 1"
$C4 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:istft_small -s 10 -c 1 -f -o gpurun_out/prof_r01_istft_small $C4 > gpurun_out/ncu_c4.log 2>&1
S="python benchmarks/bench_configs.py --only S --steps 1"
$S > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fast_stft -s 2 -c 1 -f -o gpurun_out/prof_r01_fast_stft_400 $S > gpurun_out/ncu_s.log 2>&1
R="python benchmarks/bench_configs.py --only R --steps 1"
$R > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:resample -s 2 -c 1 -f -o gpurun_out/prof_r01_resample $R > gpurun_out/ncu_r.log 2>&1
ls -la gpurun_out/prof_r01_*.ncu-rep | awk '{print $5, $9}'
cat gpurun_out/bench_r01_final.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])"
cat gpurun_out/bench_r01_reference.json | cut -c1-300
