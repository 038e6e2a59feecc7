mkdir -p gpurun_out
./scratch/ubench/smsp > gpurun_out/smsp.txt 2>&1
python -m pytest tests -q -m gpu -x 2>&1 | tail -2 > gpurun_out/tests.txt
python bench.py --clips 4096 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_a.json 2>gpurun_out/bench_a.err
CMD="python bench.py --clips 512 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
ncu --metrics smsp__inst_executed.max,smsp__inst_executed.min,smsp__inst_executed.avg,smsp__warps_active.max,smsp__warps_active.min,smsp__warps_active.avg,smsp__cycles_active.avg,smsp__issue_active.max,smsp__issue_active.min,smsp__issue_active.avg --clock-control none -k regex:fast_logmel -s 2 -c 1 --csv --log-file gpurun_out/smsp_ncu.csv $CMD > gpurun_out/ncu_smsp.log 2>&1
cat gpurun_out/smsp.txt; cat gpurun_out/tests.txt; cut -c1-400 gpurun_out/bench_a.json; grep -v "^==" gpurun_out/smsp_ncu.csv | cut -d, -f13- | tail -12
