R="python benchmarks/bench_configs.py --only R --steps 1"
ncu --set full --clock-control none --import-source on -k regex:resample -s 2 -c 1 -f -o gpurun_out/res_full $R > gpurun_out/ncu_r.log 2>&1; tail -1 gpurun_out/ncu_r.log
