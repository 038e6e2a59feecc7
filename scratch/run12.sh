C3="python benchmarks/bench_configs.py --only C3 --steps 1"
ncu --set full --clock-control none --import-source on -k regex:fast_logmel -s 6 -c 1 -f -o gpurun_out/k512_full $C3 > gpurun_out/ncu_c3.log 2>&1; tail -1 gpurun_out/ncu_c3.log
