mkdir -p gpurun_out
C5="python benchmarks/bench_configs.py --only C5 --steps 1"
ncu --set full --clock-control none --import-source on -k regex:fast_istft -s 10 -c 1 -f -o gpurun_out/k3_full $C5 > gpurun_out/ncu_k3.log 2>&1
tail -1 gpurun_out/ncu_k3.log
