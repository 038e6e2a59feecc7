mkdir -p gpurun_out
CMD="python bench.py --clips 512 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:fast_logmel -s 2 -c 1 -f -o gpurun_out/k1_full $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
