CMD="python bench.py --clips 512 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fast_logmel -s 2 -c 1 -f -o gpurun_out/prof_spec $CMD > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
