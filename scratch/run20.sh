mkdir -p gpurun_out
python benchmarks/bench_configs.py --out gpurun_out/configs_r01.json > /dev/null 2>&1; echo rc=$?
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-200
