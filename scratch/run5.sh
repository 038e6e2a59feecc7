mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x -k "resample or load_audio" 2>&1 | tail -5
python benchmarks/bench_configs.py --only R --steps 5 2>&1 | grep -v Warning | cut -c1-500
